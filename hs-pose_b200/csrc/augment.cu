// K10 — the on-device batched augmentation of HSPose.data_augment as ONE launch.
//
// Replaces the ~40 element-wise launches of the reference's network/HSPose.py:185-256 over
// datasets/data_augmentation.py (`defor_3D_bb_in_batch` :70-79, `defor_3D_rt_in_batch` :183-190,
// `defor_3D_bc_in_batch` :106-127, `defor_3D_pc` :134-140): four Bernoulli-gated deformations applied in
// sequence — bounding-box scaling in the object frame, rigid perturbation, the bowl / mug taper along y
// (categories 1 and 5; the new size comes from the tapered model points' extent), and per-point noise.
// The uniform random numbers are INPUTS (drawn by torch in the reference's order: gate_bb, gate_rt, gate_bc,
// ey_up, ey_down, gate_pc, defor), so the RNG stream is the caller's and parity is exact.
// CTA = object: the per-object pose chain is evaluated once (thread 0; the taper's extent by a block
// min/max reduction over the model points), then the threads walk the points.
#include "common.cuh"

namespace hsp {

constexpr int AUG_THREADS = 256;

struct AugObj {
  float R0[9], t0[3];      // pose used by the box scaling
  float scale[3];
  float Rr[9], dt[3];      // rigid perturbation
  float R1[9], t1[3];      // pose after the perturbation (used by the taper and the noise)
  float sy, ey_up, ey_down;
  int f_bb, f_rt, f_bc, f_pc;
};

__device__ __forceinline__ void matvec(const float* M, const float* v, float* o) {   // o = M v  (row-major M)
#pragma unroll
  for (int j = 0; j < 3; ++j) o[j] = M[3 * j] * v[0] + M[3 * j + 1] * v[1] + M[3 * j + 2] * v[2];
}
__device__ __forceinline__ void matTvec(const float* M, const float* v, float* o) {  // o = M^T v
#pragma unroll
  for (int k = 0; k < 3; ++k) o[k] = M[k] * v[0] + M[3 + k] * v[1] + M[6 + k] * v[2];
}

__global__ void __launch_bounds__(AUG_THREADS)
augment_kernel(const float* __restrict__ PC, const float* __restrict__ R, const float* __restrict__ t,
               const float* __restrict__ s, const float* __restrict__ mean_shape, const float* __restrict__ sym,
               const float* __restrict__ aug_bb, const float* __restrict__ aug_rt_t, const float* __restrict__ aug_rt_r,
               const float* __restrict__ model_point, const float* __restrict__ nocs_scale,
               const float* __restrict__ obj_id, const float* __restrict__ gates /*(B,4)*/,
               const float* __restrict__ ey /*(B,2)*/, const float* __restrict__ defor /*(B,N,3) in [0,1)*/,
               float p_bb, float p_rt, float p_bc, float p_pc, float pc_r, int N, int Nm, float* __restrict__ PC_out,
               float* __restrict__ R_out, float* __restrict__ t_out, float* __restrict__ s_out) {
  __shared__ AugObj o;
  __shared__ float s_mm[AUG_THREADS / 32][6];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    const float* g = gates + 4 * b;
    const float oid = obj_id[b];
    o.f_bb = g[0] < p_bb;
    o.f_rt = g[1] < p_rt;
    o.f_bc = (g[2] < p_bc) && (oid == 5.f || oid == 1.f);
    o.f_pc = g[3] < p_pc;
    for (int i = 0; i < 9; ++i) { o.R0[i] = R[9 * b + i]; o.Rr[i] = aug_rt_r[9 * b + i]; }
    const float* ab = aug_bb + 3 * b;
    const bool is_sym = sym[4 * b] == 1.f;
    const float sxz = (ab[0] + ab[2]) / 2.0f;
    o.scale[0] = is_sym ? sxz : ab[0];
    o.scale[1] = ab[1];
    o.scale[2] = is_sym ? sxz : ab[2];
    for (int j = 0; j < 3; ++j) { o.t0[j] = t[3 * b + j]; o.dt[j] = aug_rt_t[3 * b + j]; }
    // pose after the rigid perturbation: R1 = Rr R0, t1 = Rr (t0 + dt)
    if (o.f_rt) {
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k)
          o.R1[3 * j + k] = o.Rr[3 * j] * o.R0[k] + o.Rr[3 * j + 1] * o.R0[3 + k] + o.Rr[3 * j + 2] * o.R0[6 + k];
      const float tt[3] = {o.t0[0] + o.dt[0], o.t0[1] + o.dt[1], o.t0[2] + o.dt[2]};
      matvec(o.Rr, tt, o.t1);
    } else {
      for (int i = 0; i < 9; ++i) o.R1[i] = o.R0[i];
      for (int j = 0; j < 3; ++j) o.t1[j] = o.t0[j];
    }
    float sz[3];
    for (int j = 0; j < 3; ++j) {
      sz[j] = s[3 * b + j] + mean_shape[3 * b + j];
      if (o.f_bb) sz[j] *= o.scale[j];
    }
    o.sy = sz[1];
    o.ey_up = ey[2 * b] * 0.4f + 0.8f;          // rand * (1.2 - 0.8) + 0.8
    o.ey_down = ey[2 * b + 1] * 0.4f + 0.8f;
    for (int i = 0; i < 9; ++i) R_out[9 * b + i] = o.R1[i];
    for (int j = 0; j < 3; ++j) {
      t_out[3 * b + j] = o.t1[j];
      s_out[3 * b + j] = sz[j] - mean_shape[3 * b + j];       // overwritten below when the taper fires
    }
  }
  __syncthreads();
  if (o.f_bc) {   // new size = extent of the tapered (and box-scaled) model points x nocs_scale
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < Nm; i += AUG_THREADS) {
      float p[3];
      for (int j = 0; j < 3; ++j) p[j] = model_point[((size_t)b * Nm + i) * 3 + j] * (o.f_bb ? o.scale[j] : 1.f);
      const float f = (p[1] + o.sy / 2.0f) / o.sy * (o.ey_up - o.ey_down) + o.ey_down;
      p[0] *= f; p[2] *= f;
      for (int j = 0; j < 3; ++j) { mn[j] = fminf(mn[j], p[j]); mx[j] = fmaxf(mx[j], p[j]); }
    }
    for (int j = 0; j < 3; ++j)
      for (int off = 16; off > 0; off >>= 1) {
        mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], off));
        mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], off));
      }
    if ((tid & 31) == 0)
      for (int j = 0; j < 3; ++j) { s_mm[tid >> 5][j] = mn[j]; s_mm[tid >> 5][3 + j] = mx[j]; }
    __syncthreads();
    if (tid < 3) {
      float a = INFINITY, c = -INFINITY;
      for (int w = 0; w < AUG_THREADS / 32; ++w) { a = fminf(a, s_mm[w][tid]); c = fmaxf(c, s_mm[w][3 + tid]); }
      s_out[3 * b + tid] = (c - a) * nocs_scale[b] - mean_shape[3 * b + tid];
    }
  }
  for (int n = tid; n < N; n += AUG_THREADS) {
    const size_t pi = ((size_t)b * N + n) * 3;
    float p[3] = {PC[pi], PC[pi + 1], PC[pi + 2]};
    if (o.f_bb) {          // R0 (scale * R0^T (p - t0)) + t0
      const float d[3] = {p[0] - o.t0[0], p[1] - o.t0[1], p[2] - o.t0[2]};
      float q[3];
      matTvec(o.R0, d, q);
      for (int j = 0; j < 3; ++j) q[j] *= o.scale[j];
      matvec(o.R0, q, p);
      for (int j = 0; j < 3; ++j) p[j] += o.t0[j];
    }
    if (o.f_rt) {          // Rr (p + dt)
      const float d[3] = {p[0] + o.dt[0], p[1] + o.dt[1], p[2] + o.dt[2]};
      matvec(o.Rr, d, p);
    }
    if (o.f_bc) {          // taper along y in the object frame
      const float d[3] = {p[0] - o.t1[0], p[1] - o.t1[1], p[2] - o.t1[2]};
      float q[3];
      matTvec(o.R1, d, q);
      const float f = (q[1] + o.sy / 2.0f) / o.sy * (o.ey_up - o.ey_down) + o.ey_down;
      q[0] *= f; q[2] *= f;
      matvec(o.R1, q, p);
      for (int j = 0; j < 3; ++j) p[j] += o.t1[j];
    }
    if (o.f_pc) {
      for (int j = 0; j < 3; ++j) p[j] += defor[pi + j] * pc_r * (p[j] - o.t1[j]);
    }
    PC_out[pi] = p[0]; PC_out[pi + 1] = p[1]; PC_out[pi + 2] = p[2];
  }
}

}  // namespace hsp

extern "C" int hsp_augment(const float* PC, const float* R, const float* t, const float* s, const float* mean_shape,
                           const float* sym, const float* aug_bb, const float* aug_rt_t, const float* aug_rt_r,
                           const float* model_point, const float* nocs_scale, const float* obj_id, const float* gates,
                           const float* ey, const float* defor, float p_bb, float p_rt, float p_bc, float p_pc,
                           float pc_r, int B, int N, int Nm, float* PC_out, float* R_out, float* t_out, float* s_out,
                           void* stream) {
  using namespace hsp;
  if (!PC || !R || !t || !s || !mean_shape || !sym || !aug_bb || !aug_rt_t || !aug_rt_r || !model_point ||
      !nocs_scale || !obj_id || !gates || !ey || !defor || !PC_out || !R_out || !t_out || !s_out || B <= 0 || N <= 0 ||
      Nm <= 0)
    return HSP_EINVAL;
  augment_kernel<<<B, AUG_THREADS, 0, (cudaStream_t)stream>>>(PC, R, t, s, mean_shape, sym, aug_bb, aug_rt_t, aug_rt_r,
                                                           model_point, nocs_scale, obj_id, gates, ey, defor, p_bb,
                                                           p_rt, p_bc, p_pc, pc_r, N, Nm, PC_out, R_out, t_out, s_out);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
