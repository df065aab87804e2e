// K9 — gradient clipping + optimiser step over the flat parameter / gradient buffers as three
// launches: the reference runs `clip_grad_norm_` and `Ranger.step`
// (tools/torch_utils/solver/ranger2020.py:135-246, selected at tools/solver_utils.py:49-50) as a few
// hundred per-parameter element-wise launches with fp32 copies of every tensor.
//
//   optim_sqsum_kernel   per-CTA partial sums of g^2 (fixed order: deterministic norm)
//   optim_prep_kernel    one CTA: ||g||, the clip coefficient min(1, max_norm / (||g|| + 1e-6))
//                        (torch.nn.utils.clip_grad_norm_), step += 1 ON THE DEVICE (the whole train step
//                        is one CUDA graph: nothing here may depend on a host-side counter) and the
//                        step-dependent scalars of RAdam (N_sma, step size) / Adam (bias corrections)
//   optim_update_kernel  warp = one "segment" of the flat buffer: a row of a >= 2-D parameter
//                        (gradient centralisation subtracts the row mean: ranger2020.py:30-40 with
//                        gc_loc = True, gc_conv_only = False) or a whole 1-D parameter; moments, RAdam /
//                        Adam update and, every k-th step, the Lookahead interpolation with the slow
//                        weights — one read and one write of every state tensor.
// The learning rate is read from device memory so a scheduler can change it without re-capturing.
#include "common.cuh"

namespace hsp {
namespace optim {

constexpr int KIND_ADAM = 0, KIND_RANGER = 1;
constexpr int SQ_CTAS = 592, SQ_THREADS = 256;

struct Scalars {          // written by optim_prep_kernel, read by optim_update_kernel
  float clip_coef;        // gradient scale
  float step_size;        // RAdam: rectification / (1 - beta1^t);  Adam: 1 / (1 - beta1^t)
  float denom_scale;      // Adam: 1 / sqrt(1 - beta2^t);  Ranger: 1
  int adaptive;           // RAdam: N_sma > threshold (divide by sqrt(v) + eps)
  int lookahead;          // this step interpolates with the slow weights
  float grad_norm;        // ||g|| before clipping (what clip_grad_norm_ returns)
};

__global__ void __launch_bounds__(SQ_THREADS)
optim_sqsum_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
  __shared__ float sh[SQ_THREADS / 32];
  float s = 0.f;
  const long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = (long)blockIdx.x * SQ_THREADS + threadIdx.x; i < n4; i += (long)gridDim.x * SQ_THREADS) {
    const float4 v = __ldg(g4 + i);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0) for (long i = n4 * 4 + threadIdx.x; i < n; i += SQ_THREADS) s = fmaf(g[i], g[i], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SQ_THREADS / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32)
optim_prep_kernel(const float* __restrict__ partial, int nparts, float max_norm, int kind, float beta1, float beta2,
                  int k_look, int nsma_threshold, int* __restrict__ step, Scalars* __restrict__ sc) {
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  for (int i = 0; i < nparts; ++i) tot += (double)partial[i];
  const float norm = (float)sqrt(tot);
  sc->grad_norm = norm;
  sc->clip_coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.f) : 1.f;
  const int t = *step + 1;
  *step = t;
  const double b1t = pow((double)beta1, (double)t), b2t = pow((double)beta2, (double)t);
  if (kind == KIND_RANGER) {
    const double nmax = 2.0 / (1.0 - (double)beta2) - 1.0;
    const double nsma = nmax - 2.0 * t * b2t / (1.0 - b2t);
    const bool adaptive = nsma > (double)nsma_threshold;
    double ss;
    if (adaptive)
      ss = sqrt((1.0 - b2t) * (nsma - 4.0) / (nmax - 4.0) * (nsma - 2.0) / nsma * nmax / (nmax - 2.0)) / (1.0 - b1t);
    else
      ss = 1.0 / (1.0 - b1t);
    sc->step_size = (float)ss;
    sc->denom_scale = 1.f;
    sc->adaptive = adaptive ? 1 : 0;
    sc->lookahead = (t % k_look) == 0 ? 1 : 0;
  } else {
    sc->step_size = (float)(1.0 / (1.0 - b1t));
    sc->denom_scale = (float)(1.0 / sqrt(1.0 - b2t));
    sc->adaptive = 1;
    sc->lookahead = 0;
  }
}

// seg_off[s], seg_len[s] (len < 0: -len elements WITHOUT gradient centralisation)
__global__ void __launch_bounds__(256)
optim_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    float* __restrict__ slow, const int* __restrict__ seg_off, const int* __restrict__ seg_len, int nseg,
                    const Scalars* __restrict__ sc, const float* __restrict__ lr_dev, int kind, float beta1, float beta2,
                    float eps, float wd, float alpha) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nseg) return;
  const int off = seg_off[s];
  int len = seg_len[s];
  const bool gc = kind == KIND_RANGER && len > 0;
  if (len < 0) len = -len;
  const float coef = sc->clip_coef, lr = *lr_dev, step_size = sc->step_size, dscale = sc->denom_scale;
  const bool adaptive = sc->adaptive != 0, look = sc->lookahead != 0;
  float mean = 0.f;
  if (gc) {
    float t = 0.f;
    for (int i = lane; i < len; i += 32) t += g[off + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    mean = coef * t / (float)len;
  }
  for (int i = lane; i < len; i += 32) {
    const int e = off + i;
    const float gr = coef * g[e] - mean;
    const float vv = beta2 * v[e] + (1.f - beta2) * gr * gr;
    const float mm = beta1 * m[e] + (1.f - beta1) * gr;
    v[e] = vv;
    m[e] = mm;
    float pw = p[e];
    float G = adaptive ? mm / (sqrtf(vv) * dscale + eps) : mm;
    if (wd != 0.f) G += wd * pw;
    pw -= step_size * lr * G;
    if (look) {
      const float sl = slow[e] + alpha * (pw - slow[e]);
      slow[e] = sl;
      pw = sl;
    }
    p[e] = pw;
  }
}

}  // namespace optim
}  // namespace hsp

extern "C" size_t hsp_optim_workspace_bytes(void) {
  return (size_t)hsp::optim::SQ_CTAS * sizeof(float) + sizeof(hsp::optim::Scalars) + 64;
}

extern "C" int hsp_optim_step(int kind, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                              float* slow, long n, const int* seg_off, const int* seg_len, int nseg, int* step,
                              const float* lr, float beta1, float beta2, float eps, float weight_decay,
                              float clip_max_norm, float alpha, int k, int nsma_threshold, float* grad_norm_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  using namespace hsp;
  using namespace hsp::optim;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !seg_off || !seg_len || !step || !lr || n <= 0 || nseg <= 0 ||
      (kind != KIND_ADAM && kind != KIND_RANGER) || (kind == KIND_RANGER && (!slow || k < 1)))
    return HSP_EINVAL;
  if (!workspace || workspace_bytes < hsp_optim_workspace_bytes()) return HSP_EWORKSPACE;
  if (((uintptr_t)grad % 16) != 0) return HSP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  Scalars* sc = (Scalars*)(((uintptr_t)(partial + SQ_CTAS) + 15) & ~(uintptr_t)15);
  int nparts = SQ_CTAS;
  const long per = (long)SQ_THREADS * 4;
  if ((n + per - 1) / per < nparts) nparts = (int)((n + per - 1) / per);
  optim_sqsum_kernel<<<nparts, SQ_THREADS, 0, st>>>(grad, n, partial);
  HSP_LAUNCH_CHECK();
  optim_prep_kernel<<<1, 32, 0, st>>>(partial, nparts, clip_max_norm, kind, beta1, beta2, k, nsma_threshold, step, sc);
  HSP_LAUNCH_CHECK();
  optim_update_kernel<<<(nseg + 7) / 8, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, slow, seg_off, seg_len, nseg, sc,
                                                     lr, kind, beta1, beta2, eps, weight_decay, alpha);
  HSP_LAUNCH_CHECK();
  if (grad_norm_out &&
      cudaMemcpyAsync(grad_norm_out, &sc->grad_norm, sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return HSP_ELAUNCH;
  return HSP_OK;
}
