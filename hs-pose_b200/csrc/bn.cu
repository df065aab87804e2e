// K6b — batch-statistics BatchNorm (+ReLU) over the channel axis of a (M, C)
// row-major activation matrix, forward and backward, as streaming HBM-bound
// kernels (bf16 or fp32 storage, fp32 arithmetic).
//
// Replaces, for the dense per-point MLPs of the reference
// (network/fs_net_repo/FaceRecon.py:27-29,38-68,89-95, PoseR.py:22-29,
// PoseTs.py:24-34), the sequence  BatchNorm1d(train) -> ReLU  that PyTorch runs
// as 3 kernels forward (statistics, transform, clamp) and 3 backward
// (threshold, reduce, elementwise).  Here: forward = one read for the
// statistics + one read/write for the fused normalise+ReLU; backward = one
// 2-read reduction + one 2-read/1-write pass; the ReLU mask is recomputed from
// x, so y is never re-read.
//
// Thread mapping: a thread owns 8 consecutive channels (one 16-byte bf16
// vector) of a column tile and walks rows; per-CTA partial sums go to a
// workspace and are combined in a fixed order (double) — deterministic, no
// float atomics.  x may be a column slice of a wider matrix (row stride ldx).
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {

constexpr int BN_THREADS = 256;
constexpr int BN_VEC = 8;

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void round(const float* v, float* r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i];
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ void round(const float* v, float* r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  }
};

struct BnGeom {
  int ct;        // vector-columns per CTA tile (power of two <= 32 dividing C/8)
  int rl;        // row lanes per CTA = BN_THREADS / ct
  int ctiles;    // column tiles
  int rchunks;   // row chunks (grid.y)
};
static BnGeom bn_geom(int M, int C) {
  BnGeom g;
  const int cv = C / BN_VEC;
  g.ct = 32;
  while (g.ct > 1 && (cv % g.ct) != 0) g.ct >>= 1;
  g.rl = BN_THREADS / g.ct;
  g.ctiles = cv / g.ct;
  long want = (148L * 6 + g.ctiles - 1) / g.ctiles;
  long maxr = (M + g.rl * 4 - 1) / (g.rl * 4);
  if (want > maxr) want = maxr;
  if (want < 1) want = 1;
  g.rchunks = (int)want;
  return g;
}

// Two accumulators per channel: A = sum a_i, Bq = sum b_i, where
//   MODE 0 (forward stats) : a = x,   b = x*x
//   MODE 1 (backward)      : a = dy', b = dy' * x,  dy' = relu-masked dy (mask = the forward's own
//                            y = fma(x, scale, shift) <= 0); dgamma = invstd * (Bq - mean * A) is
//                            formed in the finalize, so no per-element normalisation is needed here.
// Four rows are in flight per thread (8 x 16-byte loads): these kernels are pure HBM streams and
// need the bytes in flight more than they need occupancy.
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy,
                 const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int relu, int M,
                 int C, int ct, float* __restrict__ partial) {
  __shared__ float s_red[BN_THREADS * 2 * BN_VEC];
  const int rl = BN_THREADS / ct;
  const int col = threadIdx.x % ct, rlane = threadIdx.x / ct;
  const int c0 = (blockIdx.x * ct + col) * BN_VEC;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float A[BN_VEC], Bq[BN_VEC], sc[BN_VEC], sh[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    A[i] = Bq[i] = 0.f;
    sc[i] = 0.f; sh[i] = 1.f;                    // mask never fires without ReLU
    if (MODE == 1 && relu) {
      const float ga = gamma ? gamma[c0 + i] : 1.f, be = beta ? beta[c0 + i] : 0.f;
      sc[i] = ga * invstd[c0 + i];               // same expressions as bn_finalize_fwd_kernel
      sh[i] = be - mean[c0 + i] * ga * invstd[c0 + i];
    }
  }
  int r = r0 + rlane;
  for (; r + 3 * rl < r1; r += 4 * rl) {
    float xv[4][BN_VEC], gv[4][BN_VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      Vec8<T>::load(x + (size_t)(r + u * rl) * ldx + c0, xv[u]);
      if (MODE == 1) Vec8<T>::load(dy + (size_t)(r + u * rl) * lddy + c0, gv[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) {
        if (MODE == 0) {
          A[i] += xv[u][i]; Bq[i] = fmaf(xv[u][i], xv[u][i], Bq[i]);
        } else {
          const float g = fmaf(xv[u][i], sc[i], sh[i]) <= 0.f ? 0.f : gv[u][i];
          A[i] += g; Bq[i] = fmaf(g, xv[u][i], Bq[i]);
        }
      }
  }
  for (; r < r1; r += rl) {
    float xv[BN_VEC], gv[BN_VEC];
    Vec8<T>::load(x + (size_t)r * ldx + c0, xv);
    if (MODE == 1) Vec8<T>::load(dy + (size_t)r * lddy + c0, gv);
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) {
      if (MODE == 0) {
        A[i] += xv[i]; Bq[i] = fmaf(xv[i], xv[i], Bq[i]);
      } else {
        const float g = fmaf(xv[i], sc[i], sh[i]) <= 0.f ? 0.f : gv[i];
        A[i] += g; Bq[i] = fmaf(g, xv[i], Bq[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    s_red[(i * 2) * BN_THREADS + threadIdx.x] = A[i];
    s_red[(i * 2 + 1) * BN_THREADS + threadIdx.x] = Bq[i];
  }
  __syncthreads();
  // thread t < ct*16 sums column (t % ct), quantity q = t / ct over the rl row lanes (fixed order)
  for (int t = threadIdx.x; t < ct * 2 * BN_VEC; t += BN_THREADS) {
    const int cc = t % ct, q = t / ct;
    float s = 0.f;
    for (int l = 0; l < rl; ++l) s += s_red[q * BN_THREADS + l * ct + cc];
    const int ch = (blockIdx.x * ct + cc) * BN_VEC + (q >> 1);
    partial[((size_t)blockIdx.y * 2 + (q & 1)) * C + ch] = s;
  }
}

// Fixed-order sum of the per-CTA partials: one CTA of 1024 threads owns BN_FC = 8 channels,
// thread (cx = tid % 8, ry = tid / 8) adds chunks ry, ry+128, ... (<= 9 independent loads at
// M = 131584, all in flight at once: the kernel is pure load latency, so it is kept to ONE round trip),
// then the 128 group sums are added in a fixed order (16 threads per channel over 8 groups each,
// then 16 in order) — deterministic.  Was 32 groups x 32 sequential loads: 12-13 us per call, 44 calls a step.
constexpr int BN_FC = 8, BN_FG = 128;
__device__ __forceinline__ bool bn_sum_partials(const float* __restrict__ partial, int rchunks,
                                                int C, int& c, double& s, double& q, int ldp = 0) {
  if (ldp == 0) ldp = C;    // row pitch of a partial plane (>= C when the planes are column slices)
  __shared__ double sh[2][BN_FG][BN_FC];
  const int cx = threadIdx.x % BN_FC, ry = threadIdx.x / BN_FC;
  c = blockIdx.x * BN_FC + cx;
  double a = 0.0, b = 0.0;
  if (c < C) {
#pragma unroll 8
    for (int r = ry; r < rchunks; r += BN_FG) {
      a += (double)__ldg(partial + ((size_t)r * 2) * ldp + c);
      b += (double)__ldg(partial + ((size_t)r * 2 + 1) * ldp + c);
    }
  }
  sh[0][ry][cx] = a;
  sh[1][ry][cx] = b;
  __syncthreads();
  if (ry < 16) {            // second level: groups 8 ry .. 8 ry + 7, in order
    a = 0.0; b = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) { a += sh[0][8 * ry + g][cx]; b += sh[1][8 * ry + g][cx]; }
  }
  __syncthreads();
  if (ry < 16) { sh[0][ry][cx] = a; sh[1][ry][cx] = b; }
  __syncthreads();
  s = 0.0; q = 0.0;
  if (ry != 0 || c >= C) return false;
#pragma unroll
  for (int g = 0; g < 16; ++g) { s += sh[0][g][cx]; q += sh[1][g][cx]; }
  return true;
}

// Forward finalize: mean / invstd, fused affine (scale, shift), running-stat update.
__global__ void __launch_bounds__(BN_FC * BN_FG)
bn_finalize_fwd_kernel(const float* __restrict__ partial, int rchunks, int M, int C,
                       float eps, float momentum, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float* __restrict__ mean,
                       float* __restrict__ invstd, float* __restrict__ scale,
                       float* __restrict__ shift, float* __restrict__ running_mean,
                       float* __restrict__ running_var, int ldp = 0) {
  int c;
  double s, q;
  if (!bn_sum_partials(partial, rchunks, C, c, s, q, ldp)) return;
  const double m = s / M;
  double var = q / M - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)m;
  invstd[c] = is;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * is;
  shift[c] = b - (float)m * g * is;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// Backward finalize: dbeta = sum dy', dgamma = sum dy' xhat = invstd * (sum dy' x - mean * sum dy').
// mean == NULL: plain column sums (first plane -> dbeta, second -> dgamma).
__global__ void __launch_bounds__(BN_FC * BN_FG)
bn_finalize_bwd_kernel(const float* __restrict__ partial, int rchunks, int C,
                       const float* __restrict__ mean, const float* __restrict__ invstd,
                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int c;
  double s, q;
  if (!bn_sum_partials(partial, rchunks, C, c, s, q)) return;
  dbeta[c] = (float)s;
  dgamma[c] = mean ? (float)((double)invstd[c] * (q - (double)mean[c] * s)) : (float)q;
}

// y = relu?(x * scale + shift)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ scale,
                const float* __restrict__ shift, int relu, int M, int C, int ct,
                T* __restrict__ y, int ldy) {
  const int rl = BN_THREADS / ct;
  const int col = threadIdx.x % ct, rlane = threadIdx.x / ct;
  const int c0 = (blockIdx.x * ct + col) * BN_VEC;
  float sc[BN_VEC], sh[BN_VEC];
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) { sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i]; }
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
#pragma unroll 2
  for (int r = r0 + rlane; r < r1; r += rl) {
    float v[BN_VEC];
    Vec8<T>::load(x + (size_t)r * ldx + c0, v);
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) {
      v[i] = fmaf(v[i], sc[i], sh[i]);
      if (relu) v[i] = fmaxf(v[i], 0.f);
    }
    Vec8<T>::store(y + (size_t)r * ldy + c0, v);
  }
}

// dx = gamma*invstd * (dy' - dbeta/M - xhat * dgamma/M), folded per channel into
//   dx = A1 * dy' + B1 * x + C1,   A1 = gamma*invstd, B1 = -A1*invstd*dgamma/M, C1 = -A1*dbeta/M - B1*mean
// with the forward's own mask y = fma(x, scale, shift) <= 0.  Four rows in flight per thread.
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu,
                    int M, int C, int ct, T* __restrict__ dx, int lddx,
                    float* __restrict__ colsum_partial) {
  __shared__ float s_cs[BN_THREADS * BN_VEC];
  const int rl = BN_THREADS / ct;
  const int col = threadIdx.x % ct, rlane = threadIdx.x / ct;
  const int c0 = (blockIdx.x * ct + col) * BN_VEC;
  float A1[BN_VEC], B1[BN_VEC], C1[BN_VEC], sc[BN_VEC], sh[BN_VEC], cs[BN_VEC];
  const float invM = 1.f / (float)M;
#pragma unroll
  for (int i = 0; i < BN_VEC; ++i) {
    const float mu = mean[c0 + i], is = invstd[c0 + i];
    const float ga = gamma ? gamma[c0 + i] : 1.f, be = beta ? beta[c0 + i] : 0.f;
    A1[i] = ga * is;
    B1[i] = -A1[i] * is * (dgamma[c0 + i] * invM);
    C1[i] = -A1[i] * (dbeta[c0 + i] * invM) - B1[i] * mu;
    sc[i] = relu ? ga * is : 0.f;                 // same expressions as bn_finalize_fwd_kernel
    sh[i] = relu ? be - mu * ga * is : 1.f;
    cs[i] = 0.f;
  }
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  int r = r0 + rlane;
  for (; r + 3 * rl < r1; r += 4 * rl) {
    float xv[4][BN_VEC], gv[4][BN_VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      Vec8<T>::load(x + (size_t)(r + u * rl) * ldx + c0, xv[u]);
      Vec8<T>::load(dy + (size_t)(r + u * rl) * lddy + c0, gv[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) {
        const float g = fmaf(xv[u][i], sc[i], sh[i]) <= 0.f ? 0.f : gv[u][i];
        xv[u][i] = fmaf(A1[i], g, fmaf(B1[i], xv[u][i], C1[i]));
      }
      Vec8<T>::store(dx + (size_t)(r + u * rl) * lddx + c0, xv[u]);
      if (colsum_partial) {   // column sums of dx AS STORED (rounded to T): the bias gradient
        float rv[BN_VEC];     // of the Linear in front of this BatchNorm (= sum_rows dY)
        Vec8<T>::round(xv[u], rv);
#pragma unroll
        for (int i = 0; i < BN_VEC; ++i) cs[i] += rv[i];
      }
    }
  }
  for (; r < r1; r += rl) {
    float xv[BN_VEC], gv[BN_VEC];
    Vec8<T>::load(x + (size_t)r * ldx + c0, xv);
    Vec8<T>::load(dy + (size_t)r * lddy + c0, gv);
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) {
      const float g = fmaf(xv[i], sc[i], sh[i]) <= 0.f ? 0.f : gv[i];
      xv[i] = fmaf(A1[i], g, fmaf(B1[i], xv[i], C1[i]));
    }
    Vec8<T>::store(dx + (size_t)r * lddx + c0, xv);
    if (colsum_partial) {
      float rv[BN_VEC];
      Vec8<T>::round(xv, rv);
#pragma unroll
      for (int i = 0; i < BN_VEC; ++i) cs[i] += rv[i];
    }
  }
  if (colsum_partial) {
#pragma unroll
    for (int i = 0; i < BN_VEC; ++i) s_cs[i * BN_THREADS + threadIdx.x] = cs[i];
    __syncthreads();
    for (int t = threadIdx.x; t < ct * BN_VEC; t += BN_THREADS) {
      const int cc = t % ct, q = t / ct;
      float a = 0.f;
      for (int l = 0; l < rl; ++l) a += s_cs[q * BN_THREADS + l * ct + cc];
      const int ch = (blockIdx.x * ct + cc) * BN_VEC + q;
      colsum_partial[((size_t)blockIdx.y * 2) * C + ch] = a;   // same layout as bn_reduce partials
      colsum_partial[((size_t)blockIdx.y * 2 + 1) * C + ch] = 0.f;
    }
  }
}

static bool bn_bad(const void* x, int M, int C, int ld, int dtype) {
  const int esz = dtype == HSP_DTYPE_BF16 ? 2 : 4;
  return !x || M <= 0 || C <= 0 || (C % BN_VEC) != 0 || ld < C || (ld * esz) % 16 != 0 ||
         ((uintptr_t)x % 16) != 0 || (dtype != HSP_DTYPE_F32 && dtype != HSP_DTYPE_BF16);
}

}  // namespace hsp

extern "C" size_t hsp_bn_workspace_bytes(int M, int C) {
  using namespace hsp;
  if (M <= 0 || C <= 0 || (C % BN_VEC) != 0) return 0;
  return ((size_t)bn_geom(M, C).rchunks * 2 * C + C) * sizeof(float);
}

extern "C" int hsp_bn_relu_fwd(const void* x, int ldx, int M, int C, int dtype, const float* gamma,
                               const float* beta, float eps, float momentum, int relu,
                               float* running_mean, float* running_var, float* mean,
                               float* invstd, float* scale_shift, void* y, int ldy,
                               void* workspace, size_t workspace_bytes, void* stream) {
  using namespace hsp;
  if (bn_bad(x, M, C, ldx, dtype) || bn_bad(y, M, C, ldy, dtype) || !mean || !invstd || !scale_shift)
    return HSP_EINVAL;
  if (!workspace || workspace_bytes < hsp_bn_workspace_bytes(M, C)) return HSP_EWORKSPACE;
  const BnGeom g = bn_geom(M, C);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(g.ctiles, g.rchunks);
  float* part = (float*)workspace;
  float* scale = scale_shift;
  float* shift = scale_shift + C;
  if (dtype == HSP_DTYPE_BF16)
    bn_reduce_kernel<__nv_bfloat16, 0><<<grid, BN_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, M, C, g.ct, part);
  else
    bn_reduce_kernel<float, 0><<<grid, BN_THREADS, 0, st>>>(
        (const float*)x, ldx, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, M, C, g.ct, part);
  HSP_LAUNCH_CHECK();
  bn_finalize_fwd_kernel<<<(C + BN_FC - 1) / BN_FC, BN_FC * BN_FG, 0, st>>>(part, g.rchunks, M, C, eps, momentum, gamma,
                                                          beta, mean, invstd, scale, shift,
                                                          running_mean, running_var);
  HSP_LAUNCH_CHECK();
  if (dtype == HSP_DTYPE_BF16)
    bn_apply_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, scale, shift, relu, M, C, g.ct, (__nv_bfloat16*)y, ldy);
  else
    bn_apply_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)x, ldx, scale, shift, relu, M, C,
                                                        g.ct, (float*)y, ldy);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_bn_relu_bwd(const void* x, int ldx, const void* dy, int lddy, int M, int C,
                               int dtype, const float* gamma, const float* beta, const float* mean,
                               const float* invstd, int relu, float* dgamma, float* dbeta, void* dx,
                               int lddx, float* dx_colsum, void* workspace, size_t workspace_bytes,
                               void* stream) {
  using namespace hsp;
  if (bn_bad(x, M, C, ldx, dtype) || bn_bad(dy, M, C, lddy, dtype) || bn_bad(dx, M, C, lddx, dtype) ||
      !mean || !invstd || !dgamma || !dbeta)
    return HSP_EINVAL;
  if (!workspace || workspace_bytes < hsp_bn_workspace_bytes(M, C)) return HSP_EWORKSPACE;
  const BnGeom g = bn_geom(M, C);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(g.ctiles, g.rchunks);
  float* part = (float*)workspace;
  if (dtype == HSP_DTYPE_BF16)
    bn_reduce_kernel<__nv_bfloat16, 1><<<grid, BN_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, lddy, mean, invstd, gamma, beta, relu,
        M, C, g.ct, part);
  else
    bn_reduce_kernel<float, 1><<<grid, BN_THREADS, 0, st>>>((const float*)x, ldx, (const float*)dy,
                                                            lddy, mean, invstd, gamma, beta, relu, M,
                                                            C, g.ct, part);
  HSP_LAUNCH_CHECK();
  bn_finalize_bwd_kernel<<<(C + BN_FC - 1) / BN_FC, BN_FC * BN_FG, 0, st>>>(part, g.rchunks, C, mean, invstd,
                                                                        dgamma, dbeta);
  HSP_LAUNCH_CHECK();
  if (dtype == HSP_DTYPE_BF16)
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, lddy, mean, invstd, gamma, beta, dgamma,
        dbeta, relu, M, C, g.ct, (__nv_bfloat16*)dx, lddx, dx_colsum ? part : nullptr);
  else
    bn_bwd_apply_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)x, ldx, (const float*)dy, lddy,
                                                            mean, invstd, gamma, beta, dgamma, dbeta,
                                                            relu, M, C, g.ct, (float*)dx, lddx,
                                                            dx_colsum ? part : nullptr);
  HSP_LAUNCH_CHECK();
  if (dx_colsum) {   // fixed-order sum of the per-CTA column sums (the second partial plane is zero)
    bn_finalize_bwd_kernel<<<(C + BN_FC - 1) / BN_FC, BN_FC * BN_FG, 0, st>>>(
        part, g.rchunks, C, nullptr, nullptr, (float*)workspace + (size_t)g.rchunks * 2 * C, dx_colsum);
    HSP_LAUNCH_CHECK();
  }
  return HSP_OK;
}

// Forward with the statistics partials already produced by the GEMM epilogue (gemm_tc.cu): finalize
// (fixed-order sums) + normalise/ReLU.  `partials` is (nblocks, 2, C) with row pitch ldp >= C: column sums and sums of squares.
extern "C" int hsp_bn_apply_fwd(const void* x, int ldx, int M, int C, int dtype, const float* partials,
                                int nblocks, int ldp, const float* gamma, const float* beta, float eps,
                                float momentum, int relu, float* running_mean, float* running_var,
                                float* mean, float* invstd, float* scale_shift, void* y, int ldy,
                                void* stream) {
  using namespace hsp;
  if (bn_bad(x, M, C, ldx, dtype) || bn_bad(y, M, C, ldy, dtype) || !mean || !invstd || !scale_shift ||
      !partials || nblocks <= 0 || ldp < C)
    return HSP_EINVAL;
  const BnGeom g = bn_geom(M, C);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(g.ctiles, g.rchunks);
  float* scale = scale_shift;
  float* shift = scale_shift + C;
  bn_finalize_fwd_kernel<<<(C + BN_FC - 1) / BN_FC, BN_FC * BN_FG, 0, st>>>(
      partials, nblocks, M, C, eps, momentum, gamma, beta, mean, invstd, scale, shift, running_mean, running_var,
      ldp);
  HSP_LAUNCH_CHECK();
  if (dtype == HSP_DTYPE_BF16)
    bn_apply_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, scale, shift, relu, M, C, g.ct, (__nv_bfloat16*)y, ldy);
  else
    bn_apply_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)x, ldx, scale, shift, relu, M, C,
                                                        g.ct, (float*)y, ldy);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
