// K3 / K4 — fused receptive-field gather + direction kernel + max over
// neighbours + mean over supports (+ centre term), forward and backward.
//
// Replaces HSlayer_surface.graph_conv (reference gcn3d.py:92-107) and
// HS_layer.graph_conv (gcn3d.py:158-181) together with the gathers they call
// (indexing_neighbor_new gcn3d.py:39-47, get_neighbor_direction_norm :49-59).
// The (B,N,k,S*C) tensors theta / gathered support / product of the reference
// are never written: one thread owns one output channel c of one point, keeps
// the S column-normalised support directions of that channel in registers
// (3*S floats) and the S running maxima in registers, and streams the k
// neighbour rows of the support matrix P[..., C:] straight from L2 (each warp
// reads 128 contiguous bytes per (neighbour, support)).  Unit directions of
// the tile's (point, neighbour) pairs are computed once per CTA into shared
// memory and broadcast.
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {

constexpr int GC_THREADS = 128;  // channels per CTA slice
constexpr int GC_PT = 8;         // points per CTA tile
constexpr int GC_MAXK = 64;

__device__ __forceinline__ float ld_p(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_p(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}

// F.normalize(nbr - centre, dim=-1): x / max(||x||_2, 1e-12)   (gcn3d.py:53-55)
__device__ __forceinline__ void unit_dir(const float* __restrict__ xb, int i, int j, float* r) {
  float rx = __fsub_rn(xb[3 * j], xb[3 * i]);
  float ry = __fsub_rn(xb[3 * j + 1], xb[3 * i + 1]);
  float rz = __fsub_rn(xb[3 * j + 2], xb[3 * i + 2]);
  float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
  float den = fmaxf(nrm, 1e-12f);
  r[0] = __fdiv_rn(rx, den);
  r[1] = __fdiv_rn(ry, den);
  r[2] = __fdiv_rn(rz, den);
}

// Stage idx and unit directions of a tile of points into shared memory.
__device__ __forceinline__ void stage_tile(const float* __restrict__ xb,
                                           const int32_t* __restrict__ idx_b, int i0, int npts,
                                           int k, int* s_idx, float* s_r) {
  for (int p = threadIdx.x; p < npts * k; p += blockDim.x) {
    int i = i0 + p / k;
    int nb = idx_b[(size_t)i * k + (p % k)];
    s_idx[p] = nb;
    float r[3];
    unit_dir(xb, i, nb, r);
    s_r[3 * p] = r[0];
    s_r[3 * p + 1] = r[1];
    s_r[3 * p + 2] = r[2];
  }
}

template <int S>
__global__ void __launch_bounds__(GC_THREADS)
surface_conv_fwd_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                        const float* __restrict__ dirn, int N, int k, int C,
                        float* __restrict__ out, uint8_t* __restrict__ argmax) {
  __shared__ int s_idx[GC_PT * GC_MAXK];
  __shared__ float s_r[GC_PT * GC_MAXK * 3];
  const int b = blockIdx.y, i0 = blockIdx.x * GC_PT;
  const int npts = min(GC_PT, N - i0);
  const int c = blockIdx.z * GC_THREADS + threadIdx.x;
  const int SC = S * C;
  stage_tile(xyz + (size_t)b * N * 3, idx + (size_t)b * N * k, i0, npts, k, s_idx, s_r);
  float dx[S], dy[S], dz[S];
  if (c < C) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      dx[s] = dirn[s * C + c];
      dy[s] = dirn[SC + s * C + c];
      dz[s] = dirn[2 * SC + s * C + c];
    }
  }
  __syncthreads();
  if (c >= C) return;
  for (int p = 0; p < npts; ++p) {
    float acc[S];
    int am[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s] = 0.0f; am[s] = 255; }  // max_n relu(x_n) = max(0, max_n x_n)
    if (argmax) {
      for (int n = 0; n < k; ++n) {
        const float rx = s_r[3 * (p * k + n)], ry = s_r[3 * (p * k + n) + 1],
                    rz = s_r[3 * (p * k + n) + 2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float th = fmaf(rz, dz[s], fmaf(ry, dy[s], rx * dx[s]));
          if (th > acc[s]) { acc[s] = th; am[s] = n; }
        }
      }
    } else {
      for (int n = 0; n < k; ++n) {
        const float rx = s_r[3 * (p * k + n)], ry = s_r[3 * (p * k + n) + 1],
                    rz = s_r[3 * (p * k + n) + 2];
#pragma unroll
        for (int s = 0; s < S; ++s)
          acc[s] = fmaxf(acc[s], fmaf(rz, dz[s], fmaf(ry, dy[s], rx * dx[s])));
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) sum += acc[s];
    const size_t row = (size_t)b * N + i0 + p;
    out[row * C + c] = __fdiv_rn(sum, (float)S);
    if (argmax) {
#pragma unroll
      for (int s = 0; s < S; ++s) argmax[row * SC + s * C + c] = (uint8_t)am[s];
    }
  }
}

template <int S, typename TP>
__global__ void __launch_bounds__(GC_THREADS)
graph_conv_fwd_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                      const float* __restrict__ dirn, const TP* __restrict__ P, int N, int k,
                      int C, float* __restrict__ out, uint8_t* __restrict__ argmax) {
  __shared__ int s_idx[GC_PT * GC_MAXK];
  __shared__ float s_r[GC_PT * GC_MAXK * 3];
  const int b = blockIdx.y, i0 = blockIdx.x * GC_PT;
  const int npts = min(GC_PT, N - i0);
  const int c = blockIdx.z * GC_THREADS + threadIdx.x;
  const int SC = S * C, LD = (S + 1) * C;
  stage_tile(xyz + (size_t)b * N * 3, idx + (size_t)b * N * k, i0, npts, k, s_idx, s_r);
  float dx[S], dy[S], dz[S];
  if (c < C) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      dx[s] = dirn[s * C + c];
      dy[s] = dirn[SC + s * C + c];
      dz[s] = dirn[2 * SC + s * C + c];
    }
  }
  __syncthreads();
  if (c >= C) return;
  const TP* Pb = P + (size_t)b * N * LD;
  for (int p = 0; p < npts; ++p) {
    float acc[S];
    int am[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s] = -INFINITY; am[s] = 0; }
#pragma unroll 2
    for (int n = 0; n < k; ++n) {
      const TP* sup = Pb + (size_t)s_idx[p * k + n] * LD + C + c;
      float v[S];
#pragma unroll
      for (int s = 0; s < S; ++s) v[s] = ld_p(sup + s * C);
      const float rx = s_r[3 * (p * k + n)], ry = s_r[3 * (p * k + n) + 1],
                  rz = s_r[3 * (p * k + n) + 2];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float th = fmaxf(fmaf(rz, dz[s], fmaf(ry, dy[s], rx * dx[s])), 0.0f);
        float a = th * v[s];
        if (a > acc[s]) { acc[s] = a; am[s] = n; }
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) sum += acc[s];
    const size_t row = (size_t)b * N + i0 + p;
    out[row * C + c] = ld_p(Pb + (size_t)(i0 + p) * LD + c) + __fdiv_rn(sum, (float)S);
    if (argmax) {
#pragma unroll
      for (int s = 0; s < S; ++s) argmax[row * SC + s * C + c] = (uint8_t)am[s];
    }
  }
}



// ---------------------------------------------------------------------------
// K4 forward, v2: two adjacent channels per thread on the packed-FP32 datapath.
//
// The v1 kernel above is issue-bound (ncu: 78 % issue-active, 6 % DRAM): ~16
// instructions per (neighbour, support, channel) element, half of them on the
// ALU pipe.  Here a thread owns the channel PAIR (c, c+1) of one point:
//   * theta for both channels is 3 packed instructions (FMUL2 + 2 FFMA2, sm_100
//     `__ffma2_rn`) on pre-halved unit directions d/2, and ReLU is the exact
//     identity  relu(t) = t/2 + |t/2|  (one FADD with a free |.| modifier) —
//     FMA pipe instead of an ALU-pipe FMNMX; bit-identical to max(t, 0);
//   * the activation product is one FMUL2;
//   * the support rows are read as one 4-byte (bf16x2) or 8-byte (fp32x2) load
//     per (neighbour, support): a warp still reads whole 128/256-byte segments;
//   * unit direction + neighbour index of a (point, neighbour) pair arrive as ONE
//     16-byte shared-memory broadcast.
// CTA = 128 threads = (C/2 channel pairs) x (points in parallel), tile of 8 points.
// ---------------------------------------------------------------------------
constexpr int GC2_THREADS = 128;

template <typename TP> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 load(const float* p) {
    return __ldg(reinterpret_cast<const float2*>(p));
  }
};
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 load(const __nv_bfloat16* p) {
    const unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
    // low half via an integer multiply (FMA pipe), high half via one LOP3 (ALU pipe)
    return make_float2(__uint_as_float(u * 65536u), __uint_as_float(u & 0xffff0000u));
  }
};

// AM: 0 = max only (no gradient needed), 1 = exact argmax (3 ALU ops / element: FSETP, FSEL,
// SEL), 2 = tagged argmax (2 ALU ops: the neighbour slot n replaces the 6 low mantissa bits of
// the candidate, one LOP3 + one FMNMX; value error <= 2^-17 relative — used with bf16 P, whose
// own rounding is 2^-9).  CT: compile-time channel count (0 = runtime C) so that the S support
// loads of a neighbour row are immediate offsets from one base address.
// UNR / MINB: measured on B200 (tools/kvar.py): 4 neighbours in flight per thread at 4 CTAs/SM
// (126 registers) beats both deeper unrolling and higher occupancy — the gather is latency-bound.
template <int S, typename TP, int AM, int CT, int UNR = 4, int MINB = 4>
__global__ void __launch_bounds__(GC2_THREADS, MINB)
graph_conv_fwd2_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                       const float* __restrict__ dirn, const TP* __restrict__ P, int N, int k,
                       int Crt, int lanes_c, float* __restrict__ out, uint8_t* __restrict__ argmax) {
  __shared__ float4 s_rn[GC_PT * GC_MAXK];   // (rhat.x, rhat.y, rhat.z, bits(neighbour index))
  const int C = CT ? CT : Crt;
  const int b = blockIdx.y, i0 = blockIdx.x * GC_PT;
  const int npts = min(GC_PT, N - i0);
  const int SC = S * C, LD = (S + 1) * C;
  const float* xb = xyz + (size_t)b * N * 3;
  const int32_t* ib = idx + (size_t)b * N * k;
  for (int p = threadIdx.x; p < npts * k; p += GC2_THREADS) {
    const int i = i0 + p / k;
    const int nb = ib[(size_t)i * k + (p % k)];
    float r[3];
    unit_dir(xb, i, nb, r);
    s_rn[p] = make_float4(r[0], r[1], r[2], __int_as_float(nb));
  }
  const int pts_par = GC2_THREADS / lanes_c;
  const int lane_c = threadIdx.x % lanes_c, psub = threadIdx.x / lanes_c;
  const int c = (blockIdx.z * lanes_c + lane_c) * 2;
  const bool active = psub < pts_par && c < C;
  float2 dx[S], dy[S], dz[S];
  if (active) {
#pragma unroll
    for (int s = 0; s < S; ++s) {   // halved: theta/2 = rhat . (d/2), exact
      const float2 x2 = *reinterpret_cast<const float2*>(dirn + s * C + c);
      const float2 y2 = *reinterpret_cast<const float2*>(dirn + SC + s * C + c);
      const float2 z2 = *reinterpret_cast<const float2*>(dirn + 2 * SC + s * C + c);
      dx[s] = make_float2(0.5f * x2.x, 0.5f * x2.y);
      dy[s] = make_float2(0.5f * y2.x, 0.5f * y2.y);
      dz[s] = make_float2(0.5f * z2.x, 0.5f * z2.y);
    }
  }
  __syncthreads();
  if (!active) return;
  const TP* Pb = P + (size_t)b * N * LD + C + c;   // support block, this thread's channel pair
  for (int p = psub; p < npts; p += pts_par) {
    float2 acc[S];
    int am0[S], am1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s] = make_float2(-INFINITY, -INFINITY); am0[s] = 0; am1[s] = 0; }
#pragma unroll UNR
    for (int n = 0; n < k; ++n) {
      const float4 rn = s_rn[p * k + n];
      const TP* sup = Pb + (size_t)__float_as_int(rn.w) * LD;
      float2 v[S];
#pragma unroll
      for (int s = 0; s < S; ++s) v[s] = Pair<TP>::load(sup + s * C);
      const float2 rx = make_float2(rn.x, rn.x), ry = make_float2(rn.y, rn.y),
                   rz = make_float2(rn.z, rn.z);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float2 t = __ffma2_rn(rz, dz[s], __ffma2_rn(ry, dy[s], __fmul2_rn(rx, dx[s])));
        t.x = __fadd_rn(t.x, fabsf(t.x));      // relu(theta), exact
        t.y = __fadd_rn(t.y, fabsf(t.y));
        const float2 a = __fmul2_rn(t, v[s]);
        if (AM == 1) {
          if (a.x > acc[s].x) { acc[s].x = a.x; am0[s] = n; }
          if (a.y > acc[s].y) { acc[s].y = a.y; am1[s] = n; }
        } else if (AM == 2) {
          acc[s].x = fmaxf(acc[s].x, __uint_as_float((__float_as_uint(a.x) & 0xffffffc0u) | (unsigned)n));
          acc[s].y = fmaxf(acc[s].y, __uint_as_float((__float_as_uint(a.y) & 0xffffffc0u) | (unsigned)n));
        } else {
          acc[s].x = fmaxf(acc[s].x, a.x);
          acc[s].y = fmaxf(acc[s].y, a.y);
        }
      }
    }
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (AM == 2) {
        am0[s] = (int)(__float_as_uint(acc[s].x) & 63u);
        am1[s] = (int)(__float_as_uint(acc[s].y) & 63u);
        acc[s].x = __uint_as_float(__float_as_uint(acc[s].x) & 0xffffffc0u);
        acc[s].y = __uint_as_float(__float_as_uint(acc[s].y) & 0xffffffc0u);
      }
      sx += acc[s].x; sy += acc[s].y;
    }
    const size_t row = (size_t)b * N + i0 + p;
    const float2 ctr = Pair<TP>::load(Pb + (size_t)(i0 + p) * LD - C);
    *reinterpret_cast<float2*>(out + row * C + c) =
        make_float2(ctr.x + __fdiv_rn(sx, (float)S), ctr.y + __fdiv_rn(sy, (float)S));
    if (AM != 0) {
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<uchar2*>(argmax + row * SC + s * C + c) =
            make_uchar2((unsigned char)am0[s], (unsigned char)am1[s]);
    }
  }
}

// K3 forward, v2: channel pairs on the packed-FP32 datapath (see graph_conv_fwd2_kernel).
// theta for two channels = FMUL2 + 2 FFMA2; max/argmax per element on the ALU pipe.
template <int S, bool AM>
__global__ void __launch_bounds__(GC2_THREADS, 4)
surface_conv_fwd2_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                         const float* __restrict__ dirn, int N, int k, int C, int lanes_c,
                         float* __restrict__ out, uint8_t* __restrict__ argmax) {
  __shared__ float4 s_rn[GC_PT * GC_MAXK];
  const int b = blockIdx.y, i0 = blockIdx.x * GC_PT;
  const int npts = min(GC_PT, N - i0);
  const int SC = S * C;
  const float* xb = xyz + (size_t)b * N * 3;
  const int32_t* ib = idx + (size_t)b * N * k;
  for (int p = threadIdx.x; p < npts * k; p += GC2_THREADS) {
    const int i = i0 + p / k;
    const int nb = ib[(size_t)i * k + (p % k)];
    float r[3];
    unit_dir(xb, i, nb, r);
    s_rn[p] = make_float4(r[0], r[1], r[2], 0.0f);
  }
  const int pts_par = GC2_THREADS / lanes_c;
  const int lane_c = threadIdx.x % lanes_c, psub = threadIdx.x / lanes_c;
  const int c = (blockIdx.z * lanes_c + lane_c) * 2;
  const bool active = psub < pts_par && c < C;
  float2 dx[S], dy[S], dz[S];
  if (active) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      dx[s] = *reinterpret_cast<const float2*>(dirn + s * C + c);
      dy[s] = *reinterpret_cast<const float2*>(dirn + SC + s * C + c);
      dz[s] = *reinterpret_cast<const float2*>(dirn + 2 * SC + s * C + c);
    }
  }
  __syncthreads();
  if (!active) return;
  for (int p = psub; p < npts; p += pts_par) {
    float2 acc[S];
    int am0[S], am1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s] = make_float2(0.0f, 0.0f); am0[s] = 255; am1[s] = 255; }
#pragma unroll 4
    for (int n = 0; n < k; ++n) {
      const float4 rn = s_rn[p * k + n];
      const float2 rx = make_float2(rn.x, rn.x), ry = make_float2(rn.y, rn.y), rz = make_float2(rn.z, rn.z);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float2 t = __ffma2_rn(rz, dz[s], __ffma2_rn(ry, dy[s], __fmul2_rn(rx, dx[s])));
        if (AM) {
          if (t.x > acc[s].x) { acc[s].x = t.x; am0[s] = n; }
          if (t.y > acc[s].y) { acc[s].y = t.y; am1[s] = n; }
        } else {
          acc[s].x = fmaxf(acc[s].x, t.x);
          acc[s].y = fmaxf(acc[s].y, t.y);
        }
      }
    }
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) { sx += acc[s].x; sy += acc[s].y; }
    const size_t row = (size_t)b * N + i0 + p;
    *reinterpret_cast<float2*>(out + row * C + c) =
        make_float2(__fdiv_rn(sx, (float)S), __fdiv_rn(sy, (float)S));
    if (AM) {
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<uchar2*>(argmax + row * SC + s * C + c) =
            make_uchar2((unsigned char)am0[s], (unsigned char)am1[s]);
    }
  }
}

template <int S, typename TP, int AM>
static void launch_gc2(dim3 grid2, cudaStream_t st, const float* xyz, const int32_t* idx,
                       const float* dirn, const TP* P, int N, int k, int C, int lanes_c, float* out,
                       uint8_t* argmax) {
  switch (C) {
    case 128: graph_conv_fwd2_kernel<S, TP, AM, 128><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, P, N, k, C, lanes_c, out, argmax); break;
    case 256: graph_conv_fwd2_kernel<S, TP, AM, 256><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, P, N, k, C, lanes_c, out, argmax); break;
    case 512: graph_conv_fwd2_kernel<S, TP, AM, 512><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, P, N, k, C, lanes_c, out, argmax); break;
    default:  graph_conv_fwd2_kernel<S, TP, AM, 0><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, P, N, k, C, lanes_c, out, argmax); break;
  }
}

// get_neighbor_direction_norm (gcn3d.py:49-59) as a stand-alone op (API parity;
// the fused kernels above never materialise it).
__global__ void direction_norm_kernel(const float* __restrict__ xyz,
                                      const int32_t* __restrict__ idx, int N, int k, int total,
                                      float* __restrict__ out_norm, float* __restrict__ out_raw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int bi = t / k, b = bi / N, i = bi % N;
  const float* xb = xyz + (size_t)b * N * 3;
  const int j = idx[t];
  float r[3];
  unit_dir(xb, i, j, r);
  out_norm[3 * (size_t)t] = r[0];
  out_norm[3 * (size_t)t + 1] = r[1];
  out_norm[3 * (size_t)t + 2] = r[2];
  if (out_raw) {
    out_raw[3 * (size_t)t] = __fsub_rn(xb[3 * j], xb[3 * i]);
    out_raw[3 * (size_t)t + 1] = __fsub_rn(xb[3 * j + 1], xb[3 * i + 1]);
    out_raw[3 * (size_t)t + 2] = __fsub_rn(xb[3 * j + 2], xb[3 * i + 2]);
  }
}

// ---------------------------------------------------------------- backward
// Persistent CTAs walk (object, tile) work items; every thread keeps the
// gradient of its channel's S support directions in registers, so the only
// cross-CTA reduction is one fixed-order pass over gridDim.x partial rows
// (deterministic, no float atomics on gdirn).
constexpr int GC_BWD_CTAS = 148 * 4;

template <int S>
__global__ void __launch_bounds__(GC_THREADS)
surface_conv_bwd_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                        const uint8_t* __restrict__ argmax, const float* __restrict__ gout, int B,
                        int N, int k, int C, float* __restrict__ partial) {
  __shared__ int s_idx[GC_PT * GC_MAXK];
  __shared__ float s_r[GC_PT * GC_MAXK * 3];
  const int c = blockIdx.z * GC_THREADS + threadIdx.x;
  const int SC = S * C;
  const int tiles = (N + GC_PT - 1) / GC_PT;
  float gx[S], gy[S], gz[S];
#pragma unroll
  for (int s = 0; s < S; ++s) gx[s] = gy[s] = gz[s] = 0.0f;
  for (int w = blockIdx.x; w < B * tiles; w += gridDim.x) {
    const int b = w / tiles, i0 = (w % tiles) * GC_PT;
    const int npts = min(GC_PT, N - i0);
    __syncthreads();
    stage_tile(xyz + (size_t)b * N * 3, idx + (size_t)b * N * k, i0, npts, k, s_idx, s_r);
    __syncthreads();
    if (c < C) {
      // all winners and gradients of the tile are fetched first (8 x (S + 1) independent loads in
      // flight per thread): the kernel is a latency-bound stream of 1-byte gathers otherwise
      unsigned char am[GC_PT][S];
      float gv[GC_PT];
#pragma unroll
      for (int p = 0; p < GC_PT; ++p) {
        const size_t row = (size_t)b * N + i0 + (p < npts ? p : 0);
        gv[p] = gout[row * C + c];
#pragma unroll
        for (int s = 0; s < S; ++s) am[p][s] = argmax[row * SC + s * C + c];
      }
#pragma unroll
      for (int p = 0; p < GC_PT; ++p) {
        if (p < npts) {
          const float gs = __fdiv_rn(gv[p], (float)S);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int n = am[p][s];                       // 255: relu killed every neighbour
            if (n < k) {
              const float* r = s_r + 3 * (p * k + n);
              gx[s] = fmaf(gs, r[0], gx[s]);
              gy[s] = fmaf(gs, r[1], gy[s]);
              gz[s] = fmaf(gs, r[2], gz[s]);
            }
          }
        }
      }
    }
  }
  if (c < C) {
    float* pr = partial + (size_t)blockIdx.x * 3 * SC;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pr[s * C + c] = gx[s];
      pr[SC + s * C + c] = gy[s];
      pr[2 * SC + s * C + c] = gz[s];
    }
  }
}

// Fixed-order column sums of the per-CTA partial rows: block (32 columns, 8 row groups); group g
// adds rows g, g+8, ... (independent loads), the 8 group sums are then added in order 0..7.
__global__ void __launch_bounds__(256)
dir_reduce_kernel(const float* __restrict__ partial, int rows, int cols, float* __restrict__ out,
                  int cols_a, float* __restrict__ out_b) {
  __shared__ float sh[8][32];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float s = 0.0f;
  if (j < cols) {
#pragma unroll 4
    for (int r = threadIdx.y; r < rows; r += 8) s += __ldg(partial + (size_t)r * cols + j);
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    float t = 0.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += sh[g][threadIdx.x];
    if (j < cols_a) out[j] = t;
    else if (out_b) out_b[j - cols_a] = t;
  }
}

// K4b, atomic variant (any S, C, N).  A point costs two dependent L2 round trips (winner byte -> support
// value / RED); a loop over the tile's points that serialises them is latency-bound (round-1 kernel, ncu r1h:
// issue 17 %, DRAM 19 %; running it per L2-sized chunk of objects changed nothing, tools/k4b_chunk.py).
// Here the tile's winners and output gradients are all requested first (8 x (S + 1) independent loads per
// thread), then the support values of HALF a tile (4 x S loads in flight), then the REDs — about five
// round trips per tile instead of ~18.  (point, neighbour) unit direction + index come as one 16-byte
// shared-memory broadcast as in the forward kernel.
template <int S, typename TP, int MINB>
__global__ void __launch_bounds__(GC_THREADS, MINB)
graph_conv_bwd2_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                       const float* __restrict__ dirn, const TP* __restrict__ P,
                       const uint8_t* __restrict__ argmax, const float* __restrict__ gout, int B,
                       int N, int k, int C, float* __restrict__ gP, float* __restrict__ partial,
                       int want_gbias) {
  __shared__ float4 s_rn[GC_PT * GC_MAXK];   // (rhat.x, rhat.y, rhat.z, bits(neighbour index))
  const int c = blockIdx.z * GC_THREADS + threadIdx.x;
  const bool live = c < C;
  const int cc = live ? c : 0;
  const int SC = S * C, LD = (S + 1) * C;
  const int tiles = (N + GC_PT - 1) / GC_PT;
  constexpr int HALF = GC_PT / 2;
  // this thread's S support directions live in shared memory (conflict-free column reads): the registers
  // go to loads in flight
  __shared__ float s_d[3 * S][GC_THREADS];
  float gx[S], gy[S], gz[S], gb[S], gbc = 0.0f;   // gb*: column sums of gP
#pragma unroll
  for (int s = 0; s < S; ++s) {
    gx[s] = gy[s] = gz[s] = gb[s] = 0.0f;
    s_d[s][threadIdx.x] = dirn[s * C + cc];
    s_d[S + s][threadIdx.x] = dirn[SC + s * C + cc];
    s_d[2 * S + s][threadIdx.x] = dirn[2 * SC + s * C + cc];
  }
  for (int w = blockIdx.x; w < B * tiles; w += gridDim.x) {
    const int b = w / tiles, i0 = (w % tiles) * GC_PT;
    const int npts = min(GC_PT, N - i0);
    const uint8_t* amb = argmax + (size_t)b * N * SC + cc;
    const float* gob = gout + (size_t)b * N * C + cc;
    // winners + output gradients of the first half tile: in flight while the tile is staged
    unsigned byte[HALF][S];
    float gv[HALF];
    auto request = [&](int h) {
#pragma unroll
      for (int q = 0; q < HALF; ++q) {
        const int row = i0 + (h + q < npts ? h + q : 0);
        gv[q] = __ldg(gob + row * C);
#pragma unroll
        for (int s = 0; s < S; ++s) byte[q][s] = __ldg(amb + row * SC + s * C);
      }
    };
    request(0);
    __syncthreads();                       // previous tile's readers of s_rn are done
    {
      const float* xb = xyz + (size_t)b * N * 3;
      const int32_t* ib = idx + (size_t)b * N * k;
      for (int p = threadIdx.x; p < npts * k; p += GC_THREADS) {
        const int i = i0 + p / k;
        const int nb = ib[(size_t)i * k + (p % k)];
        float r[3];
        unit_dir(xb, i, nb, r);
        s_rn[p] = make_float4(r[0], r[1], r[2], __int_as_float(nb));
      }
    }
    __syncthreads();
    if (!live) continue;
    const TP* Pb = P + (size_t)b * N * LD + C + c;
    float* gPb = gP + (size_t)b * N * LD + c;
#pragma unroll
    for (int h = 0; h < GC_PT; h += HALF) {
      float pv[HALF][S], g[HALF];
      unsigned amw[HALF][2];               // winner slots packed: supports 0-3 | 4-7
#pragma unroll
      for (int q = 0; q < HALF; ++q) {
        g[q] = gv[q];
        amw[q][0] = 0u; amw[q][1] = 0u;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int n = min((int)byte[q][s], k - 1);            // (clamp never taken for a forward-written byte)
          amw[q][s >> 2] |= (unsigned)n << (8 * (s & 3));
          const int nb = __float_as_int(s_rn[(h + q < npts ? h + q : 0) * k + n].w);   // rows past the tile: point 0's
          pv[q][s] = ld_p(Pb + nb * LD + s * C);
        }
      }
      if (h + HALF < GC_PT) request(h + HALF);                  // next half's winners: in flight under the REDs
#pragma unroll
      for (int q = 0; q < HALF; ++q) {
        if (h + q < npts) {
          gPb[(i0 + h + q) * LD] = g[q];                      // centre term
          gbc += g[q];
          const float gs = __fdiv_rn(g[q], (float)S);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float4 rn = s_rn[(h + q) * k + ((amw[q][s >> 2] >> (8 * (s & 3))) & 0xffu)];
            const float th = fmaxf(fmaf(rn.z, s_d[2 * S + s][threadIdx.x],
                                        fmaf(rn.y, s_d[S + s][threadIdx.x], rn.x * s_d[s][threadIdx.x])), 0.0f);
            if (th > 0.0f) {
              const float contrib = gs * th;
              atomicAdd(gPb + __float_as_int(rn.w) * LD + C + s * C, contrib);
              gb[s] += contrib;
              const float gvp = gs * pv[q][s];
              gx[s] = fmaf(gvp, rn.x, gx[s]);
              gy[s] = fmaf(gvp, rn.y, gy[s]);
              gz[s] = fmaf(gvp, rn.z, gz[s]);
            }
          }
        }
      }
    }
  }
  if (live) {
    const int pcols = want_gbias ? 3 * SC + (S + 1) * C : 3 * SC;
    float* pr = partial + (size_t)blockIdx.x * pcols;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pr[s * C + c] = gx[s];
      pr[SC + s * C + c] = gy[s];
      pr[2 * SC + s * C + c] = gz[s];
    }
    if (want_gbias) {       // bias layout of HS_layer: [centre C | s*C + c]
      pr[3 * SC + c] = gbc;
#pragma unroll
      for (int s = 0; s < S; ++s) pr[3 * SC + C + s * C + c] = gb[s];
    }
  }
}

// K4b v3 ("object-resident").  ncu of v2 on the 128-channel layer: the L1 -> crossbar request path is 83 %
// busy — every (point, support, channel) winner is its own 4-byte RED / 2-byte load to one of ~15 different
// rows per warp instruction (57 M RED requests, 102 M sectors for 118 M elements), and gP costs a 539 MB
// memset + a 539 MB -> 269 MB cast pass on top.  Here one CTA owns the gradient slab of (object b, support s,
// 32 channels): N x 32 floats in shared memory (131.6 KB at N = 1028), lane = channel, so the scattered adds
// are shared-memory atomics with bank = lane (conflict-free across rows), and the slab leaves the SM ONCE, as
// whole 64-byte (bf16) / 128-byte (fp32) row segments in the caller's dtype — no memset, no global RED, no
// cast pass.  Item s == S writes the centre columns.  Each warp walks batches of 8 points: winners and
// output gradients first, then the 8 support values (independent loads in flight), then the adds.
// gdirn / gbias: per-thread sums over the warp's points, fixed-order sum over the 8 warps, one partial row
// per object, fixed-order column reduction (dir_reduce_kernel) — deterministic as before.

__device__ __forceinline__ void st_g(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_g(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int S, typename TP, typename TO, int THREADS, int GO_PB>
__global__ void __launch_bounds__(THREADS)
graph_conv_bwd_obj_kernel(const float4* __restrict__ rnbuf, const float* __restrict__ dirn,
                          const TP* __restrict__ P, const uint8_t* __restrict__ argmax,
                          const float* __restrict__ gout, const float* __restrict__ gmax, int N, int k, int C,
                          TO* __restrict__ gP, float* __restrict__ partial, int want_gbias) {
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) unsigned char go_smem[];
  float* acc = reinterpret_cast<float*>(go_smem);                          // [N][32]
  float* s_red = acc + (size_t)N * 32;                                     // [4][WARPS][32]
  const int b = blockIdx.z, s = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int SC = S * C, LD = (S + 1) * C;
  const int pcols = want_gbias ? 3 * SC + (S + 1) * C : 3 * SC;
  const float* gob = gout + (size_t)b * N * C + c;
  TO* gPb = gP + (size_t)b * N * LD;
  float* prow = partial + (size_t)b * pcols;
  if (s == S) {                              // centre term: gP[:, :C] = gout, gbias[:C] = its column sums
    float gbc = 0.0f;
#pragma unroll 8
    for (int j = warp; j < N; j += WARPS) {
      const float g = __ldg(gob + (size_t)j * C);
      st_g(gPb + (size_t)j * LD + c, g);
      gbc += g;
    }
    if (want_gbias) {
      s_red[warp * 32 + lane] = gbc;
      __syncthreads();
      if (warp == 0) {
        float t = 0.0f;
#pragma unroll 8
        for (int w = 0; w < WARPS; ++w) t += s_red[w * 32 + lane];
        prow[3 * SC + c] = t;
      }
    }
    return;
  }
  for (int e = threadIdx.x; e < N * 32; e += THREADS) acc[e] = 0.0f;    // (+0.0f and integer 0 share the bit pattern)
  // bf16 output: the slab accumulates in 32-bit FIXED POINT — shared-memory integer adds are one native
  // instruction (a float add is a compare-and-swap loop) and associative, so gP is bit-reproducible.  Scale:
  // |contribution| <= max|gout| / S over the CTA's rows x channels, at most N contributions per accumulator,
  // so 1.5e9 / (N * max|gout| / S) cannot overflow; the quantum is <= 2^-20 of that maximum (bf16 rounds at 2^-9).
  constexpr bool FIXED = sizeof(TO) == 2;
  float scale = 1.0f;
  if (FIXED) {                               // max|gout| over the CTA's rows x channels (absmax_slab_kernel)
    const float m = __ldg(gmax + (size_t)b * gridDim.x + blockIdx.x);
    scale = m > 0.0f ? 1.5e9f * (float)S / ((float)N * m) : 1.0f;
  }
  const float inv_scale = __fdiv_rn(1.0f, scale);
  const float dx = dirn[s * C + c], dy = dirn[SC + s * C + c], dz = dirn[2 * SC + s * C + c];
  const uint8_t* amb = argmax + (size_t)b * N * SC + s * C + c;
  const TP* Pb = P + (size_t)b * N * LD + C + s * C + c;
  const float4* rnb = rnbuf + (size_t)b * N * k;        // (rhat, neighbour index) of every (point, slot) pair
  float gx = 0.0f, gy = 0.0f, gz = 0.0f, gb = 0.0f;
  unsigned win[GO_PB];
  float g8[GO_PB];
  auto request = [&](int base) {             // winners + output gradients of a batch (rows clamped into the object)
#pragma unroll
    for (int p = 0; p < GO_PB; ++p) {
      const int row = min(base + p, N - 1);
      win[p] = __ldg(amb + (size_t)row * SC);
      g8[p] = __ldg(gob + (size_t)row * C);
    }
  };
  request(warp * GO_PB);
  __syncthreads();
  for (int base = warp * GO_PB; base < N; base += WARPS * GO_PB) {
    const int npts = min(GO_PB, N - base);
    float4 rn[GO_PB];
    float pv[GO_PB], gs[GO_PB];
#pragma unroll
    for (int p = 0; p < GO_PB; ++p) {        // a warp's lanes read inside one point's k x 16-byte row
      const int n = min((int)win[p], k - 1);                     // (clamp never taken for a forward-written byte)
      rn[p] = __ldg(rnb + (size_t)min(base + p, N - 1) * k + n);
      gs[p] = __fdiv_rn(g8[p], (float)S);
    }
    if (base + WARPS * GO_PB < N) request(base + WARPS * GO_PB);   // next batch: in flight under this one
#pragma unroll
    for (int p = 0; p < GO_PB; ++p) pv[p] = ld_p(Pb + (size_t)__float_as_int(rn[p].w) * LD);
#pragma unroll
    for (int p = 0; p < GO_PB; ++p) {
      if (p < npts) {
        const float th = fmaxf(fmaf(rn[p].z, dz, fmaf(rn[p].y, dy, rn[p].x * dx)), 0.0f);
        if (th > 0.0f) {
          const float contrib = gs[p] * th;
          if (FIXED) atomicAdd(reinterpret_cast<int*>(acc) + __float_as_int(rn[p].w) * 32 + lane, __float2int_rn(contrib * scale));
          else atomicAdd(acc + __float_as_int(rn[p].w) * 32 + lane, contrib);
          gb += contrib;
          const float gvp = gs[p] * pv[p];
          gx = fmaf(gvp, rn[p].x, gx);
          gy = fmaf(gvp, rn[p].y, gy);
          gz = fmaf(gvp, rn[p].z, gz);
        }
      }
    }
  }
  __syncthreads();                           // slab complete
  s_red[(0 * WARPS + warp) * 32 + lane] = gx;
  s_red[(1 * WARPS + warp) * 32 + lane] = gy;
  s_red[(2 * WARPS + warp) * 32 + lane] = gz;
  s_red[(3 * WARPS + warp) * 32 + lane] = gb;
  // the slab leaves as whole row segments: two rows per warp instruction (half-warp = 16 channel pairs)
  {
    const int hl = lane & 15, hr = lane >> 4;
    for (int j = warp * 2 + hr; j < N; j += WARPS * 2) {
      const float2 v = *reinterpret_cast<const float2*>(acc + j * 32 + 2 * hl);
      TO* dst = gPb + (size_t)j * LD + C + s * C + blockIdx.x * 32 + 2 * hl;
      if (FIXED) {
        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(
            (float)__float_as_int(v.x) * inv_scale, (float)__float_as_int(v.y) * inv_scale);
      } else {
        *reinterpret_cast<float2*>(dst) = v;
      }
    }
  }
  __syncthreads();
  if (warp < 4 && (warp < 3 || want_gbias)) {
    float t = 0.0f;
#pragma unroll 8
    for (int w = 0; w < WARPS; ++w) t += s_red[(warp * WARPS + w) * 32 + lane];
    if (warp < 3) prow[warp * SC + s * C + c] = t;
    else prow[3 * SC + C + s * C + c] = t;
  }
}

// max |gout[b, :, 32 cg .. 32 cg + 31]| for the fixed-point scale of the bf16 slabs: grid (C / 32, B)
__global__ void __launch_bounds__(256)
absmax_slab_kernel(const float* __restrict__ gout, int N, int C, float* __restrict__ out) {
  __shared__ float sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* g = gout + (size_t)blockIdx.y * N * C + blockIdx.x * 32 + lane;
  float m = 0.0f;
#pragma unroll 8
  for (int j = warp; j < N; j += 8) m = fmaxf(m, fabsf(__ldg(g + (size_t)j * C)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) sh[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
    out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = m;
  }
}

// (unit direction, neighbour index) of every (point, neighbour slot) pair, once per layer call
__global__ void __launch_bounds__(256)
pair_dirs_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx, int N, int k, int total,
                 float4* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int bi = t / k, b = bi / N, i = bi % N;
  const int j = idx[t];
  float r[3];
  unit_dir(xyz + (size_t)b * N * 3, i, j, r);
  out[t] = make_float4(r[0], r[1], r[2], __int_as_float(j));
}

static int bwd_obj_threads(int N) { return N >= 512 ? 1024 : 256; }
static size_t bwd_obj_smem(int N, int k) {
  (void)k;
  return (size_t)N * 32 * sizeof(float) + (size_t)4 * (bwd_obj_threads(N) / 32) * 32 * sizeof(float);
}

static int bwd_ctas(int B, int N) {
  long items = (long)B * ((N + GC_PT - 1) / GC_PT);
  return (int)(items < GC_BWD_CTAS ? items : GC_BWD_CTAS);
}

#define HSP_DISPATCH_S(S_, CALL)                 \
  switch (S_) {                                  \
    case 1: { constexpr int S = 1; CALL; } break; \
    case 2: { constexpr int S = 2; CALL; } break; \
    case 3: { constexpr int S = 3; CALL; } break; \
    case 4: { constexpr int S = 4; CALL; } break; \
    case 5: { constexpr int S = 5; CALL; } break; \
    case 6: { constexpr int S = 6; CALL; } break; \
    case 7: { constexpr int S = 7; CALL; } break; \
    case 8: { constexpr int S = 8; CALL; } break; \
    default: return HSP_EINVAL;                  \
  }

static bool bad_dims(int B, int N, int k, int S, int C) {
  return B < 0 || N <= 0 || k <= 0 || k > GC_MAXK || S < 1 || S > 8 || C <= 0 || B > 65535;
}

}  // namespace hsp

extern "C" int hsp_surface_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn,
                                    int B, int N, int k, int S, int C, float* out,
                                    uint8_t* argmax, void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !dirn || !out || bad_dims(B, N, k, S, C)) return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  dim3 grid((N + GC_PT - 1) / GC_PT, B, (C + GC_THREADS - 1) / GC_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 1) == 0 && S == 7 && (((uintptr_t)argmax) & 1) == 0) {   // v2: channel pairs, packed FP32
    const int pairs = C / 2;
    const int lanes_c = pairs < GC2_THREADS ? pairs : GC2_THREADS;
    dim3 grid2((N + GC_PT - 1) / GC_PT, B, (pairs + lanes_c - 1) / lanes_c);
    if (argmax)
      surface_conv_fwd2_kernel<7, true><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, N, k, C, lanes_c, out, argmax);
    else
      surface_conv_fwd2_kernel<7, false><<<grid2, GC2_THREADS, 0, st>>>(xyz, idx, dirn, N, k, C, lanes_c, out, nullptr);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  HSP_DISPATCH_S(S, (surface_conv_fwd_kernel<S><<<grid, GC_THREADS, 0, st>>>(xyz, idx, dirn, N, k, C, out, argmax)));
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_graph_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn,
                                  const void* P, int p_dtype, int B, int N, int k, int S, int C,
                                  float* out, uint8_t* argmax, void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !dirn || !P || !out || bad_dims(B, N, k, S, C)) return HSP_EINVAL;
  if (argmax && k > 255) return HSP_EINVAL;
  if (p_dtype != HSP_DTYPE_F32 && p_dtype != HSP_DTYPE_BF16) return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  dim3 grid((N + GC_PT - 1) / GC_PT, B, (C + GC_THREADS - 1) / GC_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 1) == 0 && S == 7) {   // v2 (reference default S = 7): channel pairs, packed FP32
    const int pairs = C / 2;
    const int lanes_c = pairs < GC2_THREADS ? pairs : GC2_THREADS;
    dim3 grid2((N + GC_PT - 1) / GC_PT, B, (pairs + lanes_c - 1) / lanes_c);
#define HSP_GC2(TP_, AM_) \
    launch_gc2<7, TP_, AM_>(grid2, st, xyz, idx, dirn, (const TP_*)P, N, k, C, lanes_c, out, argmax)
    if (p_dtype == HSP_DTYPE_BF16) {
      if (argmax) { HSP_GC2(__nv_bfloat16, 2); } else { HSP_GC2(__nv_bfloat16, 0); }
    } else {
      if (argmax) { HSP_GC2(float, 1); } else { HSP_GC2(float, 0); }
    }
#undef HSP_GC2
  } else if (p_dtype == HSP_DTYPE_BF16) {
    HSP_DISPATCH_S(S, (graph_conv_fwd_kernel<S, __nv_bfloat16><<<grid, GC_THREADS, 0, st>>>(
                          xyz, idx, dirn, (const __nv_bfloat16*)P, N, k, C, out, argmax)));
  } else {
    HSP_DISPATCH_S(S, (graph_conv_fwd_kernel<S, float><<<grid, GC_THREADS, 0, st>>>(
                          xyz, idx, dirn, (const float*)P, N, k, C, out, argmax)));
  }
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" size_t hsp_surface_conv_bwd_workspace_bytes(int B, int N, int k, int S, int C) {
  using namespace hsp;
  if (bad_dims(B, N, k, S, C) || B == 0) return 0;
  return (size_t)bwd_ctas(B, N) * 3 * S * C * sizeof(float);
}
extern "C" size_t hsp_graph_conv_bwd_workspace_bytes(int B, int N, int k, int S, int C) {
  using namespace hsp;
  if (bad_dims(B, N, k, S, C) || B == 0) return 0;
  return (size_t)bwd_ctas(B, N) * (3 * S + S + 1) * C * sizeof(float);
}

extern "C" int hsp_surface_conv_bwd(const float* xyz, const int32_t* idx, const uint8_t* argmax,
                                    const float* gout, int B, int N, int k, int S, int C,
                                    float* gdirn, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !argmax || !gout || !gdirn || bad_dims(B, N, k, S, C)) return HSP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    return cudaMemsetAsync(gdirn, 0, sizeof(float) * 3 * S * C, st) == cudaSuccess ? HSP_OK
                                                                                   : HSP_ELAUNCH;
  }
  if (!workspace || workspace_bytes < hsp_surface_conv_bwd_workspace_bytes(B, N, k, S, C))
    return HSP_EWORKSPACE;
  const int ctas = bwd_ctas(B, N);
  dim3 grid(ctas, 1, (C + GC_THREADS - 1) / GC_THREADS);
  HSP_DISPATCH_S(S, (surface_conv_bwd_kernel<S><<<grid, GC_THREADS, 0, st>>>(
                        xyz, idx, argmax, gout, B, N, k, C, (float*)workspace)));
  HSP_LAUNCH_CHECK();
  const int cols = 3 * S * C;
  dir_reduce_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, st>>>((const float*)workspace, ctas, cols, gdirn,
                                                          cols, nullptr);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_graph_conv_bwd(const float* xyz, const int32_t* idx, const float* dirn,
                                  const void* P, int p_dtype, const uint8_t* argmax,
                                  const float* gout, int B, int N, int k, int S, int C, float* gP,
                                  float* gdirn, float* gbias, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !dirn || !P || !argmax || !gout || !gP || !gdirn ||
      bad_dims(B, N, k, S, C))
    return HSP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    if (gbias && cudaMemsetAsync(gbias, 0, sizeof(float) * (S + 1) * C, st) != cudaSuccess)
      return HSP_ELAUNCH;
    return cudaMemsetAsync(gdirn, 0, sizeof(float) * 3 * S * C, st) == cudaSuccess ? HSP_OK
                                                                                   : HSP_ELAUNCH;
  }
  if (!workspace || workspace_bytes < hsp_graph_conv_bwd_workspace_bytes(B, N, k, S, C))
    return HSP_EWORKSPACE;
  const int want_gbias = gbias != nullptr;
  if (cudaMemsetAsync(gP, 0, sizeof(float) * (size_t)B * N * (S + 1) * C, st) != cudaSuccess)
    return HSP_ELAUNCH;
  const int ctas = bwd_ctas(B, N);
  dim3 grid(ctas, 1, (C + GC_THREADS - 1) / GC_THREADS);
  if (p_dtype == HSP_DTYPE_BF16) {
    HSP_DISPATCH_S(S, (graph_conv_bwd2_kernel<S, __nv_bfloat16, 4><<<grid, GC_THREADS, 0, st>>>(
                          xyz, idx, dirn, (const __nv_bfloat16*)P, argmax, gout, B, N, k, C, gP,
                          (float*)workspace, want_gbias)));
  } else if (p_dtype == HSP_DTYPE_F32) {
    HSP_DISPATCH_S(S, (graph_conv_bwd2_kernel<S, float, 4><<<grid, GC_THREADS, 0, st>>>(
                          xyz, idx, dirn, (const float*)P, argmax, gout, B, N, k, C, gP,
                          (float*)workspace, want_gbias)));
  } else {
    return HSP_EINVAL;
  }
  HSP_LAUNCH_CHECK();
  const int cols_a = 3 * S * C, cols = want_gbias ? cols_a + (S + 1) * C : cols_a;
  dir_reduce_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, st>>>((const float*)workspace, ctas, cols, gdirn,
                                                          cols_a, gbias);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

// Object-resident variant (graph_conv_bwd_obj_kernel): gP is written once, in `gp_dtype`, with no memset.
// Supported when C % 32 == 0 and the N x 32 fp32 slab + staging fit in shared memory (N <= ~1600).
extern "C" int hsp_graph_conv_bwd_obj_supported(int N, int k, int C) {
  using namespace hsp;
  return N > 0 && k > 0 && k <= GC_MAXK && C > 0 && (C % 32) == 0 && bwd_obj_smem(N, k) <= 220 * 1024 ? 1 : 0;
}
extern "C" size_t hsp_graph_conv_bwd_obj_workspace_bytes(int B, int N, int k, int S, int C) {
  using namespace hsp;
  if (bad_dims(B, N, k, S, C) || B == 0) return 0;
  return (size_t)B * (3 * S + S + 1) * C * sizeof(float) + (size_t)B * N * k * sizeof(float4) + 16 +
         (size_t)B * (C / 32 + 1) * sizeof(float);
}
extern "C" int hsp_graph_conv_bwd_obj(const float* xyz, const int32_t* idx, const float* dirn, const void* P,
                                      int p_dtype, const uint8_t* argmax, const float* gout, int B, int N,
                                      int k, int S, int C, void* gP, int gp_dtype, float* gdirn, float* gbias,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !dirn || !P || !argmax || !gout || !gP || !gdirn || bad_dims(B, N, k, S, C))
    return HSP_EINVAL;
  if (!hsp_graph_conv_bwd_obj_supported(N, k, C) || S != 7) return HSP_EINVAL;
  if ((p_dtype != HSP_DTYPE_F32 && p_dtype != HSP_DTYPE_BF16) || (gp_dtype != HSP_DTYPE_F32 && gp_dtype != HSP_DTYPE_BF16))
    return HSP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    if (gbias && cudaMemsetAsync(gbias, 0, sizeof(float) * (S + 1) * C, st) != cudaSuccess) return HSP_ELAUNCH;
    return cudaMemsetAsync(gdirn, 0, sizeof(float) * 3 * S * C, st) == cudaSuccess ? HSP_OK : HSP_ELAUNCH;
  }
  if (!workspace || workspace_bytes < hsp_graph_conv_bwd_obj_workspace_bytes(B, N, k, S, C)) return HSP_EWORKSPACE;
  const int want_gbias = gbias != nullptr;
  const size_t smem = bwd_obj_smem(N, k);
  dim3 grid(C / 32, S + 1, B);
  if ((long)B * N * k > 0x7fffffffL) return HSP_EINVAL;
  float* partial = (float*)workspace;
  float4* rnbuf = (float4*)(((uintptr_t)(partial + (size_t)B * (3 * S + S + 1) * C) + 15) & ~(uintptr_t)15);
  const int total = B * N * k;
  pair_dirs_kernel<<<(total + 255) / 256, 256, 0, st>>>(xyz, idx, N, k, total, rnbuf);
  HSP_LAUNCH_CHECK();
  float* gmax = (float*)(rnbuf + total);
  if (gp_dtype == HSP_DTYPE_BF16) {
    absmax_slab_kernel<<<dim3(C / 32, B), 256, 0, st>>>(gout, N, C, gmax);
    HSP_LAUNCH_CHECK();
  }
#define HSP_GCO_T(TP_, TO_, T_, PB_)                                                                             \
  do {                                                                                                           \
    auto kern = graph_conv_bwd_obj_kernel<7, TP_, TO_, T_, PB_>;                                                 \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)      \
      return HSP_ELAUNCH;                                                                                        \
    kern<<<grid, T_, smem, st>>>(rnbuf, dirn, (const TP_*)P, argmax, gout, gmax, N, k, C, (TO_*)gP, partial,     \
                                 want_gbias);                                                                    \
  } while (0)
#define HSP_GCO(TP_, TO_)                                                                                        \
  do {                                                                                                           \
    if (threads == 1024) HSP_GCO_T(TP_, TO_, 1024, 4);                                                           \
    else HSP_GCO_T(TP_, TO_, 256, 4);                                                                            \
  } while (0)
  const int threads = bwd_obj_threads(N);
  if (p_dtype == HSP_DTYPE_BF16) {
    if (gp_dtype == HSP_DTYPE_BF16) HSP_GCO(__nv_bfloat16, __nv_bfloat16); else HSP_GCO(__nv_bfloat16, float);
  } else {
    if (gp_dtype == HSP_DTYPE_BF16) HSP_GCO(float, __nv_bfloat16); else HSP_GCO(float, float);
  }
#undef HSP_GCO_T
#undef HSP_GCO
  HSP_LAUNCH_CHECK();
  const int cols_a = 3 * S * C, cols = want_gbias ? cols_a + (S + 1) * C : cols_a;
  dir_reduce_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, st>>>(partial, B, cols, gdirn, cols_a, gbias);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_neighbor_direction_norm(const float* xyz, const int32_t* idx, int B, int N,
                                           int k, float* out_norm, float* out_raw,
                                           void* stream) {
  using namespace hsp;
  if (!xyz || !idx || !out_norm || B < 0 || N <= 0 || k <= 0) return HSP_EINVAL;
  if ((long)B * N * k > 0x7fffffffL) return HSP_EINVAL;
  const int total = B * N * k;
  if (total == 0) return HSP_OK;
  direction_norm_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      xyz, idx, N, k, total, out_norm, out_raw);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
