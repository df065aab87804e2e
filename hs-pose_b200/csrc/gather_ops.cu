// K5 — row gathers fused with their reductions (forward + backward).
//   hsp_gather_max_*     Pool_layer.forward (reference gcn3d.py:234-246)
//   hsp_orl_global_*     get_ORL_global     (gcn3d.py:211-218)
//   hsp_upsample_rows_*  nearest up-sampling (FaceRecon.py:100-104)
// indexing_neighbor_new (gcn3d.py:39-47) never materialises (B,N,k,C): one
// thread owns one channel, a warp reads 128 contiguous bytes of each gathered
// row, the max over neighbours stays in a register.
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {

constexpr int GO_THREADS = 128;
constexpr int GO_PT = 8;  // rows per CTA

__global__ void __launch_bounds__(GO_THREADS)
gather_max_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                      const int32_t* __restrict__ rows, int N, int C, int R, int kuse,
                      int kstride, float* __restrict__ out, uint8_t* __restrict__ argmax) {
  const int b = blockIdx.y;
  const int c = blockIdx.z * GO_THREADS + threadIdx.x;
  if (c >= C) return;
  const float* fb = feat + (size_t)b * N * C;
  const int r_end = min((int)(blockIdx.x + 1) * GO_PT, R);
  for (int r = blockIdx.x * GO_PT; r < r_end; ++r) {
    const int i = rows ? __ldg(rows + r) : r;
    const int32_t* ip = idx + ((size_t)b * N + i) * kstride;
    float m = -INFINITY;
    int am = 0;
    for (int n = 0; n < kuse; ++n) {
      float v = __ldg(fb + (size_t)__ldg(ip + n) * C + c);
      if (v > m) { m = v; am = n; }
    }
    out[((size_t)b * R + r) * C + c] = m;
    if (argmax) argmax[((size_t)b * R + r) * C + c] = (uint8_t)am;
  }
}

__global__ void __launch_bounds__(GO_THREADS)
gather_max_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                      const int32_t* __restrict__ rows, const uint8_t* __restrict__ argmax,
                      int N, int C, int R, int kstride, float* __restrict__ gfeat) {
  const int b = blockIdx.y;
  const int c = blockIdx.z * GO_THREADS + threadIdx.x;
  if (c >= C) return;
  const int r_end = min((int)(blockIdx.x + 1) * GO_PT, R);
  for (int r = blockIdx.x * GO_PT; r < r_end; ++r) {
    const int i = rows ? __ldg(rows + r) : r;
    const size_t o = ((size_t)b * R + r) * C + c;
    const int src = __ldg(idx + ((size_t)b * N + i) * kstride + argmax[o]);
    atomicAdd(gfeat + ((size_t)b * N + src) * C + c, gout[o]);
  }
}

// Deterministic form of the same scatter (the default): warp = one SOURCE row j.  The object's (sampled row, slot)
// -> source table (R x kuse ints) sits in shared memory; the warp scans it 32 entries at a time (ballot) and adds
// the matching output gradients, where the saved arg-max names that slot, in ascending (r, slot) order.  Every
// gfeat row is written exactly once: no float atomics, no zero-fill, bit-reproducible.
constexpr int GMB_WARPS = 16;
constexpr int GMB_ROWS = 2;                    // source rows per warp: the table staging is amortised over 32 rows
constexpr int GMB_MAXV = 4;                    // float4 accumulators per lane: C <= 512
__global__ void __launch_bounds__(GMB_WARPS * 32)
gather_max_bwd_det_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx,
                          const int32_t* __restrict__ rows, const uint8_t* __restrict__ argmax, int N, int C,
                          int R, int kuse, int kstride, float* __restrict__ gfeat) {
  extern __shared__ __align__(16) int32_t s_tab[];   // [R * kuse], padded to a multiple of 4 with -1
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = R * kuse, T4 = (T + 3) >> 2;
  for (int t = threadIdx.x; t < 4 * T4; t += GMB_WARPS * 32) {
    int v = -1;
    if (t < T) {
      const int r = t / kuse, sl = t - r * kuse;
      const int i = rows ? __ldg(rows + r) : r;
      v = __ldg(idx + ((size_t)b * N + i) * kstride + sl);
    }
    s_tab[t] = v;
  }
  __syncthreads();
  const int4* tab4 = reinterpret_cast<const int4*>(s_tab);
  const int c4 = C >> 2;
  for (int rr = 0; rr < GMB_ROWS; ++rr) {
    const int j = (blockIdx.x * GMB_WARPS + warp) * GMB_ROWS + rr;
    if (j >= N) return;
    float4 acc[GMB_MAXV];
#pragma unroll
    for (int v = 0; v < GMB_MAXV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q0 = 0; q0 < T4; q0 += 32) {      // 128 table entries per step
      unsigned hit = 0;
      if (q0 + lane < T4) {
        const int4 e = tab4[q0 + lane];
        hit = (e.x == j ? 1u : 0u) | (e.y == j ? 2u : 0u) | (e.z == j ? 4u : 0u) | (e.w == j ? 8u : 0u);
      }
      unsigned any = __ballot_sync(0xffffffffu, hit != 0u);
      while (any) {                             // ascending (r, slot) order
        const int L = __ffs(any) - 1;
        any &= any - 1;
        unsigned h = __shfl_sync(0xffffffffu, hit, L);
        while (h) {
          const int tt = 4 * (q0 + L) + __ffs(h) - 1;
          h &= h - 1;
          const int r = tt / kuse;
          const uint32_t sl = (uint32_t)(tt - r * kuse);
          const size_t o = ((size_t)b * R + r) * C;
#pragma unroll
          for (int v = 0; v < GMB_MAXV; ++v) {
            const int q = lane + 32 * v;
            if (q < c4) {
              const uint32_t am = __ldg(reinterpret_cast<const uint32_t*>(argmax + o) + q);
              const float4 g = __ldg(reinterpret_cast<const float4*>(gout + o) + q);
              if ((am & 0xffu) == sl) acc[v].x += g.x;
              if (((am >> 8) & 0xffu) == sl) acc[v].y += g.y;
              if (((am >> 16) & 0xffu) == sl) acc[v].z += g.z;
              if ((am >> 24) == sl) acc[v].w += g.w;
            }
          }
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(gfeat + ((size_t)b * N + j) * C);
#pragma unroll
    for (int v = 0; v < GMB_MAXV; ++v) {
      const int q = lane + 32 * v;
      if (q < c4) dst[q] = acc[v];
    }
  }
}

// Stage 1 of ORL: per-tile partial sums over points of max over neighbours.
__global__ void __launch_bounds__(GO_THREADS)
orl_partial_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, int N, int C,
                   int k, int tile, float* __restrict__ partial, uint8_t* __restrict__ argmax) {
  const int b = blockIdx.y;
  const int c = blockIdx.z * GO_THREADS + threadIdx.x;
  if (c >= C) return;
  const float* fb = feat + (size_t)b * N * C;
  const int i_end = min((int)(blockIdx.x + 1) * tile, N);
  float sum = 0.0f;
  for (int i = blockIdx.x * tile; i < i_end; ++i) {
    const int32_t* ip = idx + ((size_t)b * N + i) * k;
    float m = -INFINITY;
    int am = 0;
    for (int n = 0; n < k; ++n) {
      float v = __ldg(fb + (size_t)__ldg(ip + n) * C + c);
      if (v > m) { m = v; am = n; }
    }
    sum += m;
    if (argmax) argmax[((size_t)b * N + i) * C + c] = (uint8_t)am;
  }
  partial[((size_t)b * gridDim.x + blockIdx.x) * C + c] = sum;
}
// Stage 2: fixed-order sum of the tile partials (deterministic), mean over N.
__global__ void orl_finish_kernel(const float* __restrict__ partial, int tiles, int C, int N,
                                  float* __restrict__ G) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int t = 0; t < tiles; ++t) s += (double)partial[((size_t)b * tiles + t) * C + c];
  G[(size_t)b * C + c] = (float)(s / (double)N);
}

// ORL, shared-memory form: one CTA owns (object, 16-channel slice).  The slice of the feature
// map (N x 16 fp32, 64 B per point) is staged ONCE into shared memory with coalesced 16-byte
// loads; every gathered row is then a conflict-free 64-byte shared-memory read by a half-warp
// (lane = channel, the two halves of a warp work on two points), instead of k dependent L2
// round trips per point.  The per-object mean is finished inside the CTA in a fixed order
// (deterministic), so the whole op is one launch and needs no workspace.
constexpr int ORL_CS = 16;
constexpr int ORL_THREADS = 256;
constexpr int ORL_GROUPS = ORL_THREADS / 4;   // point groups per CTA (4 lanes x 4 channels = 16 channels)

template <bool AM>
__device__ __forceinline__ void orl_take(const float4 v, int n, float4& m, int (&am)[4]) {
  if (AM) {
    if (v.x > m.x) { m.x = v.x; am[0] = n; }
    if (v.y > m.y) { m.y = v.y; am[1] = n; }
    if (v.z > m.z) { m.z = v.z; am[2] = n; }
    if (v.w > m.w) { m.w = v.w; am[3] = n; }
  } else {
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
}

template <bool AM>
__global__ void __launch_bounds__(ORL_THREADS)
orl_smem_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, int N, int C,
                int k, float* __restrict__ G, uint8_t* __restrict__ argmax) {
  extern __shared__ __align__(16) float s_f[];            // [N][16]
  __shared__ float s_sum[ORL_GROUPS][ORL_CS];
  const int b = blockIdx.y, c0 = blockIdx.x * ORL_CS;
  const int tid = threadIdx.x;
  const float* fb = feat + (size_t)b * N * C + c0;
  for (int e = tid; e < N * (ORL_CS / 4); e += ORL_THREADS) {
    const int row = e / (ORL_CS / 4), q = e % (ORL_CS / 4);
    reinterpret_cast<float4*>(s_f)[e] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)row * C) + q);
  }
  __syncthreads();
  // thread = (point group, channel quad): one LDS.128 per gathered row serves 4 channels
  const int grp = tid >> 2, q = tid & 3;
  const float4* sf4 = reinterpret_cast<const float4*>(s_f) + q;
  const int32_t* ib = idx + (size_t)b * N * k;
  const int k4 = k >> 2;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = grp; p < N; p += ORL_GROUPS) {
    const int32_t* ip = ib + (size_t)p * k;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int am[4] = {0, 0, 0, 0};
    if ((k & 3) == 0 && ((size_t)p * k & 3) == 0) {     // 16-byte aligned index row
      const int4* ip4 = reinterpret_cast<const int4*>(ip);
      int4 jj = __ldg(ip4);
      for (int t = 0; t < k4; ++t) {
        const int4 cur = jj;
        if (t + 1 < k4) jj = __ldg(ip4 + t + 1);
        orl_take<AM>(sf4[cur.x * (ORL_CS / 4)], 4 * t, m, am);
        orl_take<AM>(sf4[cur.y * (ORL_CS / 4)], 4 * t + 1, m, am);
        orl_take<AM>(sf4[cur.z * (ORL_CS / 4)], 4 * t + 2, m, am);
        orl_take<AM>(sf4[cur.w * (ORL_CS / 4)], 4 * t + 3, m, am);
      }
    } else {
      for (int n = 0; n < k; ++n) orl_take<AM>(sf4[__ldg(ip + n) * (ORL_CS / 4)], n, m, am);
    }
    sum.x += m.x; sum.y += m.y; sum.z += m.z; sum.w += m.w;
    if (AM)
      *reinterpret_cast<uchar4*>(argmax + ((size_t)b * N + p) * C + c0 + 4 * q) =
          make_uchar4((unsigned char)am[0], (unsigned char)am[1], (unsigned char)am[2], (unsigned char)am[3]);
  }
  *reinterpret_cast<float4*>(&s_sum[grp][4 * q]) = sum;
  __syncthreads();
  if (tid < ORL_CS) {
    double t = 0.0;
    for (int g = 0; g < ORL_GROUPS; ++g) t += (double)s_sum[g][tid];
    G[(size_t)b * C + c0 + tid] = (float)(t / (double)N);
  }
}

__global__ void __launch_bounds__(GO_THREADS)
orl_bwd_kernel(const float* __restrict__ gG, const int32_t* __restrict__ idx,
               const uint8_t* __restrict__ argmax, int N, int C, int k,
               float* __restrict__ gfeat) {
  const int b = blockIdx.y;
  const int c = blockIdx.z * GO_THREADS + threadIdx.x;
  if (c >= C) return;
  const float g = gG[(size_t)b * C + c] / (float)N;
  const int i_end = min((int)(blockIdx.x + 1) * GO_PT, N);
  for (int i = blockIdx.x * GO_PT; i < i_end; ++i) {
    const int src = __ldg(idx + ((size_t)b * N + i) * k + argmax[((size_t)b * N + i) * C + c]);
    atomicAdd(gfeat + ((size_t)b * N + src) * C + c, g);
  }
}

template <typename TO> __device__ __forceinline__ TO from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TO>
__global__ void __launch_bounds__(GO_THREADS)
upsample_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ nn, int Nsrc,
                    int M, int C, TO* __restrict__ out, int ldo, int col0) {
  const int b = blockIdx.y;
  const int i_end = min((int)(blockIdx.x + 1) * GO_PT, M);
  for (int i = blockIdx.x * GO_PT; i < i_end; ++i) {
    const int r = nn ? __ldg(nn + (size_t)b * M + i) : (Nsrc == 1 ? 0 : i);
    const float* src = feat + ((size_t)b * Nsrc + r) * C;
    TO* dst = out + ((size_t)b * M + i) * ldo + col0;
    for (int c = threadIdx.x; c < C; c += GO_THREADS) dst[c] = from_f32<TO>(__ldg(src + c));
  }
}
// nn != NULL: scatter-add (float atomics; a source row is hit by ~4 points).
// nn == NULL (identity): plain converting copy of the column slice.
template <typename TO>
__global__ void __launch_bounds__(GO_THREADS)
upsample_bwd_kernel(const TO* __restrict__ gout, const int32_t* __restrict__ nn, int Nsrc,
                    int M, int C, int ldo, int col0, float* __restrict__ gfeat) {
  const int b = blockIdx.y;
  const int i_end = min((int)(blockIdx.x + 1) * GO_PT, M);
  for (int i = blockIdx.x * GO_PT; i < i_end; ++i) {
    const TO* src = gout + ((size_t)b * M + i) * ldo + col0;
    if (nn) {
      float* dst = gfeat + ((size_t)b * Nsrc + __ldg(nn + (size_t)b * M + i)) * C;
      for (int c = threadIdx.x; c < C; c += GO_THREADS) atomicAdd(dst + c, to_f32(src[c]));
    } else {
      float* dst = gfeat + ((size_t)b * Nsrc + i) * C;
      for (int c = threadIdx.x; c < C; c += GO_THREADS) dst[c] = to_f32(src[c]);
    }
  }
}

// Vectorised forms (C % 8 == 0, 16-byte aligned rows): one thread moves 8 channels.
template <typename TO> struct Out8;
template <> struct Out8<float> {
  static __device__ __forceinline__ void store(float* p, const float4 a, const float4 b) {
    reinterpret_cast<float4*>(p)[0] = a;
    reinterpret_cast<float4*>(p)[1] = b;
  }
  static __device__ __forceinline__ void load(const float* p, float4& a, float4& b) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
};
template <> struct Out8<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float4 a, const float4 b) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(b.x, b.y); h[3] = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float4& a, float4& b) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
    a = make_float4(f0.x, f0.y, f1.x, f1.y);
    b = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
};

template <typename TO>
__global__ void __launch_bounds__(256)
upsample_fwd_vec_kernel(const float* __restrict__ feat, const int32_t* __restrict__ nn, int Nsrc,
                        int M, int C, TO* __restrict__ out, int ldo, int col0) {
  const int b = blockIdx.y, c8 = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * c8) return;
  const int i = t / c8, v = t % c8;
  const int r = nn ? __ldg(nn + (size_t)b * M + i) : (Nsrc == 1 ? 0 : i);
  const float4* src = reinterpret_cast<const float4*>(feat + ((size_t)b * Nsrc + r) * C) + 2 * v;
  Out8<TO>::store(out + ((size_t)b * M + i) * ldo + col0 + 8 * v, __ldg(src), __ldg(src + 1));
}

template <typename TO>
__global__ void __launch_bounds__(256)
upsample_bwd_vec_kernel(const TO* __restrict__ gout, const int32_t* __restrict__ nn, int Nsrc,
                        int M, int C, int ldo, int col0, float* __restrict__ gfeat) {
  const int b = blockIdx.y, c8 = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * c8) return;
  const int i = t / c8, v = t % c8;
  float4 a, c;
  Out8<TO>::load(gout + ((size_t)b * M + i) * ldo + col0 + 8 * v, a, c);
  if (nn) {   // scatter-add: one 16-byte vector atomic per 4 channels (sm_90+)
    float4* dst = reinterpret_cast<float4*>(gfeat + ((size_t)b * Nsrc + __ldg(nn + (size_t)b * M + i)) * C) + 2 * v;
    atomicAdd(dst, a);
    atomicAdd(dst + 1, c);
  } else {
    Out8<float>::store(gfeat + ((size_t)b * Nsrc + i) * C + 8 * v, a, c);
  }
}

// Deterministic form of the scatter-add (the default when nn != NULL): warp = one SOURCE row j.  The object's
// nearest table (M ints) sits in shared memory; the warp scans it 32 targets at a time (ballot) and adds the rows
// of the targets that map to j in ascending target order.  Every gfeat row is written exactly once (zeros when no
// target maps to it): no float atomics, no zero-fill, bit-reproducible.
constexpr int UBD_WARPS = 16;
constexpr int UBD_ROWS = 1;                    // source rows per warp
constexpr int UBD_MAXV = 4;                    // 8-channel vectors per lane: C <= 1024
template <typename TO>
__global__ void __launch_bounds__(UBD_WARPS * 32)
upsample_bwd_det_kernel(const TO* __restrict__ gout, const int32_t* __restrict__ nn, int Nsrc, int M, int C,
                        int ldo, int col0, float* __restrict__ gfeat) {
  extern __shared__ __align__(16) int32_t s_nn[];    // [M], padded to a multiple of 4 with -1
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c8 = C >> 3;
  const int M4 = (M + 3) >> 2;
  for (int t = threadIdx.x; t < 4 * M4; t += UBD_WARPS * 32) s_nn[t] = t < M ? __ldg(nn + (size_t)b * M + t) : -1;
  __syncthreads();
  const int4* nn4 = reinterpret_cast<const int4*>(s_nn);
  for (int rr = 0; rr < UBD_ROWS; ++rr) {
    const int j = (blockIdx.x * UBD_WARPS + warp) * UBD_ROWS + rr;
    if (j >= Nsrc) return;
    float4 aa[UBD_MAXV], ab[UBD_MAXV];
#pragma unroll
    for (int v = 0; v < UBD_MAXV; ++v) aa[v] = ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q0 = 0; q0 < M4; q0 += 32) {      // 128 targets per step
      unsigned hit = 0;
      if (q0 + lane < M4) {
        const int4 e = nn4[q0 + lane];
        hit = (e.x == j ? 1u : 0u) | (e.y == j ? 2u : 0u) | (e.z == j ? 4u : 0u) | (e.w == j ? 8u : 0u);
      }
      unsigned any = __ballot_sync(0xffffffffu, hit != 0u);
      while (any) {                             // ascending target order
        const int L = __ffs(any) - 1;
        any &= any - 1;
        unsigned h = __shfl_sync(0xffffffffu, hit, L);
        while (h) {
          const int i = 4 * (q0 + L) + __ffs(h) - 1;
          h &= h - 1;
          const TO* src = gout + ((size_t)b * M + i) * ldo + col0;
#pragma unroll
          for (int v = 0; v < UBD_MAXV; ++v) {
            const int q = lane + 32 * v;
            if (q < c8) {
              float4 a, c;
              Out8<TO>::load(src + 8 * q, a, c);
              aa[v].x += a.x; aa[v].y += a.y; aa[v].z += a.z; aa[v].w += a.w;
              ab[v].x += c.x; ab[v].y += c.y; ab[v].z += c.z; ab[v].w += c.w;
            }
          }
        }
      }
    }
    float* dst = gfeat + ((size_t)b * Nsrc + j) * C;
#pragma unroll
    for (int v = 0; v < UBD_MAXV; ++v) {
      const int q = lane + 32 * v;
      if (q < c8) Out8<float>::store(dst + 8 * q, aa[v], ab[v]);
    }
  }
}

static bool vec8_ok(const void* a, const void* b, int C, int ldo, int col0, int esz) {
  return (C % 8) == 0 && ((size_t)ldo * esz) % 16 == 0 && ((size_t)col0 * esz) % 16 == 0 &&
         ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0;
}

// Max over the points of one object (the `torch.max(x, 2)` that feeds the last block of every pose
// head, reference PoseR.py:30 / PoseTs.py:35) on a (B, N, C) bf16 / fp32 activation: CTA = (object,
// 64-channel slab), thread = 8 channels x a row lane, 16-byte loads, fixed-order combine in shared
// memory (ties -> lowest row, deterministic).  out (B, C) same dtype as x, arg (B, C) int32.
template <typename T>
__global__ void __launch_bounds__(256)
colmax_kernel(const T* __restrict__ x, int N, int C, T* __restrict__ out, int32_t* __restrict__ arg) {
  __shared__ float s_v[32][64];
  __shared__ int s_i[32][64];
  const int b = blockIdx.y, cv = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cv * 8;
  float m[8];
  int am[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; am[i] = 0; }
  if (c0 < C) {
    const T* xb = x + (size_t)b * N * C + c0;
#pragma unroll 2
    for (int r = rl; r < N; r += 32) {
      float4 a, c;
      Out8<T>::load(xb + (size_t)r * C, a, c);
      const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (v[i] > m[i]) { m[i] = v[i]; am[i] = r; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { s_v[rl][cv * 8 + i] = m[i]; s_i[rl][cv * 8 + i] = am[i]; }
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < C) {
    float best = s_v[0][threadIdx.x];
    int bi = s_i[0][threadIdx.x];
    for (int l = 1; l < 32; ++l) {
      const float v = s_v[l][threadIdx.x];
      const int vi = s_i[l][threadIdx.x];
      if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
    }
    out[(size_t)b * C + blockIdx.x * 64 + threadIdx.x] = from_f32<T>(best);
    arg[(size_t)b * C + blockIdx.x * 64 + threadIdx.x] = bi;
  }
}

static int orl_tile(int N) { return 16; }

}  // namespace hsp

extern "C" int hsp_gather_max_fwd(const float* feat, const int32_t* idx, const int32_t* rows,
                                  int B, int N, int C, int R, int kuse, int kstride,
                                  float* out, uint8_t* argmax, void* stream) {
  using namespace hsp;
  if (!feat || !idx || !out || B < 0 || N <= 0 || C <= 0 || R < 0 || kuse <= 0 ||
      kuse > kstride || kuse > 255 || B > 65535 || (!rows && R != N))
    return HSP_EINVAL;
  if (B == 0 || R == 0) return HSP_OK;
  dim3 grid((R + GO_PT - 1) / GO_PT, B, (C + GO_THREADS - 1) / GO_THREADS);
  gather_max_fwd_kernel<<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(
      feat, idx, rows, N, C, R, kuse, kstride, out, argmax);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_gather_max_bwd(const float* gout, const int32_t* idx, const int32_t* rows,
                                  const uint8_t* argmax, int B, int N, int C, int R, int kuse,
                                  int kstride, float* gfeat, void* stream) {
  using namespace hsp;
  if (!gout || !idx || !argmax || !gfeat || B < 0 || N <= 0 || C <= 0 || R < 0 || kuse <= 0 ||
      kuse > kstride || B > 65535 || (!rows && R != N))
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tab = (size_t)R * kuse * sizeof(int32_t);
  if (R > 0 && (C % 4) == 0 && C <= 128 * GMB_MAXV && tab + 16 <= 48 * 1024 && ((uintptr_t)gout % 16) == 0 &&
      ((uintptr_t)gfeat % 16) == 0 && ((uintptr_t)argmax % 4) == 0) {
    gather_max_bwd_det_kernel<<<dim3((N + GMB_WARPS * GMB_ROWS - 1) / (GMB_WARPS * GMB_ROWS), B), GMB_WARPS * 32,
                                tab + 16, st>>>(
        gout, idx, rows, argmax, N, C, R, kuse, kstride, gfeat);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  // other shapes: float atomics into a zeroed buffer (order-dependent in the last bit)
  if (cudaMemsetAsync(gfeat, 0, (size_t)B * N * C * sizeof(float), st) != cudaSuccess) return HSP_ELAUNCH;
  if (R == 0) return HSP_OK;
  dim3 grid((R + GO_PT - 1) / GO_PT, B, (C + GO_THREADS - 1) / GO_THREADS);
  gather_max_bwd_kernel<<<grid, GO_THREADS, 0, st>>>(gout, idx, rows, argmax, N, C, R, kstride, gfeat);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" size_t hsp_orl_global_workspace_bytes(int B, int N, int C) {
  using namespace hsp;
  if (B <= 0 || N <= 0 || C <= 0) return 0;
  int tile = orl_tile(N);
  return (size_t)B * ((N + tile - 1) / tile) * C * sizeof(float);
}

extern "C" int hsp_orl_global_fwd(const float* feat, const int32_t* idx, int B, int N, int C,
                                  int k, float* G, uint8_t* argmax, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  using namespace hsp;
  if (!feat || !idx || !G || B < 0 || N <= 0 || C <= 0 || k <= 0 || k > 255 || B > 65535)
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  if (!workspace || workspace_bytes < hsp_orl_global_workspace_bytes(B, N, C))
    return HSP_EWORKSPACE;
  const int tile = orl_tile(N), tiles = (N + tile - 1) / tile;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)N * ORL_CS * sizeof(float);
  if ((C % ORL_CS) == 0 && smem <= 200 * 1024 && (((uintptr_t)argmax) & 3) == 0) {
    if (smem > 48 * 1024 &&
        (cudaFuncSetAttribute(orl_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem) != cudaSuccess ||
         cudaFuncSetAttribute(orl_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem) != cudaSuccess))
      return HSP_ELAUNCH;
    if (argmax)
      orl_smem_kernel<true><<<dim3(C / ORL_CS, B), ORL_THREADS, smem, st>>>(feat, idx, N, C, k, G, argmax);
    else
      orl_smem_kernel<false><<<dim3(C / ORL_CS, B), ORL_THREADS, smem, st>>>(feat, idx, N, C, k, G, nullptr);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  dim3 grid(tiles, B, (C + GO_THREADS - 1) / GO_THREADS);
  orl_partial_kernel<<<grid, GO_THREADS, 0, st>>>(feat, idx, N, C, k, tile, (float*)workspace,
                                                  argmax);
  HSP_LAUNCH_CHECK();
  dim3 grid2((C + 127) / 128, B);
  orl_finish_kernel<<<grid2, 128, 0, st>>>((const float*)workspace, tiles, C, N, G);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_orl_global_bwd(const float* gG, const int32_t* idx, const uint8_t* argmax,
                                  int B, int N, int C, int k, float* gfeat, void* stream) {
  using namespace hsp;
  if (!gG || !idx || !argmax || !gfeat || B < 0 || N <= 0 || C <= 0 || k <= 0 || B > 65535)
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  dim3 grid((N + GO_PT - 1) / GO_PT, B, (C + GO_THREADS - 1) / GO_THREADS);
  orl_bwd_kernel<<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(gG, idx, argmax, N, C, k, gfeat);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_upsample_rows_fwd(const float* feat, const int32_t* nn, int B, int Nsrc,
                                     int M, int C, void* out, int ldo, int col0, int out_dtype,
                                     void* stream) {
  using namespace hsp;
  if (!feat || !out || B < 0 || Nsrc <= 0 || M < 0 || C <= 0 || col0 < 0 ||
      col0 + C > ldo || B > 65535)
    return HSP_EINVAL;
  if (!nn && Nsrc != 1 && Nsrc != M) return HSP_EINVAL;  /* identity / broadcast modes */
  if (out_dtype != HSP_DTYPE_F32 && out_dtype != HSP_DTYPE_BF16) return HSP_EINVAL;
  if (B == 0 || M == 0) return HSP_OK;
  if (vec8_ok(feat, out, C, ldo, col0, out_dtype == HSP_DTYPE_BF16 ? 2 : 4)) {
    dim3 gv((M * (C / 8) + 255) / 256, B);
    if (out_dtype == HSP_DTYPE_BF16)
      upsample_fwd_vec_kernel<__nv_bfloat16><<<gv, 256, 0, (cudaStream_t)stream>>>(
          feat, nn, Nsrc, M, C, (__nv_bfloat16*)out, ldo, col0);
    else
      upsample_fwd_vec_kernel<float><<<gv, 256, 0, (cudaStream_t)stream>>>(feat, nn, Nsrc, M, C,
                                                                             (float*)out, ldo, col0);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  dim3 grid((M + GO_PT - 1) / GO_PT, B);
  if (out_dtype == HSP_DTYPE_BF16)
    upsample_fwd_kernel<__nv_bfloat16><<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(
        feat, nn, Nsrc, M, C, (__nv_bfloat16*)out, ldo, col0);
  else
    upsample_fwd_kernel<float><<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(
        feat, nn, Nsrc, M, C, (float*)out, ldo, col0);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_upsample_rows_bwd(const void* gout, const int32_t* nn, int B, int Nsrc,
                                     int M, int C, int ldo, int col0, int gout_dtype,
                                     float* gfeat, void* stream) {
  using namespace hsp;
  if (!gout || !gfeat || B < 0 || Nsrc <= 0 || M < 0 || C <= 0 || col0 < 0 ||
      col0 + C > ldo || B > 65535)
    return HSP_EINVAL;
  if (!nn && Nsrc != M) return HSP_EINVAL;
  if (gout_dtype != HSP_DTYPE_F32 && gout_dtype != HSP_DTYPE_BF16) return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = vec8_ok(gout, gfeat, C, ldo, col0, gout_dtype == HSP_DTYPE_BF16 ? 2 : 4);
  if (nn && M > 0 && vec && C <= 256 * UBD_MAXV && (size_t)M * sizeof(int32_t) + 16 <= 48 * 1024) {
    // scatter-add as a gather over the source rows: deterministic, writes every gfeat row once
    dim3 gd((Nsrc + UBD_WARPS * UBD_ROWS - 1) / (UBD_WARPS * UBD_ROWS), B);
    const size_t sm = (size_t)M * sizeof(int32_t) + 16;
    if (gout_dtype == HSP_DTYPE_BF16)
      upsample_bwd_det_kernel<__nv_bfloat16><<<gd, UBD_WARPS * 32, sm, st>>>((const __nv_bfloat16*)gout, nn, Nsrc, M, C,
                                                                         ldo, col0, gfeat);
    else
      upsample_bwd_det_kernel<float><<<gd, UBD_WARPS * 32, sm, st>>>((const float*)gout, nn, Nsrc, M, C, ldo, col0,
                                                                 gfeat);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  // the float-atomics scatter accumulates: it starts from zeros (the identity copy overwrites everything)
  if (nn && cudaMemsetAsync(gfeat, 0, (size_t)B * Nsrc * C * sizeof(float), st) != cudaSuccess) return HSP_ELAUNCH;
  if (M == 0) return HSP_OK;
  if (vec) {
    dim3 gv((M * (C / 8) + 255) / 256, B);
    if (gout_dtype == HSP_DTYPE_BF16)
      upsample_bwd_vec_kernel<__nv_bfloat16><<<gv, 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)gout, nn, Nsrc, M, C, ldo, col0, gfeat);
    else
      upsample_bwd_vec_kernel<float><<<gv, 256, 0, (cudaStream_t)stream>>>(
          (const float*)gout, nn, Nsrc, M, C, ldo, col0, gfeat);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  dim3 grid((M + GO_PT - 1) / GO_PT, B);
  if (gout_dtype == HSP_DTYPE_BF16)
    upsample_bwd_kernel<__nv_bfloat16><<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)gout, nn, Nsrc, M, C, ldo, col0, gfeat);
  else
    upsample_bwd_kernel<float><<<grid, GO_THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)gout, nn, Nsrc, M, C, ldo, col0, gfeat);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_colmax_fwd(const void* x, int dtype, int B, int N, int C, void* out, int32_t* arg,
                              void* stream) {
  using namespace hsp;
  if (!x || !out || !arg || B < 0 || N <= 0 || C <= 0 || (C % 8) != 0 || B > 65535 ||
      (((uintptr_t)x) & 15) != 0)
    return HSP_EINVAL;
  if (dtype != HSP_DTYPE_F32 && dtype != HSP_DTYPE_BF16) return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  dim3 grid((C + 63) / 64, B);
  if (dtype == HSP_DTYPE_BF16)
    colmax_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, N, C, (__nv_bfloat16*)out, arg);
  else
    colmax_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, N, C, (float*)out, arg);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
