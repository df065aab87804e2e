// K1 — fused 3-D pairwise distance + top-k (never materialises the MxN matrix).
//
// Replaces get_neighbor_index (reference network/fs_net_repo/gcn3d.py:15-24)
// and get_nearest_index (gcn3d.py:27-36).  One CTA = one object x a tile of
// queries.  The object's candidate set is staged once in shared memory as
// float4 (x, y, z, |p|^2); one WARP owns one query at a time: each lane
// evaluates the reference's exact FP32 expression for one candidate
// (conflict-free LDS.128), candidates below the running K-th distance are
// compacted into a per-warp shared-memory queue with ballot/popc, and every 32
// queued candidates are bitonic-sorted and merged into the register-resident
// sorted list with warp shuffles (WarpTopK, common.cuh).
//
// Arithmetic (bit-exact with the reference's CPU path, pinned by
// tests/golden):  inner = fma(a2,b2, fma(a1,b1, a0*b0)),  |p|^2 = (x*x+y*y)+z*z,
//   NEIGHBOR: d = ((-2*inner) + |c_j|^2) + |q_i|^2
//   NEAREST : d = (|c_j|^2 + |q_i|^2) - 2*inner
#include "common.cuh"

namespace hsp {

constexpr int KNN3_THREADS = 256;
constexpr int KNN3_WARPS = KNN3_THREADS / 32;
constexpr int KNN3_UNROLL = 4;

__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

template <int FORMULA>
__device__ __forceinline__ float dist3(float qx, float qy, float qz, float qq, float4 c) {
  float t = __fmul_rn(qx, c.x);
  t = __fmaf_rn(qy, c.y, t);
  t = __fmaf_rn(qz, c.z, t);
  if (FORMULA == HSP_DIST_NEIGHBOR)
    return __fadd_rn(__fadd_rn(__fmul_rn(t, -2.0f), c.w), qq);
  else
    return __fsub_rn(__fadd_rn(c.w, qq), __fmul_rn(t, 2.0f));
}

template <int NL, int FORMULA>
__global__ void __launch_bounds__(KNN3_THREADS)
knn3_kernel(const float* __restrict__ query, const float* __restrict__ cand, int M, int N,
            int K, int drop, int qtile, int64_t* __restrict__ idx64,
            int32_t* __restrict__ idx32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_c = reinterpret_cast<float4*>(smem_raw);
  uint64_t* s_q = reinterpret_cast<uint64_t*>(smem_raw + (size_t)N * sizeof(float4));

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cb = cand + (size_t)b * N * 3;
  for (int j = tid; j < N; j += KNN3_THREADS) {
    float x = cb[3 * j], y = cb[3 * j + 1], z = cb[3 * j + 2];
    s_c[j] = make_float4(x, y, z, sqnorm3(x, y, z));
  }
  __syncthreads();

  WarpTopK<NL> top;
  const int q_end = min((int)(blockIdx.x + 1) * qtile, M);
  const int k_out = K - drop;
  for (int qi = blockIdx.x * qtile + warp; qi < q_end; qi += KNN3_WARPS) {
    const float* qp = query + ((size_t)b * M + qi) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const float qq = sqnorm3(qx, qy, qz);
    top.reset(s_q + warp * 64);
    for (int base = 0; base < N; base += 32 * KNN3_UNROLL) {
      uint64_t key[KNN3_UNROLL];
      bool any = false;
#pragma unroll
      for (int u = 0; u < KNN3_UNROLL; ++u) {
        int j = base + u * 32 + lane;
        key[u] = KEY_MAX;
        if (j < N) key[u] = make_key(dist3<FORMULA>(qx, qy, qz, qq, s_c[j]), (uint32_t)j);
        any |= key[u] < top.thr;
      }
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int u = 0; u < KNN3_UNROLL; ++u) top.push(key[u], lane, K);
      }
    }
    top.finish(lane, K);
    const size_t o = ((size_t)b * M + qi) * k_out;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      int r = l * 32 + lane - drop;
      if (r >= 0 && r < k_out) {
        uint32_t j = (uint32_t)(top.L[l] & 0xffffffffu);
        if (idx64) idx64[o + r] = (int64_t)j;
        if (idx32) idx32[o + r] = (int32_t)j;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// v2: distances stay in REGISTERS and the threshold comes first.
//
// The streaming kernel above spends ~85 % of its instructions in the selection (ncu: ALU pipe
// 83 %): every candidate under the running K-th distance costs a ballot/queue round and every
// 32 of them a 64-bit bitonic sort+merge, ~K(1+ln(N/K)) ~ 100 events per query at N = 1028.
// Here a warp evaluates a block of 32*RPL candidates into RPL registers per lane (orderable
// 32-bit distances), keeps two running minima per lane (even / odd chunks = 64 "slots") and
// takes the K-th smallest of the 64 slot minima as an upper bound tau on the K-th smallest
// distance of the block (K distinct candidates are <= tau).  Only candidates with d <= tau
// survive (~K + 5 of them): they are compacted lane-locally into a small shared-memory queue and
// sorted ONCE.  Same keys (distance, index), same total order, same result — bit-identical to
// the streaming kernel and the oracle.
// ---------------------------------------------------------------------------
constexpr int KNN3_QCAP = 128;   // per-warp survivor queue (keys); more survivors -> streaming fallback

template <int NL, int FORMULA, int RPL>
__global__ void __launch_bounds__(KNN3_THREADS)
knn3_reg_kernel(const float* __restrict__ query, const float* __restrict__ cand, int M, int N,
                int K, int drop, int qtile, int64_t* __restrict__ idx64,
                int32_t* __restrict__ idx32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_c = reinterpret_cast<float4*>(smem_raw);
  uint64_t* s_q = reinterpret_cast<uint64_t*>(smem_raw + (size_t)N * sizeof(float4));

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cb = cand + (size_t)b * N * 3;
  for (int j = tid; j < N; j += KNN3_THREADS) {
    float x = cb[3 * j], y = cb[3 * j + 1], z = cb[3 * j + 2];
    s_c[j] = make_float4(x, y, z, sqnorm3(x, y, z));
  }
  __syncthreads();

  WarpTopK<NL> top;
  uint64_t* q = s_q + warp * KNN3_QCAP;
  const int q_end = min((int)(blockIdx.x + 1) * qtile, M);
  const int k_out = K - drop;
  for (int qi = blockIdx.x * qtile + warp; qi < q_end; qi += KNN3_WARPS) {
    const float* qp = query + ((size_t)b * M + qi) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const float qq = sqnorm3(qx, qy, qz);
    top.reset(q);
    for (int base = 0; base < N; base += 32 * RPL) {
      uint32_t o[RPL];
      uint32_t mA = 0xffffffffu, mB = 0xffffffffu;
#pragma unroll
      for (int c = 0; c < RPL; ++c) {
        const int j = base + 32 * c + lane;
        o[c] = 0xffffffffu;
        if (j < N) o[c] = float_orderable(dist3<FORMULA>(qx, qy, qz, qq, s_c[j]));
        if (c & 1) mB = min(mB, o[c]); else mA = min(mA, o[c]);
      }
      // upper bound on the K-th smallest distance of this block, tightened by the running list
      const uint32_t tau_blk = warp_kth_smallest64(mA, mB, K, lane);
      const uint64_t cap = umin64(top.thr, ((uint64_t)tau_blk << 32) | 0xffffffffull);
      const uint32_t tau = (uint32_t)(cap >> 32);
      // lane-local count of survivors, exclusive scan over lanes
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < RPL; ++c) cnt += (o[c] <= tau && base + 32 * c + lane < N) ? 1 : 0;
      int off = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, off, d);
        if (lane >= d) off += t;
      }
      const int total = __shfl_sync(0xffffffffu, off, 31);
      off -= cnt;
      if (total <= KNN3_QCAP) {
#pragma unroll
        for (int c = 0; c < RPL; ++c) {
          const int j = base + 32 * c + lane;
          if (o[c] <= tau && j < N) q[off++] = ((uint64_t)o[c] << 32) | (uint32_t)j;
        }
        __syncwarp();
        for (int t = 0; t < total; t += 32) {
          const uint64_t key = (t + lane < total) ? q[t + lane] : KEY_MAX;
          top.merge(key, lane, K);
        }
        __syncwarp();
      } else {   // heavy ties / duplicates: stream this block through the queue-and-merge path
        top.thr = cap;
#pragma unroll
        for (int c = 0; c < RPL; ++c) {
          const int j = base + 32 * c + lane;
          const uint64_t key = (j < N) ? (((uint64_t)o[c] << 32) | (uint32_t)j) : KEY_MAX;
          if (__any_sync(0xffffffffu, key < top.thr)) {
            top.push(key, lane, K);
            top.thr = umin64(top.thr, cap);
          }
        }
        top.finish(lane, K);
      }
    }
    const size_t oo = ((size_t)b * M + qi) * k_out;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      int r = l * 32 + lane - drop;
      if (r >= 0 && r < k_out) {
        uint32_t j = (uint32_t)(top.L[l] & 0xffffffffu);
        if (idx64) idx64[oo + r] = (int64_t)j;
        if (idx32) idx32[oo + r] = (int32_t)j;
      }
    }
  }
}

// K = 1 (get_nearest_index, gcn3d.py:27-36; k = 1 queries in general): no selection network at all —
// every lane keeps the smallest (distance, index) key of its candidates, one warp min-reduction.
template <int FORMULA>
__global__ void __launch_bounds__(KNN3_THREADS)
knn3_nearest_kernel(const float* __restrict__ query, const float* __restrict__ cand, int M, int N,
                    int qtile, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_c = reinterpret_cast<float4*>(smem_raw);
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cb = cand + (size_t)b * N * 3;
  for (int j = tid; j < N; j += KNN3_THREADS) {
    float x = cb[3 * j], y = cb[3 * j + 1], z = cb[3 * j + 2];
    s_c[j] = make_float4(x, y, z, sqnorm3(x, y, z));
  }
  __syncthreads();
  const int q_end = min((int)(blockIdx.x + 1) * qtile, M);
  for (int qi = blockIdx.x * qtile + warp; qi < q_end; qi += KNN3_WARPS) {
    const float* qp = query + ((size_t)b * M + qi) * 3;
    const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    const float qq = sqnorm3(qx, qy, qz);
    uint64_t best = KEY_MAX;
#pragma unroll 4
    for (int j = lane; j < N; j += 32)
      best = umin64(best, make_key(dist3<FORMULA>(qx, qy, qz, qq, s_c[j]), (uint32_t)j));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = umin64(best, shfl_xor_u64(best, d));
    if (lane == 0) {
      const size_t o = (size_t)b * M + qi;
      if (idx64) idx64[o] = (int64_t)(best & 0xffffffffu);
      if (idx32) idx32[o] = (int32_t)(best & 0xffffffffu);
    }
  }
}

template <int FORMULA>
static int launch_knn3_nearest(const float* query, const float* cand, int B, int M, int N,
                               int64_t* idx64, int32_t* idx32, cudaStream_t st) {
  size_t smem = (size_t)N * sizeof(float4);
  auto kern = knn3_nearest_kernel<FORMULA>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return HSP_ELAUNCH;
  int qtile = 128;
  while (qtile > KNN3_WARPS && (long)B * ((M + qtile - 1) / qtile) < 148L * 8) qtile >>= 1;
  dim3 grid((M + qtile - 1) / qtile, B);
  kern<<<grid, KNN3_THREADS, smem, st>>>(query, cand, M, N, qtile, idx64, idx32);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

template <int NL, int FORMULA, int RPL>
static int launch_knn3_reg(const float* query, const float* cand, int B, int M, int N, int K,
                           int drop, int64_t* idx64, int32_t* idx32, cudaStream_t st) {
  size_t smem = (size_t)N * sizeof(float4) + KNN3_WARPS * KNN3_QCAP * sizeof(uint64_t);
  auto kern = knn3_reg_kernel<NL, FORMULA, RPL>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return HSP_ELAUNCH;
  }
  int qtile = 64;
  while (qtile > KNN3_WARPS && (long)B * ((M + qtile - 1) / qtile) < 148L * 8) qtile >>= 1;
  dim3 grid((M + qtile - 1) / qtile, B);
  kern<<<grid, KNN3_THREADS, smem, st>>>(query, cand, M, N, K, drop, qtile, idx64, idx32);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

template <int NL, int FORMULA>
static int dispatch_knn3_reg(const float* query, const float* cand, int B, int M, int N, int K,
                             int drop, int64_t* idx64, int32_t* idx32, cudaStream_t st) {
#define HSP_K3(R) return launch_knn3_reg<NL, FORMULA, R>(query, cand, B, M, N, K, drop, idx64, idx32, st)
  if (N <= 64) HSP_K3(2);
  if (N <= 160) HSP_K3(5);
  if (N <= 288) HSP_K3(9);
  if (N <= 544) HSP_K3(17);
  HSP_K3(33);
#undef HSP_K3
}

template <int NL, int FORMULA>
static int launch_knn3(const float* query, const float* cand, int B, int M, int N, int K,
                       int drop, int64_t* idx64, int32_t* idx32, cudaStream_t st) {
  size_t smem = (size_t)N * sizeof(float4) + KNN3_WARPS * 64 * sizeof(uint64_t);
  auto kern = knn3_kernel<NL, FORMULA>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return HSP_ELAUNCH;
  }
  // One full machine of resident CTAs is 148 SMs x 8 CTAs (2048 threads / SM).
  int qtile = 32;
  while (qtile > KNN3_WARPS && (long)B * ((M + qtile - 1) / qtile) < 148L * 8) qtile >>= 1;
  dim3 grid((M + qtile - 1) / qtile, B);
  kern<<<grid, KNN3_THREADS, smem, st>>>(query, cand, M, N, K, drop, qtile, idx64, idx32);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

}  // namespace hsp

extern "C" int hsp_knn3(const float* query, const float* cand, int B, int M, int N, int k,
                        int drop_first, int formula, int64_t* idx64, int32_t* idx32,
                        void* stream) {
  using namespace hsp;
  if (!query || !cand || (!idx64 && !idx32)) return HSP_EINVAL;
  if (B < 0 || M < 0 || N <= 0 || k <= 0 || drop_first < 0) return HSP_EINVAL;
  const int K = k + drop_first;
  if (K > N || K > 64 || N > 8192) return HSP_EINVAL;
  if (formula != HSP_DIST_NEIGHBOR && formula != HSP_DIST_NEAREST) return HSP_EINVAL;
  if (B == 0 || M == 0) return HSP_OK;
  if (B > 65535) return HSP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 1) {
    if (formula == HSP_DIST_NEIGHBOR)
      return launch_knn3_nearest<HSP_DIST_NEIGHBOR>(query, cand, B, M, N, idx64, idx32, st);
    return launch_knn3_nearest<HSP_DIST_NEAREST>(query, cand, B, M, N, idx64, idx32, st);
  }
  if (K <= 32) {
    if (formula == HSP_DIST_NEIGHBOR)
      return dispatch_knn3_reg<1, HSP_DIST_NEIGHBOR>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
    return dispatch_knn3_reg<1, HSP_DIST_NEAREST>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
  }
  // K in (32, 64]: two sorted registers per lane
  if (formula == HSP_DIST_NEIGHBOR)
    return dispatch_knn3_reg<2, HSP_DIST_NEIGHBOR>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
  return dispatch_knn3_reg<2, HSP_DIST_NEAREST>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
}
