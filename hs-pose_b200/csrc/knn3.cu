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
  if (K <= 32) {
    if (formula == HSP_DIST_NEIGHBOR)
      return launch_knn3<1, HSP_DIST_NEIGHBOR>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
    return launch_knn3<1, HSP_DIST_NEAREST>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
  }
  if (formula == HSP_DIST_NEIGHBOR)
    return launch_knn3<2, HSP_DIST_NEIGHBOR>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
  return launch_knn3<2, HSP_DIST_NEAREST>(query, cand, B, M, N, K, drop_first, idx64, idx32, st);
}
