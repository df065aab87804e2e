// K2 — fused D-dimensional ("RF-F", feature-space) distance + top-k.
//
// Replaces get_neighbor_index(feature_map, k) (reference gcn3d.py:15-24, called
// from get_receptive_fields gcn3d.py:189-209 with D = 128 / 256).  The NxN
// distance matrix is produced tile by tile in shared memory and consumed
// immediately by the warp top-k; it never reaches HBM.
//
// One CTA = one object x TQ query rows.  The query block stays resident in
// shared memory for the whole kernel; candidate rows stream through in
// (TC x DK) chunks.  Each thread owns an MQ x MC register tile of inner
// products and accumulates every one of them as a single sequential FP32 FMA
// chain over d = 0..D-1 (the order oracle/hsp_oracle.c:inner() restates), so
// indices are bit-reproducible.  |f|^2 comes from a small pre-kernel (rounded
// squares added left to right, as for D = 3).  After the last chunk the
// distance tile ((-2*inner) + q_j) + q_i is written to shared memory and each
// warp updates the sorted top-K lists of its rows: chunks with no candidate
// under the row's K-th distance cost one ballot; a few candidates are inserted
// by warp-cooperative shifting; many are bitonic-sorted and merged.
#include <stdlib.h>

#include "common.cuh"

namespace hsp {

constexpr int KF_THREADS = 256;
constexpr int KF_MQ = 4, KF_MC = 4;          // register micro-tile
constexpr int KF_TQ = 16 * KF_MQ;            // query rows per CTA
constexpr int KF_TC = 16 * KF_MC;            // candidate rows per tile
constexpr int KF_DK = 32;                    // feature chunk
constexpr int KF_LDB = KF_DK + 4;            // padded chunk row (conflict-free LDS.128)
constexpr int KF_LDD = KF_TC + 1;

__global__ void sqnorm_rows_kernel(const float* __restrict__ feat, int rows, int D,
                                   float* __restrict__ q) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(feat + (size_t)r * D);
  float s = 0.0f;
  for (int d4 = 0; d4 < D / 4; ++d4) {
    float4 v = __ldg(p + d4);
    if (d4 == 0) s = __fmul_rn(v.x, v.x); else s = __fadd_rn(s, __fmul_rn(v.x, v.x));
    s = __fadd_rn(s, __fmul_rn(v.y, v.y));
    s = __fadd_rn(s, __fmul_rn(v.z, v.z));
    s = __fadd_rn(s, __fmul_rn(v.w, v.w));
  }
  q[r] = s;
}

// Insert one key into the ascending per-lane list (keys are unique).
__device__ __forceinline__ void warp_insert(uint64_t& L, uint64_t key, int lane) {
  const int pos = __popc(__ballot_sync(0xffffffffu, L < key));
  const uint64_t up = __shfl_up_sync(0xffffffffu, L, 1);
  if (lane == pos) L = key;
  else if (lane > pos) L = up;
}

// Orderable key -> the float it encodes (inverse of float_orderable).
__device__ __forceinline__ float key_to_float(uint64_t key) {
  const uint32_t o = (uint32_t)(key >> 32);
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// v2: the inner products run on the packed-FP32 datapath.  A thread owns 4 query rows x 4
// PAIRS of adjacent candidates; the candidate chunk is staged TRANSPOSED ([d][j]) so one
// LDS.64 yields (B[j][d], B[j+1][d]) and one FFMA2 advances two accumulators:
//   acc2[r][c] = fma2((a,a), (b_j, b_j+1), acc2[r][c])
// Every accumulator is still ONE sequential fp32 FMA chain over d = 0..D-1 (FFMA2 is two
// independent IEEE fmas) — bit-identical to the scalar kernel and the oracle.
constexpr int KF2_TC = 128;                 // candidates per tile (16 tx x 4 pairs x 2)
constexpr int KF2_LDBT = KF2_TC + 2;        // transposed chunk pitch (8-byte aligned rows)
constexpr int KF2_LDD = KF2_TC + 1;
constexpr int KF2_PF = KF2_TC * (KF_DK / 4) / KF_THREADS;   // float4 per thread per chunk

template <int NL>
__global__ void __launch_bounds__(KF_THREADS, 2)
knn_feat_kernel(const float* __restrict__ feat, const float* __restrict__ qn, int N, int D, int K,
                int drop, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int LDA = D + 4;
  float* s_A = reinterpret_cast<float*>(smem_raw);                  // [TQ][LDA]      row-major queries
  float* s_BT = s_A + KF_TQ * LDA;                                  // [DK][LDBT]     transposed chunk
  float* s_D = s_BT + KF_DK * KF2_LDBT;                             // [TQ][LDD]      distance tile
  uint64_t* s_L = reinterpret_cast<uint64_t*>(s_D + KF_TQ * KF2_LDD + ((KF_TQ * KF2_LDD) & 1));
  const int b = blockIdx.y, q0 = blockIdx.x * KF_TQ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const float* fb = feat + (size_t)b * N * D;
  const float* qb = qn + (size_t)b * N;

  // resident query block (rows past N are zero-filled and never written out)
  for (int e = tid; e < KF_TQ * (D / 4); e += KF_THREADS) {
    const int r = e / (D / 4), d4 = e % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < N) v = __ldg(reinterpret_cast<const float4*>(fb + (size_t)(q0 + r) * D) + d4);
    *reinterpret_cast<float4*>(s_A + r * LDA + 4 * d4) = v;
  }
  for (int e = tid; e < KF_TQ * NL * 32; e += KF_THREADS) s_L[e] = KEY_MAX;
  float qi[KF_MQ];
#pragma unroll
  for (int r = 0; r < KF_MQ; ++r) {
    const int i = q0 + ty + 16 * r;
    qi[r] = i < N ? __ldg(qb + i) : 0.0f;
  }

  float4 pf[KF2_PF];   // register prefetch of the next transposed chunk (128 rows x 32 features)
#pragma unroll
  for (int u = 0; u < KF2_PF; ++u) {
    const int e = tid + u * KF_THREADS;
    const int r = e / (KF_DK / 4), d4 = e % (KF_DK / 4);
    pf[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < N) pf[u] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)r * D) + d4);
  }
  for (int j0 = 0; j0 < N; j0 += KF2_TC) {
    float2 acc[KF_MQ][4];
#pragma unroll
    for (int r = 0; r < KF_MQ; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.0f, 0.0f);

    for (int d0 = 0; d0 < D; d0 += KF_DK) {
      __syncthreads();  // previous chunk / previous tile's selection done with s_BT / s_D
      // stage the chunk transposed from the registers prefetched during the previous chunk
#pragma unroll
      for (int u = 0; u < KF2_PF; ++u) {
        const int e = tid + u * KF_THREADS;
        const int r = e / (KF_DK / 4), d4 = e % (KF_DK / 4);
        float* dst = s_BT + (4 * d4) * KF2_LDBT + r;
        dst[0] = pf[u].x; dst[KF2_LDBT] = pf[u].y; dst[2 * KF2_LDBT] = pf[u].z; dst[3 * KF2_LDBT] = pf[u].w;
      }
      __syncthreads();
      {   // prefetch the next chunk (next d0, or chunk 0 of the next tile): in flight during the FMAs
        int nj0 = j0, nd0 = d0 + KF_DK;
        if (nd0 >= D) { nd0 = 0; nj0 = j0 + KF2_TC; }
#pragma unroll
        for (int u = 0; u < KF2_PF; ++u) {
          const int e = tid + u * KF_THREADS;
          const int r = e / (KF_DK / 4), d4 = e % (KF_DK / 4);
          pf[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nj0 + r < N)
            pf[u] = __ldg(reinterpret_cast<const float4*>(fb + (size_t)(nj0 + r) * D + nd0) + d4);
        }
      }
#pragma unroll
      for (int d4 = 0; d4 < KF_DK / 4; ++d4) {
        float4 a[KF_MQ];
#pragma unroll
        for (int r = 0; r < KF_MQ; ++r)
          a[r] = *reinterpret_cast<const float4*>(s_A + (ty + 16 * r) * LDA + d0 + 4 * d4);
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          float2 bb[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            bb[c] = *reinterpret_cast<const float2*>(s_BT + (4 * d4 + dd) * KF2_LDBT + 2 * tx + 32 * c);
#pragma unroll
          for (int r = 0; r < KF_MQ; ++r) {
            const float av = dd == 0 ? a[r].x : dd == 1 ? a[r].y : dd == 2 ? a[r].z : a[r].w;
            const float2 a2 = make_float2(av, av);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = __ffma2_rn(a2, bb[c], acc[r][c]);
          }
        }
      }
    }
    // distance tile: ((-2*inner) + q_j) + q_i        (gcn3d.py:21)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int jl = 2 * tx + 32 * c;
      const int j = j0 + jl;
      const float qj0 = j < N ? __ldg(qb + j) : 0.0f;
      const float qj1 = j + 1 < N ? __ldg(qb + j + 1) : 0.0f;
#pragma unroll
      for (int r = 0; r < KF_MQ; ++r) {
        float* dst = s_D + (ty + 16 * r) * KF2_LDD + jl;
        dst[0] = __fadd_rn(__fadd_rn(__fmul_rn(acc[r][c].x, -2.0f), qj0), qi[r]);
        dst[1] = __fadd_rn(__fadd_rn(__fmul_rn(acc[r][c].y, -2.0f), qj1), qi[r]);
      }
    }
    __syncthreads();
    // selection: warp w owns rows w, w+8, ...
    for (int r = warp; r < KF_TQ; r += KF_THREADS / 32) {
      if (q0 + r >= N) break;
      uint64_t L[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) L[l] = s_L[(r * NL + l) * 32 + lane];
      uint64_t thr = (NL == 2 && K > 32) ? shfl_u64(L[NL - 1], K - 33) : shfl_u64(L[0], K - 1);
      float thr_f = key_to_float(thr);   // KEY_MAX decodes to NaN: !(d > NaN) lets everything through
      bool dirty = false;
#pragma unroll
      for (int c = 0; c < KF2_TC / 32; ++c) {
        const int j = j0 + c * 32 + lane;
        const float dv = s_D[r * KF2_LDD + c * 32 + lane];
        // cheap float pre-filter; the exact (distance, index) order is decided on the keys below
        if (__ballot_sync(0xffffffffu, j < N && !(dv > thr_f)) == 0) continue;
        uint64_t key = KEY_MAX;
        if (j < N) key = make_key(dv, (uint32_t)j);
        const bool pass = key < thr;
        unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m == 0) continue;
        dirty = true;
        if (NL == 1 && __popc(m) <= 6) {
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            warp_insert(L[0], shfl_u64(key, src), lane);
          }
          thr = shfl_u64(L[0], K - 1);
        } else {
          uint64_t cnd = warp_sort32(pass ? key : KEY_MAX, lane);
          uint64_t crev = shfl_u64(cnd, 31 - lane);
          uint64_t lo = umin64(L[0], crev);
          if (NL == 2) {
            uint64_t hi = warp_bitonic_merge32(umax64(L[0], crev), lane);
            uint64_t hrev = shfl_u64(hi, 31 - lane);
            L[NL - 1] = warp_bitonic_merge32(umin64(L[NL - 1], hrev), lane);
          }
          L[0] = warp_bitonic_merge32(lo, lane);
          thr = (NL == 2 && K > 32) ? shfl_u64(L[NL - 1], K - 33) : shfl_u64(L[0], K - 1);
        }
        thr_f = key_to_float(thr);
      }
      if (dirty) {
#pragma unroll
        for (int l = 0; l < NL; ++l) s_L[(r * NL + l) * 32 + lane] = L[l];
      }
    }
  }
  __syncthreads();
  const int k_out = K - drop;
  for (int r = warp; r < KF_TQ; r += KF_THREADS / 32) {
    const int i = q0 + r;
    if (i >= N) break;
    const size_t o = ((size_t)b * N + i) * k_out;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int rank = l * 32 + lane - drop;
      if (rank >= 0 && rank < k_out) {
        const uint32_t j = (uint32_t)(s_L[(r * NL + l) * 32 + lane] & 0xffffffffu);
        if (idx64) idx64[o + rank] = (int64_t)j;
        if (idx32) idx32[o + rank] = (int32_t)j;
      }
    }
  }
}

template <int NL>
static int launch_knn_feat(const float* feat, const float* qn, int B, int N, int D, int K,
                           int drop, int64_t* idx64, int32_t* idx32, cudaStream_t st) {
  size_t fl = (size_t)KF_TQ * (D + 4) + KF_DK * KF2_LDBT + KF_TQ * KF2_LDD;
  fl += fl & 1;  // 8-byte align the lists
  size_t smem = fl * sizeof(float) + (size_t)KF_TQ * NL * 32 * sizeof(uint64_t);
  auto kern = knn_feat_kernel<NL>;
  if (smem > 227 * 1024) return HSP_EINVAL;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess)
    return HSP_ELAUNCH;
  dim3 grid((N + KF_TQ - 1) / KF_TQ, B);
  kern<<<grid, KF_THREADS, smem, st>>>(feat, qn, N, D, K, drop, idx64, idx32);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

}  // namespace hsp

namespace hsp {
size_t knn_feat_tc_workspace_bytes(int B, int N);
bool knn_feat_tc_supported(int N, int D, int K);
int knn_feat_tc_launch(const float* feat, int B, int N, int D, int K, int drop, int64_t* idx64,
                       int32_t* idx32, void* workspace, cudaStream_t st);
}  // namespace hsp

extern "C" size_t hsp_knn_feat_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t exact = (size_t)B * N * sizeof(float);
  const size_t tcw = hsp::knn_feat_tc_workspace_bytes(B, N);   // D = 128 tensor-core filter path
  return exact > tcw ? exact : tcw;
}

extern "C" int hsp_knn_feat(const float* feat, int B, int N, int D, int k, int drop_first,
                            int64_t* idx64, int32_t* idx32, void* workspace,
                            size_t workspace_bytes, void* stream) {
  using namespace hsp;
  if (!feat || (!idx64 && !idx32)) return HSP_EINVAL;
  if (B < 0 || N <= 0 || D <= 0 || (D % KF_DK) != 0 || k <= 0 || drop_first < 0)
    return HSP_EINVAL;
  const int K = k + drop_first;
  if (K > N || K > 64 || B > 65535) return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  if (!workspace || workspace_bytes < hsp_knn_feat_workspace_bytes(B, N)) return HSP_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  // D = 128 / 256, N >= 128: tensor-core filter (tcgen05 + TMEM) + exact FP32 refine, bit-identical results
  // (knn_feat_tc.cu).  HSP_KNN_FEAT_EXACT=1 forces the all-FP32 kernel below (A/B testing).
  static const bool force_exact = getenv("HSP_KNN_FEAT_EXACT") != nullptr;
  if (knn_feat_tc_supported(N, D, K) && !force_exact && (((uintptr_t)feat) & 15) == 0)
    return knn_feat_tc_launch(feat, B, N, D, K, drop_first, idx64, idx32, workspace, st);
  float* qn = (float*)workspace;
  sqnorm_rows_kernel<<<(B * N + 127) / 128, 128, 0, st>>>(feat, B * N, D, qn);
  HSP_LAUNCH_CHECK();
  if (K <= 32) return launch_knn_feat<1>(feat, qn, B, N, D, K, drop_first, idx64, idx32, st);
  return launch_knn_feat<2>(feat, qn, B, N, D, K, drop_first, idx64, idx32, st);
}
