// Shared device helpers for libhspose_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hspose_b200.h"

#define HSP_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return HSP_ELAUNCH;              \
  } while (0)

namespace hsp {

constexpr uint64_t KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

// Total order on fp32 that matches "<" on floats, with -0 folded onto +0 so
// equal distances tie on the index.  NaN sorts after +inf.
__device__ __forceinline__ uint32_t float_orderable(float d) {
  d = __fadd_rn(d, 0.0f);
  uint32_t u = __float_as_uint(d);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t j) {
  return ((uint64_t)float_orderable(d) << 32) | (uint64_t)j;
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a < b ? b : a; }

// Full bitonic sort of one key per lane, ascending in lane order (15 stages).
__device__ __forceinline__ uint64_t warp_sort32(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      uint64_t other = shfl_xor_u64(key, j);
      bool asc = ((lane & k) == 0);
      bool lower = ((lane & j) == 0);
      key = (asc == lower) ? umin64(key, other) : umax64(key, other);
    }
  }
  return key;
}
// Bitonic merge (5 stages): input bitonic across lanes, output ascending.
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t key, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    uint64_t other = shfl_xor_u64(key, j);
    key = ((lane & j) == 0) ? umin64(key, other) : umax64(key, other);
  }
  return key;
}

// 32-bit variants (orderable distances only) and the K-th smallest of 64 values held two per lane.
__device__ __forceinline__ uint32_t warp_sort32_u32(uint32_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool asc = ((lane & k) == 0), lower = ((lane & j) == 0);
      key = (asc == lower) ? min(key, other) : max(key, other);
    }
  }
  return key;
}
__device__ __forceinline__ uint32_t warp_bitonic_merge32_u32(uint32_t key, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint32_t other = __shfl_xor_sync(0xffffffffu, key, j);
    key = ((lane & j) == 0) ? min(key, other) : max(key, other);
  }
  return key;
}
// K-th smallest (1-based K <= 64) of the 64 values {a, b} x 32 lanes.
__device__ __forceinline__ uint32_t warp_kth_smallest64(uint32_t a, uint32_t b, int K, int lane) {
  a = warp_sort32_u32(a, lane);
  b = warp_sort32_u32(b, lane);
  const uint32_t brev = __shfl_sync(0xffffffffu, b, 31 - lane);
  if (K <= 32) return __shfl_sync(0xffffffffu, warp_bitonic_merge32_u32(min(a, brev), lane), K - 1);
  return __shfl_sync(0xffffffffu, warp_bitonic_merge32_u32(max(a, brev), lane), K - 33);
}

// Running top-K (K <= 32*NL) of a stream of keys, one warp per query.
// L[0] holds ranks 0..31 (lane = rank), L[1] ranks 32..63 when NL == 2.
template <int NL>
struct WarpTopK {
  uint64_t L[NL];
  uint64_t thr;     // current K-th smallest key (warp-uniform)
  int qn;           // entries waiting in the queue (warp-uniform)
  uint64_t* queue;  // 64 slots of shared memory private to the warp

  __device__ __forceinline__ void reset(uint64_t* q) {
#pragma unroll
    for (int i = 0; i < NL; ++i) L[i] = KEY_MAX;
    thr = KEY_MAX;
    qn = 0;
    queue = q;
  }
  // Merge 32 candidate keys (one per lane, any order) into the lists.
  __device__ __forceinline__ void merge(uint64_t c, int lane, int K) {
    c = warp_sort32(c, lane);
    uint64_t crev = shfl_u64(c, 31 - lane);
    uint64_t lo = umin64(L[0], crev);
    if (NL == 2) {
      uint64_t hi = umax64(L[0], crev);
      hi = warp_bitonic_merge32(hi, lane);
      uint64_t hrev = shfl_u64(hi, 31 - lane);
      L[1] = warp_bitonic_merge32(umin64(L[1], hrev), lane);
    }
    L[0] = warp_bitonic_merge32(lo, lane);
    if (NL == 2 && K > 32) thr = shfl_u64(L[1], K - 33);
    else thr = shfl_u64(L[0], K - 1);
  }
  // Offer one key per lane (KEY_MAX for inactive lanes).
  __device__ __forceinline__ void push(uint64_t key, int lane, int K) {
    bool pass = key < thr;
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    if (pass) queue[qn + __popc(m & ((1u << lane) - 1u))] = key;
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
      uint64_t c = queue[lane];
      uint64_t rest = (32 + lane < qn) ? queue[32 + lane] : KEY_MAX;
      __syncwarp();
      queue[lane] = rest;
      qn -= 32;
      __syncwarp();
      merge(c, lane, K);
    }
  }
  __device__ __forceinline__ void finish(int lane, int K) {
    if (qn > 0) {
      uint64_t c = (lane < qn) ? queue[lane] : KEY_MAX;
      qn = 0;
      __syncwarp();
      merge(c, lane, K);
    }
  }
  // rank r (0-based) lives in L[r/32] at lane r%32.
};

}  // namespace hsp
