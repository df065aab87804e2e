// K11: input pre-stage — depth ROI -> camera-frame cloud -> n_pts sampled points, on the device.
//
// Replaces, per object of a batch,
//   * datasets/load_data.py:322-333 `_depth_to_pcl` (numpy, float64 arithmetic) + :277 `/ 1000.0`
//     and :307-320 `_sample_points` (tile when short, random subset when long), and
//   * network/point_sample/pc_sample.py:8-77 `PC_sample` (torch, float32 arithmetic, the `depth=` path of
//     network/HSPose.py:40-48).
// Both keep the pixels with depth > 0 and mask > 0 IN RASTER ORDER, back-project them with the pinhole
// intrinsics and pick n_pts of them.  CTA = object; the compaction is an order-preserving block scan over
// 1024-pixel chunks, so the compacted cloud is bit-identical to the reference's boolean-mask indexing.
// Sampling: with caller-drawn indices (the reference's numpy draw -> exact parity) it is a gather; without,
// the device rule = tile (i % total, exactly load_data.py:316-317) or, when total > n_pts, a uniformly random
// subset WITHOUT replacement (the n_pts smallest of counter-hashed 32-bit keys found by a 4-pass radix select;
// same distribution as permutation(total)[:n_pts], emitted in raster order; not numpy's stream).
#include "common.cuh"

namespace hsp {

constexpr int PS_THREADS = 1024;

// exclusive rank of `flag` among the block's threads (thread order) and the block total
__device__ __forceinline__ int block_rank(bool flag, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_warp[w] = __popc(m);
  __syncthreads();
  if (w == 0) {
    const int v = s_warp[lane];
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    s_warp[lane] = inc - v;
    if (lane == 31) s_warp[32] = inc;
  }
  __syncthreads();
  const int r = s_warp[w] + __popc(m & ((1u << lane) - 1u));
  total = s_warp[32];
  __syncthreads();   // s_warp is reused by the next call
  return r;
}

template <bool FP64>
__global__ void __launch_bounds__(PS_THREADS)
depth_to_cloud_kernel(const float* __restrict__ depth, const float* __restrict__ mask,
                      const float* __restrict__ xymap, const void* __restrict__ camK, int HW,
                      float* __restrict__ cloud, int* __restrict__ count) {
  __shared__ int s_warp[33];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* dp = depth + (size_t)b * HW;
  const float* mk = mask + (size_t)b * HW;
  const float* xm = xymap + (size_t)b * 2 * HW;
  const float* ym = xm + HW;
  float* out = cloud + (size_t)b * HW * 3;
  double fx64 = 0, fy64 = 0, cx64 = 0, cy64 = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  if (FP64) {
    const double* K = (const double*)camK + 9 * b;
    fx64 = K[0]; cx64 = K[2]; fy64 = K[4]; cy64 = K[5];
  } else {
    const float* K = (const float*)camK + 9 * b;
    fx = K[0]; cx = K[2]; fy = K[4]; cy = K[5];
  }
  int base_out = 0;
  for (int base = 0; base < HW; base += PS_THREADS) {
    const int i = base + tid;
    float d = 0.f;
    bool valid = false;
    if (i < HW) {
      d = dp[i];
      valid = (d > 0.f) && (mk[i] > 0.f);
    }
    int total;
    const int r = block_rank(valid, s_warp, total);
    if (valid) {
      float x, y;
      if (FP64) {   // numpy: ((x_map - cx) * depth / fx) in float64, then astype(float32)
        x = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)xm[i], cx64), (double)d), fx64);
        y = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)ym[i], cy64), (double)d), fy64);
      } else {      // torch float32: (x - ux) * dp / fx
        x = __fdiv_rn(__fmul_rn(__fsub_rn(xm[i], cx), d), fx);
        y = __fdiv_rn(__fmul_rn(__fsub_rn(ym[i], cy), d), fy);
      }
      float* o = out + (size_t)(base_out + r) * 3;
      o[0] = __fdiv_rn(x, 1000.0f);   // millimetres -> metres, float32 true division (numpy / torch-CPU)
      o[1] = __fdiv_rn(y, 1000.0f);
      o[2] = __fdiv_rn(d, 1000.0f);
    }
    base_out += total;
  }
  if (tid == 0) count[b] = base_out;
}

__device__ __forceinline__ uint32_t ps_key(uint32_t seed_lo, uint32_t seed_hi, uint32_t b, uint32_t i) {
  uint32_t h = seed_lo ^ (b * 0x9E3779B9u);
  h ^= i * 0x85EBCA6Bu + seed_hi;
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  h += i; h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}

__global__ void __launch_bounds__(PS_THREADS)
sample_points_kernel(const float* __restrict__ cloud, const int* __restrict__ count,
                     const int* __restrict__ choose, uint32_t seed_lo, uint32_t seed_hi, int cap, int n_pts,
                     float* __restrict__ out, int* __restrict__ status) {
  __shared__ int s_warp[33];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_remaining;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = cloud + (size_t)b * cap * 3;
  float* dst = out + (size_t)b * n_pts * 3;
  const int total = min(count[b], cap);
  if (total <= 0) {           // nothing valid: zeros, flagged (the reference loader skips such a sample)
    for (int i = tid; i < n_pts * 3; i += PS_THREADS) dst[i] = 0.f;
    if (tid == 0 && status) atomicOr(status, 1);
    return;
  }
  if (choose) {               // the caller's draw (numpy permutation / choice on the host): plain gather
    const int* ch = choose + (size_t)b * n_pts;
    for (int i = tid; i < n_pts; i += PS_THREADS) {
      int j = ch[i];
      if (j < 0 || j >= total) {
        if (status) atomicOr(status, 2);
        j = min(max(j, 0), total - 1);
      }
      dst[3 * i] = src[3 * j]; dst[3 * i + 1] = src[3 * j + 1]; dst[3 * i + 2] = src[3 * j + 2];
    }
    return;
  }
  if (total <= n_pts) {       // load_data.py:316-317: tile the cloud, then its head (== i mod total)
    for (int i = tid; i < n_pts; i += PS_THREADS) {
      const int j = i % total;
      dst[3 * i] = src[3 * j]; dst[3 * i + 1] = src[3 * j + 1]; dst[3 * i + 2] = src[3 * j + 2];
    }
    return;
  }
  // total > n_pts: the n_pts smallest keys.  Radix select, most significant byte first.
  if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)n_pts; }
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = tid; i < 256; i += PS_THREADS) s_hist[i] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
    for (int i = tid; i < total; i += PS_THREADS) {
      const uint32_t k = ps_key(seed_lo, seed_hi, b, i);
      if ((k & himask) == prefix) atomicAdd(&s_hist[(k >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned rem = s_remaining, acc = 0;
      int d = 0;
      for (; d < 255; ++d) {
        if (acc + s_hist[d] >= rem) break;
        acc += s_hist[d];
      }
      s_prefix = prefix | ((unsigned)d << (8 * pass));
      s_remaining = rem - acc;
    }
    __syncthreads();
  }
  const unsigned T = s_prefix;          // the n_pts-th smallest key
  const int take_eq = (int)s_remaining;  // how many keys == T belong to the subset (first ones in raster order)
  int base_out = 0, eq_seen = 0;
  for (int base = 0; base < total; base += PS_THREADS) {
    const int i = base + tid;
    uint32_t k = 0xFFFFFFFFu;
    const bool in = i < total;
    if (in) k = ps_key(seed_lo, seed_hi, b, i);
    const bool eq = in && k == T;
    int tot_eq;
    const int r_eq = block_rank(eq, s_warp, tot_eq);
    const bool sel = in && (k < T || (eq && eq_seen + r_eq < take_eq));
    int tot_sel;
    const int r = block_rank(sel, s_warp, tot_sel);
    if (sel) {
      const int o = base_out + r;
      dst[3 * o] = src[3 * (size_t)i]; dst[3 * o + 1] = src[3 * (size_t)i + 1]; dst[3 * o + 2] = src[3 * (size_t)i + 2];
    }
    base_out += tot_sel;
    eq_seen += tot_eq;
  }
}

}  // namespace hsp

extern "C" int hsp_depth_to_cloud(const float* depth, const float* mask, const float* xymap, const void* camK,
                                  int camK_is_f64, int B, int H, int W, float* cloud, int* count, void* stream) {
  using namespace hsp;
  if (!depth || !mask || !xymap || !camK || !cloud || !count || B <= 0 || H <= 0 || W <= 0) return HSP_EINVAL;
  if ((long long)H * W > (1ll << 24)) return HSP_EINVAL;
  if (camK_is_f64)
    depth_to_cloud_kernel<true><<<B, PS_THREADS, 0, (cudaStream_t)stream>>>(depth, mask, xymap, camK, H * W, cloud,
                                                                            count);
  else
    depth_to_cloud_kernel<false><<<B, PS_THREADS, 0, (cudaStream_t)stream>>>(depth, mask, xymap, camK, H * W, cloud,
                                                                             count);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_sample_points(const float* cloud, const int* count, const int* choose, unsigned long long seed,
                                 int B, int cap, int n_pts, float* out, int* status, void* stream) {
  using namespace hsp;
  if (!cloud || !count || !out || B <= 0 || cap <= 0 || n_pts <= 0) return HSP_EINVAL;
  sample_points_kernel<<<B, PS_THREADS, 0, (cudaStream_t)stream>>>(cloud, count, choose, (uint32_t)seed,
                                                                   (uint32_t)(seed >> 32), cap, n_pts, out, status);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
