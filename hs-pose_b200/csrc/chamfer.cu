// K7 — Chamfer distance forward/backward.
//
// sm_100a counterpart of the reference's (unused, JIT-built) extension
// tools/pyTorchChamferDistance/chamfer_distance.cu:6-187: for every point of
// `a` the squared distance to / index of its nearest point in `b` (and the
// reverse direction in the same launch sequence), and the gradient
//   d dist_a[i] / d a_i = 2 (a_i - b_nn(i)),   d dist_a[i] / d b_nn(i) = -2 (a_i - b_nn(i)).
// Candidates are staged through shared memory in float4 tiles; one thread owns
// one query; first minimum wins on ties (ascending j), as in the reference loop.
#include "common.cuh"

namespace hsp {

constexpr int CH_THREADS = 256;
constexpr int CH_TILE = 1024;

__global__ void __launch_bounds__(CH_THREADS)
chamfer_nn_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int M,
                  float* __restrict__ dist, int32_t* __restrict__ idx) {
  __shared__ float4 s_b[CH_TILE];
  const int o = blockIdx.y;
  const int i = blockIdx.x * CH_THREADS + threadIdx.x;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (i < N) {
    const float* p = a + ((size_t)o * N + i) * 3;
    px = p[0]; py = p[1]; pz = p[2];
  }
  float best = INFINITY;
  int bj = 0;
  for (int j0 = 0; j0 < M; j0 += CH_TILE) {
    const int cnt = min(CH_TILE, M - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += CH_THREADS) {
      const float* q = b + ((size_t)o * M + j0 + j) * 3;
      s_b[j] = make_float4(q[0], q[1], q[2], 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 q = s_b[j];
      const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < best) { best = d; bj = j0 + j; }
    }
  }
  if (i < N) {
    dist[(size_t)o * N + i] = best;
    idx[(size_t)o * N + i] = bj;
  }
}

// ga[i] = 2 g[i] (a_i - b_nn) ; gb[nn] -= the same (atomic, as the reference).
__global__ void chamfer_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const int32_t* __restrict__ idx, const float* __restrict__ g,
                                   int N, int M, int total, float* __restrict__ ga,
                                   float* __restrict__ gb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int o = t / N;
  const int j = idx[t];
  const float* p = a + (size_t)t * 3;
  const float* q = b + ((size_t)o * M + j) * 3;
  const float w = 2.0f * g[t];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float v = w * (p[d] - q[d]);
    atomicAdd(ga + (size_t)t * 3 + d, v);
    atomicAdd(gb + ((size_t)o * M + j) * 3 + d, -v);
  }
}

// Deterministic form (the default).  CTA = object (x a slice of its points): the inverse of the idx_b table — for
// every a_i the b_j whose nearest neighbour it is, ascending j — is built in shared memory (count, scan, fill,
// per-point insertion sort of the short lists), then thread = point a_i adds its own term and those of its list in
// that fixed order and writes ga once: no float atomics, no memset, O(N + M) work per object.
constexpr int CHB_THREADS = 512;
constexpr int CHB_LONG = 24;      // lists longer than this are summed by the whole CTA
constexpr int CHB_HEAVY = 512;    // >= M / CHB_LONG for every M the shared-memory tables admit
__global__ void __launch_bounds__(CHB_THREADS)
chamfer_bwd_det_kernel(const float* __restrict__ a, const float* __restrict__ b, const int32_t* __restrict__ idx_a,
                       const int32_t* __restrict__ idx_b, const float* __restrict__ g_a,
                       const float* __restrict__ g_b, int N, int M, int per, float* __restrict__ ga) {
  extern __shared__ int32_t s_mem[];
  int32_t* s_start = s_mem;                    // [N + 1]
  int32_t* s_cur = s_mem + N + 1;              // [N]
  int32_t* s_list = s_cur + N;                 // [M]
  const int o = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int32_t* ib = idx_b + (size_t)o * M;
  for (int v = tid; v < N; v += CHB_THREADS) s_cur[v] = 0;
  __syncthreads();
  for (int e = tid; e < M; e += CHB_THREADS) {
    const int k = __ldg(ib + e);
    if (k >= 0 && k < N) atomicAdd(&s_cur[k], 1);
  }
  __syncthreads();
  if (tid < 32) {                              // exclusive scan of the counts, 32 at a time
    int run = 0;
    for (int v0 = 0; v0 < N; v0 += 32) {
      const int v = v0 + lane;
      const int c = v < N ? s_cur[v] : 0;
      int inc = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (v < N) s_start[v] = run + inc - c;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_start[N] = run;
  }
  __syncthreads();
  for (int v = tid; v < N; v += CHB_THREADS) s_cur[v] = s_start[v];
  __syncthreads();
  for (int e = tid; e < M; e += CHB_THREADS) {
    const int k = __ldg(ib + e);
    if (k >= 0 && k < N) s_list[atomicAdd(&s_cur[k], 1)] = e;
  }
  __syncthreads();
  __shared__ int s_heavy[CHB_HEAVY];
  __shared__ int s_nheavy;
  __shared__ float s_red[3][CHB_THREADS / 32];
  if (tid == 0) s_nheavy = 0;
  __syncthreads();
  const int i_end = min(N, (int)(blockIdx.x + 1) * per);
  for (int i = blockIdx.x * per + tid; i < i_end; i += CHB_THREADS) {
    const int s0 = s_start[i], s1 = s_start[i + 1];
    if (s1 - s0 > CHB_LONG) {                  // a point many b_j collapse onto: summed by the whole CTA below
      const int slot = atomicAdd(&s_nheavy, 1);
      if (slot < CHB_HEAVY) s_heavy[slot] = i;
      continue;
    }
    for (int x = s0 + 1; x < s1; ++x) {        // the fill order is arbitrary: sort the (short) list
      const int key = s_list[x];
      int y = x - 1;
      while (y >= s0 && s_list[y] > key) { s_list[y + 1] = s_list[y]; --y; }
      s_list[y + 1] = key;
    }
    const size_t t = (size_t)o * N + i;
    const float p[3] = {a[t * 3], a[t * 3 + 1], a[t * 3 + 2]};
    const float* q = b + ((size_t)o * M + idx_a[t]) * 3;
    const float w = 2.0f * g_a[t];
    float acc[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[d] = w * (p[d] - q[d]);
    for (int x = s0; x < s1; ++x) {
      const int j = s_list[x];
      const float* bj = b + ((size_t)o * M + j) * 3;
      const float wj = 2.0f * __ldg(g_b + (size_t)o * M + j);
#pragma unroll
      for (int d = 0; d < 3; ++d) acc[d] -= wj * (__ldg(bj + d) - p[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) ga[t * 3 + d] = acc[d];
  }
  __syncthreads();
  // heavy points: thread t owns the contiguous slice [t * chunk, (t + 1) * chunk) of the b_j (ascending j inside),
  // lanes are combined by a shuffle tree and the warps in order 0..15 — a fixed order again
  const int nheavy = min(s_nheavy, CHB_HEAVY);
  const int chunk = (M + CHB_THREADS - 1) / CHB_THREADS;
  for (int h = 0; h < nheavy; ++h) {
    const int i = s_heavy[h];
    const size_t t = (size_t)o * N + i;
    const float p[3] = {a[t * 3], a[t * 3 + 1], a[t * 3 + 2]};
    float acc[3] = {0.f, 0.f, 0.f};
    const int j1 = min(M, (tid + 1) * chunk);
    for (int j = tid * chunk; j < j1; ++j) {
      if (__ldg(ib + j) == i) {
        const float* bj = b + ((size_t)o * M + j) * 3;
        const float wj = 2.0f * __ldg(g_b + (size_t)o * M + j);
#pragma unroll
        for (int d = 0; d < 3; ++d) acc[d] -= wj * (__ldg(bj + d) - p[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc[d] += __shfl_down_sync(0xffffffffu, acc[d], sft);
      if (lane == 0) s_red[d][tid >> 5] = acc[d];
    }
    __syncthreads();
    if (tid < 3) {
      const float* q = b + ((size_t)o * M + idx_a[t]) * 3;
      float r = 2.0f * g_a[t] * (p[tid] - q[tid]);
      for (int w = 0; w < CHB_THREADS / 32; ++w) r += s_red[tid][w];
      ga[t * 3 + tid] = r;
    }
    __syncthreads();
  }
}

}  // namespace hsp

extern "C" int hsp_chamfer_fwd(const float* a, const float* b, int B, int N, int M,
                               float* dist_a, int32_t* idx_a, float* dist_b, int32_t* idx_b,
                               void* stream) {
  using namespace hsp;
  if (!a || !b || !dist_a || !idx_a || !dist_b || !idx_b || B < 0 || N <= 0 || M <= 0 ||
      B > 65535)
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  chamfer_nn_kernel<<<dim3((N + CH_THREADS - 1) / CH_THREADS, B), CH_THREADS, 0, st>>>(
      a, b, N, M, dist_a, idx_a);
  HSP_LAUNCH_CHECK();
  chamfer_nn_kernel<<<dim3((M + CH_THREADS - 1) / CH_THREADS, B), CH_THREADS, 0, st>>>(
      b, a, M, N, dist_b, idx_b);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_chamfer_bwd(const float* a, const float* b, const int32_t* idx_a,
                               const int32_t* idx_b, const float* gdist_a,
                               const float* gdist_b, int B, int N, int M, float* ga, float* gb,
                               void* stream) {
  using namespace hsp;
  if (!a || !b || !idx_a || !idx_b || !gdist_a || !gdist_b || !ga || !gb || B < 0 || N <= 0 ||
      M <= 0)
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sm_a = ((size_t)2 * N + 1 + M) * sizeof(int32_t), sm_b = ((size_t)2 * M + 1 + N) * sizeof(int32_t);
  if (sm_a <= 48 * 1024 && sm_b <= 48 * 1024 && B <= 65535) {
    const int gxa = (N + 2 * CHB_THREADS - 1) / (2 * CHB_THREADS), gxb = (M + 2 * CHB_THREADS - 1) / (2 * CHB_THREADS);
    chamfer_bwd_det_kernel<<<dim3(gxa, B), CHB_THREADS, sm_a, st>>>(a, b, idx_a, idx_b, gdist_a, gdist_b, N, M,
                                                                   (N + gxa - 1) / gxa, ga);
    HSP_LAUNCH_CHECK();
    chamfer_bwd_det_kernel<<<dim3(gxb, B), CHB_THREADS, sm_b, st>>>(b, a, idx_b, idx_a, gdist_b, gdist_a, M, N,
                                                                   (M + gxb - 1) / gxb, gb);
    HSP_LAUNCH_CHECK();
    return HSP_OK;
  }
  // very large clouds: float atomics (order-dependent in the last bit)
  if (cudaMemsetAsync(ga, 0, sizeof(float) * (size_t)B * N * 3, st) != cudaSuccess ||
      cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)B * M * 3, st) != cudaSuccess)
    return HSP_ELAUNCH;
  chamfer_bwd_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(a, b, idx_a, gdist_a, N, M, B * N, ga, gb);
  HSP_LAUNCH_CHECK();
  chamfer_bwd_kernel<<<(B * M + 255) / 256, 256, 0, st>>>(b, a, idx_b, gdist_b, M, N, B * M, gb, ga);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
