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
  if (cudaMemsetAsync(ga, 0, sizeof(float) * (size_t)B * N * 3, st) != cudaSuccess ||
      cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)B * M * 3, st) != cudaSuccess)
    return HSP_ELAUNCH;
  chamfer_bwd_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(a, b, idx_a, gdist_a, N, M, B * N, ga, gb);
  HSP_LAUNCH_CHECK();
  chamfer_bwd_kernel<<<(B * M + 255) / 256, 256, 0, st>>>(b, a, idx_b, gdist_b, M, N, B * M, gb, ga);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
