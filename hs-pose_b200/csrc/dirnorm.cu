// Unit support directions of the 3D-GCN layers: `F.normalize(self.directions, dim=0)` of the reference
// (network/fs_net_repo/gcn3d.py:95, :162) and its autograd, one launch each way.
//   fwd: out[:, j] = d[:, j] / max(||d[:, j]||_2, eps)           d, out (3, n) row-major, nrm (n)
//   bwd: gd[:, j]  = (g[:, j] - out[:, j] <out[:, j], g[:, j]>) / ||d[:, j]||   (||d|| >= eps)
//                  = g[:, j] / eps                                             (clamped columns)
// PyTorch spends ~4 element-wise / reduction launches forward and ~12 backward per layer on these
// 3 x 896..3584 element tensors; inside the captured step that is pure launch latency.
#include "common.cuh"

namespace hsp {

__global__ void __launch_bounds__(256)
normalize_cols_fwd_kernel(const float* __restrict__ d, int n, float eps, float* __restrict__ out,
                          float* __restrict__ nrm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float x = d[j], y = d[n + j], z = d[2 * n + j];
  const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  const float den = fmaxf(len, eps);
  out[j] = __fdiv_rn(x, den);
  out[n + j] = __fdiv_rn(y, den);
  out[2 * n + j] = __fdiv_rn(z, den);
  nrm[j] = len;
}

__global__ void __launch_bounds__(256)
normalize_cols_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out,
                          const float* __restrict__ nrm, int n, float eps, float* __restrict__ gd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float gx = g[j], gy = g[n + j], gz = g[2 * n + j];
  const float len = nrm[j];
  if (len < eps) {   // clamp_min(eps) is constant there: the quotient's gradient only
    gd[j] = gx / eps; gd[n + j] = gy / eps; gd[2 * n + j] = gz / eps;
    return;
  }
  const float ox = out[j], oy = out[n + j], oz = out[2 * n + j];
  const float dot = fmaf(oz, gz, fmaf(oy, gy, ox * gx));
  gd[j] = (gx - ox * dot) / len;
  gd[n + j] = (gy - oy * dot) / len;
  gd[2 * n + j] = (gz - oz * dot) / len;
}

}  // namespace hsp

extern "C" int hsp_normalize_cols_fwd(const float* d, int n, float eps, float* out, float* nrm, void* stream) {
  if (!d || !out || !nrm || n <= 0) return HSP_EINVAL;
  hsp::normalize_cols_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d, n, eps, out, nrm);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_normalize_cols_bwd(const float* g, const float* out, const float* nrm, int n, float eps,
                                      float* gd, void* stream) {
  if (!g || !out || !nrm || !gd || n <= 0) return HSP_EINVAL;
  hsp::normalize_cols_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, out, nrm, n, eps, gd);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
