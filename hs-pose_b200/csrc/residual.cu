// K5d — the residual sum that closes every HS layer, fused:
//   out[b,i,c] = feature[b,i,c] + lin[b,i,c] + gproj[b,c] + ste[b,i,c]
// = ORL_forward's `conv2(cat[feature, f_global]) + feature` (reference gcn3d.py:109-113,
// :183-187; the cat is split algebraically: lin = feature @ W2[:, :C]^T, gproj = G @ W2[:, C:]^T
// is a per-object constant) followed by `+ f_STE` (gcn3d.py:90, :156).  The reference (and plain
// PyTorch) runs this as three full-size element-wise passes forward and four backward; here it is
// one pass each way.  lin / ste may be fp32 or bf16 (tensor-core GEMM outputs under autocast).
// Backward: d feature = d out (no copy), d lin = d ste = cast(d out), d gproj = column sums of
// d out per object (fixed-order in-CTA reduction, deterministic).
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TL, typename TS>
__global__ void __launch_bounds__(256)
residual_sum_fwd_kernel(const float* __restrict__ feature, const TL* __restrict__ lin,
                        const float* __restrict__ gproj, const TS* __restrict__ ste,
                        const float* __restrict__ xyz, const float* __restrict__ wxyz, int N, int C,
                        size_t total4, float* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of the (B*N, C) matrix
  if (t >= total4) return;
  const size_t e = t * 4;
  const int c = (int)(e % C);
  const size_t row = e / C;
  const size_t b = row / N;
  float4 v = ld4(feature + e);
  if (lin) { const float4 l = ld4(lin + e); v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w; }
  if (gproj) { const float4 g = ld4(gproj + b * C + c); v.x += g.x; v.y += g.y; v.z += g.z; v.w += g.w; }
  if (ste) { const float4 s = ld4(ste + e); v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
  if (xyz) {   // STE of the surface layer: a 3 -> C linear on the coordinates, evaluated in place
    const float px = __ldg(xyz + 3 * row), py = __ldg(xyz + 3 * row + 1), pz = __ldg(xyz + 3 * row + 2);
    const float4 w0 = ld4(wxyz + 3 * c), w1 = ld4(wxyz + 3 * c + 4), w2 = ld4(wxyz + 3 * c + 8);
    v.x += fmaf(pz, w0.z, fmaf(py, w0.y, px * w0.x));
    v.y += fmaf(pz, w1.y, fmaf(py, w1.x, px * w0.w));
    v.z += fmaf(pz, w2.x, fmaf(py, w1.w, px * w1.z));
    v.w += fmaf(pz, w2.w, fmaf(py, w2.z, px * w2.y));
  }
  *reinterpret_cast<float4*>(out + e) = v;
}

// CTA = (object, 32-channel slab): 8 channel quads x 32 row lanes.
__global__ void __launch_bounds__(256)
residual_sum_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xyz, int N, int C,
                        __nv_bfloat16* __restrict__ g16, float* __restrict__ colsum,
                        float* __restrict__ gw_part) {
  __shared__ float4 sh[32][8];
  const int b = blockIdx.y, q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 32 + 4 * q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ax = acc, ay = acc, az = acc;          // sum_i g[i,c] * xyz[i,d]: gradient of the 3 -> C weight
  if (c < C) {
    const size_t base = (size_t)b * N * C + c;
#pragma unroll 4
    for (int i = rl; i < N; i += 32) {
      const float4 v = ld4(g + base + (size_t)i * C);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (xyz) {
        const float* p = xyz + 3 * ((size_t)b * N + i);
        const float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
        ax.x = fmaf(v.x, px, ax.x); ax.y = fmaf(v.y, px, ax.y); ax.z = fmaf(v.z, px, ax.z); ax.w = fmaf(v.w, px, ax.w);
        ay.x = fmaf(v.x, py, ay.x); ay.y = fmaf(v.y, py, ay.y); ay.z = fmaf(v.z, py, ay.z); ay.w = fmaf(v.w, py, ay.w);
        az.x = fmaf(v.x, pz, az.x); az.y = fmaf(v.y, pz, az.y); az.z = fmaf(v.z, pz, az.z); az.w = fmaf(v.w, pz, az.w);
      }
      if (g16) {
        uint2 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
        h[0] = __floats2bfloat162_rn(v.x, v.y);
        h[1] = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(g16 + base + (size_t)i * C) = u;
      }
    }
  }
  auto reduce = [&](const float4 val) {          // fixed-order sum over the 32 row lanes (result in lane 0)
    __syncthreads();
    sh[rl][q] = val;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl == 0) {
#pragma unroll
      for (int r = 0; r < 32; ++r) { const float4 v = sh[r][q]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    }
    return t;
  };
  const float4 t = reduce(acc);
  if (colsum && rl == 0 && c < C) *reinterpret_cast<float4*>(colsum + (size_t)b * C + c) = t;
  if (gw_part) {                                 // per-object partial (B, C, 3); the caller sums over B
    const float4 tx = reduce(ax), ty = reduce(ay), tz = reduce(az);
    if (rl == 0 && c < C) {
      float* o = gw_part + ((size_t)b * C + c) * 3;
      o[0] = tx.x; o[1] = ty.x; o[2] = tz.x;  o[3] = tx.y; o[4] = ty.y; o[5] = tz.y;
      o[6] = tx.z; o[7] = ty.z; o[8] = tz.z;  o[9] = tx.w; o[10] = ty.w; o[11] = tz.w;
    }
  }
}

}  // namespace hsp

extern "C" int hsp_residual_sum_fwd(const float* feature, const void* lin, int lin_dtype,
                                    const float* gproj, const void* ste, int ste_dtype,
                                    const float* xyz, const float* wxyz, int B, int N, int C,
                                    float* out, void* stream) {
  using namespace hsp;
  if (!feature || !out || B < 0 || N <= 0 || C <= 0 || (C % 4) != 0) return HSP_EINVAL;
  if ((xyz != nullptr) != (wxyz != nullptr)) return HSP_EINVAL;
  if ((lin_dtype != HSP_DTYPE_F32 && lin_dtype != HSP_DTYPE_BF16) ||
      (ste_dtype != HSP_DTYPE_F32 && ste_dtype != HSP_DTYPE_BF16))
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  const size_t total4 = (size_t)B * N * C / 4;
  const unsigned blocks = (unsigned)((total4 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
#define HSP_RS(TL_, TS_)                                                                       \
  residual_sum_fwd_kernel<TL_, TS_><<<blocks, 256, 0, st>>>(feature, (const TL_*)lin, gproj, \
                                                            (const TS_*)ste, xyz, wxyz, N, C, total4, out)
  if (lin_dtype == HSP_DTYPE_BF16) {
    if (ste_dtype == HSP_DTYPE_BF16) HSP_RS(__nv_bfloat16, __nv_bfloat16); else HSP_RS(__nv_bfloat16, float);
  } else {
    if (ste_dtype == HSP_DTYPE_BF16) HSP_RS(float, __nv_bfloat16); else HSP_RS(float, float);
  }
#undef HSP_RS
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_residual_sum_bwd(const float* g, const float* xyz, int B, int N, int C, void* g_bf16,
                                    float* g_gproj, float* g_wxyz_partial, void* stream) {
  using namespace hsp;
  if (!g || (!g_bf16 && !g_gproj && !g_wxyz_partial) || B < 0 || N <= 0 || C <= 0 || (C % 4) != 0 ||
      B > 65535 || ((g_wxyz_partial != nullptr) && !xyz))
    return HSP_EINVAL;
  if (B == 0) return HSP_OK;
  residual_sum_bwd_kernel<<<dim3((C + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(
      g, g_wxyz_partial ? xyz : nullptr, N, C, (__nv_bfloat16*)g_bf16, g_gproj, g_wxyz_partial);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
