// fp32 -> bf16 multi-term split for fp32-accurate contractions on the bf16 tensor cores (configs[1]: the fp32
// evaluation forward, 1e-5 contract).  x = x1 + x2 + x3 with x1 = bf16(x), x2 = bf16(x - x1),
// x3 = bf16(x - x1 - x2) (3 x 8 mantissa bits: the residual is below 2^-24 |x|).  A product x.w is then
//   x1 w1 + (x1 w2 + x2 w1) + (x1 w3 + x2 w2 + x3 w1)        (dropped terms <= 2^-24 |x||w|)
// i.e. ONE bf16 GEMM (hsp_gemm_bf16, fp32 accumulation in TMEM) over a 6x longer reduction axis whose operands
// are the concatenations  A' = [x1 x1 x2 x1 x2 x3],  B' = [w1 w2 w1 w3 w2 w1].  This kernel writes such a
// concatenation: term t of the output holds component comp[t] of the input, each term padded to Kpad columns
// (a multiple of 64 = one pipeline k-block, so split-K can give every term its own accumulator plane and the
// large and the small terms never share a rounding).
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {

__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, int ld, int M, int K, int Kpad, int nterms, int comp_packed,
                  __nv_bfloat16* __restrict__ out) {
  const int kq = Kpad / 4;                       // 4 columns per thread
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)M * kq) return;
  const int r = (int)(t / kq), c0 = (int)(t % kq) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 + 3 < K && ((ld & 3) == 0)) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ld + c0));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) if (c0 + i < K) v[i] = x[(size_t)r * ld + c0 + i];
  }
  __nv_bfloat16 parts[3][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 a = __float2bfloat16_rn(v[i]);
    const float r1 = v[i] - __bfloat162float(a);
    const __nv_bfloat16 b = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b);
    parts[0][i] = a; parts[1][i] = b; parts[2][i] = __float2bfloat16_rn(r2);
  }
  const size_t ldo = (size_t)nterms * Kpad;
  for (int tm = 0; tm < nterms; ++tm) {
    const int c = (comp_packed >> (2 * tm)) & 3;
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(parts[c][0]) | ((uint32_t)__bfloat16_as_ushort(parts[c][1]) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(parts[c][2]) | ((uint32_t)__bfloat16_as_ushort(parts[c][3]) << 16);
    *reinterpret_cast<uint2*>(out + (size_t)r * ldo + (size_t)tm * Kpad + c0) = o;
  }
}

}  // namespace hsp

extern "C" int hsp_split_bf16(const float* x, int ld, int M, int K, int Kpad, int nterms, const int* comp /*host*/,
                              void* out, void* stream) {
  using namespace hsp;
  if (!x || !out || !comp || M <= 0 || K <= 0 || Kpad < K || (Kpad % 8) != 0 || nterms < 1 || nterms > 8 || ld < K)
    return HSP_EINVAL;
  int packed = 0;
  for (int t = 0; t < nterms; ++t) {
    if (comp[t] < 0 || comp[t] > 2) return HSP_EINVAL;
    packed |= comp[t] << (2 * t);
  }
  const long threads = (long)M * (Kpad / 4);
  split_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ld, M, K, Kpad, nterms, packed,
                                                                                     (__nv_bfloat16*)out);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
