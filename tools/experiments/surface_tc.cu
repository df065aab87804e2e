// EXPERIMENT (not built into the library): K3 forward with theta on the tensor cores.  RESULT: correct (max |diff| 3e-7
// vs the packed-FP32 kernel, arg-max mismatch fraction 1.6e-7) but SLOWER — 1.06 ms vs 0.60 ms at B = 128, N = 1028,
// k = 20, C = 128 (0.31 vs 0.16 at N = 257), profiles/r2_k3_tc_experiment.md.  The packed-FP32 kernel spends 1.5 of its
// 4.5 instructions per element on theta; moving theta to tcgen05 leaves the 3 ALU operations per element of the exact
// max / arg-max (FSETP + FSEL + SEL) — at best 1.5x fewer instructions — and adds the slot hand-shake of a 64-column
// TMEM ring per (channel block, neighbour) item (545 M warp instructions executed, 46 % issue slots, ncu).
// To try it again: copy into hs-pose_b200/csrc/, declare the three functions at the bottom in graph_conv.cu and route
// hsp_surface_conv_fwd to surface_conv_tc_launch with a (C / 8) x 4 KB workspace.
//
// K3-TC — HSlayer_surface.graph_conv forward (reference gcn3d.py:92-107 with :39-59) with theta on the tensor cores.
//
//   out[b,i,c] = 1/S * sum_s max(0, max_n rhat[b,i,n] . dirn[:, s*C + c])
//
// theta is a (points x neighbours) x 3 x (S*C) matrix product.  The packed-FP32 kernel (graph_conv.cu,
// surface_conv_fwd2_kernel) spends 1.5 of its 4.5 instructions per element on it and is ALU-pipe bound (ncu r1f: ALU
// 69 %, 0.61 ms at B = 128, N = 1028 with no gather at all).  Here
//   * every FP32 operand is split into three bf16 parts (x = x1 + x2 + x3, 24 mantissa bits) and the six cross terms
//     >= 2^-16 (r1d1, r1d2, r2d1, r1d3, r2d2, r3d1) x 3 components form ONE K = 32 reduction (18 used): theta differs
//     from the FP32 FMA chain by ~1e-7, far inside the 1e-5 contract;
//   * one CTA = 128 consecutive (object, point) rows (persistent over tiles).  The 8 epilogue warps (thread = point =
//     TMEM lane, two warps per lane quadrant splitting the columns) build the A operand — unit directions of the point's k neighbours, split, in UMMA core-matrix order — for all k
//     neighbours (k x 8 KB of shared memory); the B operand (direction columns, split once per call by
//     surf_dirn_prep_kernel, 64 columns = 8 channels x 7 supports + 8 zero columns per block) streams through a
//     4-stage ring of 4 KB TMA bulk copies;
//   * warp 4 issues tcgen05.mma (M = 128 points, N = 64, K = 32) per (channel block, neighbour) into a ring of
//     eight 64-column TMEM slots; the epilogue warps read a slot with tcgen05.ld and keep, per thread, the running
//     maximum and arg-max of 56 (support, channel) columns in registers: 3 ALU operations per element and nothing
//     else.  After the k neighbours of a block: ReLU, mean over S, 32 bytes of `out` and 7 x 8 arg-max bytes per point.
// Arg-max semantics are the packed-FP32 kernel's: strict >, first winner, 255 when ReLU kills every neighbour.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace hsp {
namespace stc {

using namespace hsp::tc;

constexpr int ST_S = 7;                 // supports (reference default)
constexpr int ST_CB = 8;                // channels per column block
constexpr int ST_COLS = 64;             // MMA N: 56 used columns + 8 zero columns
constexpr int ST_K = 32;                // reduction length (18 used)
constexpr int ST_ROWS = 128;            // points per tile = MMA M
constexpr int ST_EPI_THREADS = 256;     // warps 0-7: A builders + epilogue
constexpr int ST_THREADS = ST_EPI_THREADS + 32;   // warp 8: TMA + MMA
constexpr int ST_SLOTS = 8;             // TMEM accumulator slots of 64 columns
constexpr int ST_BSTAGES = 4;
constexpr int ST_A_BYTES = ST_ROWS * ST_K * 2;      // 8 KB per neighbour slot
constexpr int ST_B_BYTES = ST_COLS * ST_K * 2;      // 4 KB per channel block
constexpr int ST_MAXK = 26;             // k * 8 KB + ring must fit shared memory

__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}

// term t multiplies part ST_TA(t) of rhat with part ST_TB(t) of the direction (0-based parts, largest first):
// (0,0) (0,1) (1,0) (0,2) (1,1) (2,0)
__host__ __device__ constexpr int ST_TA(int t) { return t == 2 || t == 4 ? 1 : (t == 5 ? 2 : 0); }
__host__ __device__ constexpr int ST_TB(int t) { return t == 1 || t == 4 ? 1 : (t == 3 ? 2 : 0); }

// B operand: [block][k-chunk (4)][column (64)][8 k-elements] bf16; column = s * 8 + channel-in-block, k = 3 * term + comp.
__global__ void __launch_bounds__(256)
surf_dirn_prep_kernel(const float* __restrict__ dirn, int C, __nv_bfloat16* __restrict__ bop) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (block, column)
  const int nblk = C / ST_CB;
  if (t >= nblk * ST_COLS) return;
  const int blk = t / ST_COLS, col = t % ST_COLS;
  __align__(16) __nv_bfloat16 kv[ST_K];
#pragma unroll
  for (int i = 0; i < ST_K; ++i) kv[i] = __float2bfloat16_rn(0.0f);
  if (col < ST_S * ST_CB) {
    const int s = col / ST_CB, c = blk * ST_CB + col % ST_CB;
    __nv_bfloat16 part[3][3];
#pragma unroll
    for (int comp = 0; comp < 3; ++comp)
      split3(dirn[(size_t)comp * ST_S * C + s * C + c], part[0][comp], part[1][comp], part[2][comp]);
#pragma unroll
    for (int term = 0; term < 6; ++term)
#pragma unroll
      for (int comp = 0; comp < 3; ++comp) kv[3 * term + comp] = part[ST_TB(term)][comp];
  }
  __nv_bfloat16* dst = bop + (size_t)blk * ST_COLS * ST_K;
#pragma unroll
  for (int kc = 0; kc < ST_K / 8; ++kc)
    *reinterpret_cast<uint4*>(dst + ((size_t)kc * ST_COLS + col) * 8) = *reinterpret_cast<const uint4*>(kv + kc * 8);
}

// asynchronous 32-column TMEM load (no wait): the caller overlaps it with arithmetic and waits with tmem_wait_ld
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool AM>
__global__ void __launch_bounds__(ST_THREADS, 1)
surface_conv_tc_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                       const __nv_bfloat16* __restrict__ bop, int B, int N, int k, int C, float* __restrict__ out,
                       uint8_t* __restrict__ argmax) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                        // [k][4 chunks][128 rows][8] bf16
  unsigned char* sB = sA + (size_t)k * ST_A_BYTES;                 // ring of ST_BSTAGES x 4 KB
  float* s_part = reinterpret_cast<float*>(sB + ST_BSTAGES * ST_B_BYTES);   // [128 rows][8 channels]: supports 4-6
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + ST_ROWS * ST_CB);
  uint64_t* a_ready = bars;                                        // [1]   256 arrivals per tile
  uint64_t* b_full = bars + 1;                                     // [ST_BSTAGES]
  uint64_t* b_empty = b_full + ST_BSTAGES;                         // [ST_BSTAGES]
  uint64_t* slot_full = b_empty + ST_BSTAGES;                      // [ST_SLOTS]
  uint64_t* slot_empty = slot_full + ST_SLOTS;                     // [ST_SLOTS]  8 arrivals (one per epilogue warp)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(slot_empty + ST_SLOTS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long rows_total = (long)B * N;                             // tiles run over the flattened (object, point) rows
  const int n_tiles = (int)((rows_total + ST_ROWS - 1) / ST_ROWS);
  const int nblk = C / ST_CB;
  const int SC = ST_S * C;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(ST_SLOTS * ST_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(a_ready, ST_EPI_THREADS);
    for (int i = 0; i < ST_BSTAGES; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < ST_SLOTS; ++i) { mbar_init(slot_full + i, 1); mbar_init(slot_empty + i, ST_EPI_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  if (warp == ST_EPI_THREADS / 32) {
    // ============================ TMA (B blocks) + MMA issue ============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(ST_ROWS, ST_COLS);
      const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(bop);
      int g = 0;                  // (block, neighbour) items issued so far: slot = g % ST_SLOTS
      int bl = 0, bu = 0;         // B blocks loaded / used so far
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        auto load_b = [&]() {
          const int st = bl % ST_BSTAGES;
          if (bl >= ST_BSTAGES) mbar_wait(b_empty + st, ((bl / ST_BSTAGES) - 1) & 1);
          mbar_expect_tx(b_full + st, ST_B_BYTES);
          bulk_g2s(sB + st * ST_B_BYTES, bsrc + (size_t)(bl % nblk) * ST_B_BYTES, ST_B_BYTES, b_full + st);
          ++bl;
        };
        const int b_end = (it + 1) * nblk;             // blocks needed up to the end of this tile
        while (bl < b_end && bl < bu + ST_BSTAGES) load_b();
        mbar_wait(a_ready, it & 1);                    // the A operand of this tile is in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int cb = 0; cb < nblk; ++cb, ++bu) {
          const int st = bu % ST_BSTAGES;
          mbar_wait(b_full + st, (bu / ST_BSTAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b0 = smem_u32(sB + st * ST_B_BYTES);
          for (int n = 0; n < k; ++n, ++g) {
            const int slot = g % ST_SLOTS;
            if (g >= ST_SLOTS) {
              mbar_wait(slot_empty + slot, ((g / ST_SLOTS) - 1) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t a0 = smem_u32(sA + (size_t)n * ST_A_BYTES);
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * ST_COLS);
#pragma unroll
            for (int ks = 0; ks < ST_K / 16; ++ks)
              umma_bf16(d_tmem, umma_desc(a0 + ks * 2 * (ST_ROWS * 16), ST_ROWS * 16, 128),
                        umma_desc(b0 + ks * 2 * (ST_COLS * 16), ST_COLS * 16, 128), idesc, ks != 0);
            umma_commit(slot_full + slot);
          }
          umma_commit(b_empty + st);
          if (bl < b_end) load_b();
        }
      }
    }
  } else {
    // ===== A builders + epilogue: 8 warps; warps w and w + 4 share TMEM lane quadrant w & 3 (thread = point) and
    // split a slot's columns: half 0 = columns 0-31 (supports 0-3 of the block's 8 channels), half 1 = 32-63
    // (supports 4-6 + 8 zero columns)
    const int quad = warp & 3, half = warp >> 2;
    const int prow = quad * 32 + lane;                             // row of the tile = TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32);
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long row = (long)tile * ST_ROWS + prow;
      const bool row_ok = row < rows_total;
      const int b = row_ok ? (int)(row / N) : 0, i = row_ok ? (int)(row % N) : 0;
      // ---- A operand: this point's unit directions (neighbours half, half + 2, ...), split, [n][k-chunk][row][8]
      {
        const float* xb = xyz + (size_t)b * N * 3;
        const int32_t* ib = idx + ((size_t)b * N + i) * k;
        const float px = xb[3 * i], py = xb[3 * i + 1], pz = xb[3 * i + 2];
        for (int n = half; n < k; n += 2) {
          float r[3] = {0.0f, 0.0f, 0.0f};
          if (row_ok) {
            const int nb = ib[n];
            const float rx = __fsub_rn(xb[3 * nb], px), ry = __fsub_rn(xb[3 * nb + 1], py), rz = __fsub_rn(xb[3 * nb + 2], pz);
            const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
            const float den = fmaxf(nrm, 1e-12f);
            r[0] = __fdiv_rn(rx, den); r[1] = __fdiv_rn(ry, den); r[2] = __fdiv_rn(rz, den);
          }
          __nv_bfloat16 part[3][3];
#pragma unroll
          for (int comp = 0; comp < 3; ++comp) split3(r[comp], part[0][comp], part[1][comp], part[2][comp]);
          __align__(16) __nv_bfloat16 kv[ST_K];
#pragma unroll
          for (int q = 0; q < ST_K; ++q) kv[q] = __float2bfloat16_rn(0.0f);
#pragma unroll
          for (int term = 0; term < 6; ++term)
#pragma unroll
            for (int comp = 0; comp < 3; ++comp) kv[3 * term + comp] = part[ST_TA(term)][comp];
          unsigned char* dst = sA + (size_t)n * ST_A_BYTES + prow * 16;
#pragma unroll
          for (int kc = 0; kc < ST_K / 8; ++kc)
            *reinterpret_cast<uint4*>(dst + kc * (ST_ROWS * 16)) = *reinterpret_cast<const uint4*>(kv + kc * 8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the tensor core
        mbar_arrive(a_ready);
      }
      // ---- epilogue: per channel block, running max / arg-max over the k neighbours of this half's 32 columns.
      // The TMEM load of item g + 1 is in flight while item g is folded in.
      uint32_t va[32], vb[32];                                             // k is even: item parity is static
      auto fetch = [&](int gg, uint32_t (&dst)[32]) {
        const int slot = gg % ST_SLOTS;
        mbar_wait(slot_full + slot, (gg / ST_SLOTS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32_async(t_row + (uint32_t)(slot * ST_COLS), dst);
      };
      const int g_end = g + nblk * k;
      fetch(g, va);
      for (int cb = 0; cb < nblk; ++cb) {
        float acc[32];
        int am[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) { acc[q] = 0.0f; am[q] = 255; }       // max_n relu(x_n) = max(0, max_n x_n)
        auto step = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int n) {
          tmem_wait_ld();                                                  // item g is in registers: its slot is free
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(slot_empty + (g % ST_SLOTS));
          ++g;
          if (g < g_end) fetch(g, nxt);
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float v = __uint_as_float(cur[q]);
            if (AM) { if (v > acc[q]) { acc[q] = v; am[q] = n; } }
            else acc[q] = fmaxf(acc[q], v);
          }
        };
        for (int n = 0; n < k; n += 2) {
          step(va, vb, n);
          step(vb, va, n + 1);
        }
        // sum over this half's supports per channel; half 1 hands its partial sums to half 0 through shared memory
        float o[ST_CB];
#pragma unroll
        for (int c = 0; c < ST_CB; ++c) {
          float sum = 0.0f;
#pragma unroll
          for (int s = 0; s < 4; ++s)
            if (half == 0 || s < 3) sum += acc[s * ST_CB + c];
          o[c] = sum;
        }
        if (half == 1) {
          float4* pp = reinterpret_cast<float4*>(s_part + prow * ST_CB);
          pp[0] = make_float4(o[0], o[1], o[2], o[3]);
          pp[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ST_EPI_THREADS) : "memory");
        if (half == 0 && row_ok) {
          const float4* pp = reinterpret_cast<const float4*>(s_part + prow * ST_CB);
          const float4 p0 = pp[0], p1 = pp[1];
          // same order as the packed-FP32 kernel: supports 0..6 left to right
          float4* op = reinterpret_cast<float4*>(out + (size_t)row * C + cb * ST_CB);
          op[0] = make_float4(__fdiv_rn(o[0] + p0.x, (float)ST_S), __fdiv_rn(o[1] + p0.y, (float)ST_S),
                              __fdiv_rn(o[2] + p0.z, (float)ST_S), __fdiv_rn(o[3] + p0.w, (float)ST_S));
          op[1] = make_float4(__fdiv_rn(o[4] + p1.x, (float)ST_S), __fdiv_rn(o[5] + p1.y, (float)ST_S),
                              __fdiv_rn(o[6] + p1.z, (float)ST_S), __fdiv_rn(o[7] + p1.w, (float)ST_S));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ST_EPI_THREADS) : "memory");   // s_part may be overwritten
        if (AM && row_ok) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (half == 0 || s < 3) {
              uint2 w;
              w.x = (unsigned)am[s * ST_CB] | ((unsigned)am[s * ST_CB + 1] << 8) | ((unsigned)am[s * ST_CB + 2] << 16) |
                    ((unsigned)am[s * ST_CB + 3] << 24);
              w.y = (unsigned)am[s * ST_CB + 4] | ((unsigned)am[s * ST_CB + 5] << 8) | ((unsigned)am[s * ST_CB + 6] << 16) |
                    ((unsigned)am[s * ST_CB + 7] << 24);
              *reinterpret_cast<uint2*>(argmax + (size_t)row * SC + (half * 4 + s) * C + cb * ST_CB) = w;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ST_SLOTS * ST_COLS)
                 : "memory");
}

static size_t st_smem(int k) {
  return (size_t)k * ST_A_BYTES + ST_BSTAGES * ST_B_BYTES + ST_ROWS * ST_CB * 4 + (1 + 2 * ST_BSTAGES + 2 * ST_SLOTS) * 8 + 16;
}

}  // namespace stc

bool surface_conv_tc_supported(int N, int k, int S, int C, const void* out, const void* argmax) {
  return S == stc::ST_S && C % stc::ST_CB == 0 && k >= 2 && k <= stc::ST_MAXK && (k & 1) == 0 && N >= 1 &&
         (((uintptr_t)out) & 15) == 0 && (((uintptr_t)argmax) & 7) == 0;
}
size_t surface_conv_tc_workspace_bytes(int C) { return (size_t)(C / stc::ST_CB) * stc::ST_B_BYTES + 256; }

int surface_conv_tc_launch(const float* xyz, const int32_t* idx, const float* dirn, int B, int N, int k, int C,
                           float* out, uint8_t* argmax, void* workspace, cudaStream_t st) {
  using namespace stc;
  __nv_bfloat16* bop = (__nv_bfloat16*)(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
  const int nblk = C / ST_CB;
  surf_dirn_prep_kernel<<<(nblk * ST_COLS + 255) / 256, 256, 0, st>>>(dirn, C, bop);
  HSP_LAUNCH_CHECK();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (int)(((long)B * N + ST_ROWS - 1) / ST_ROWS);
  const int grid = n_tiles < sms ? n_tiles : sms;
  const size_t smem = st_smem(k);
  if (argmax) {
    if (cudaFuncSetAttribute(surface_conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return HSP_ELAUNCH;
    surface_conv_tc_kernel<true><<<grid, ST_THREADS, smem, st>>>(xyz, idx, bop, B, N, k, C, out, argmax);
  } else {
    if (cudaFuncSetAttribute(surface_conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return HSP_ELAUNCH;
    surface_conv_tc_kernel<false><<<grid, ST_THREADS, smem, st>>>(xyz, idx, bop, B, N, k, C, out, nullptr);
  }
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

}  // namespace hsp
