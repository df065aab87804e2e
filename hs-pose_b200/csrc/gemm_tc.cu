// K6 — dense per-point MLP contractions on the 5th-generation tensor cores.
//
// Replaces the library GEMMs behind every Conv1d(kernel_size=1) / `@` of the reference's dense
// stages: the heads' 1286|1289 -> 1024 -> 256 stacks (network/fs_net_repo/PoseR.py:27-34,
// PoseTs.py:32-39), the train-only recon / face stacks (FaceRecon.py:38-68,114-124), the HS
// layers' `feature_map @ self.weights + self.bias` (gcn3d.py:171) and the STE / conv2 1x1
// convolutions (gcn3d.py:85,112,149,186), forward and backward (dgrad, wgrad).
//
//   D[M,N] (+bias[N]) = A[M,K] . B[N,K]^T        bf16 operands, fp32 accumulation in TMEM
//
// Either operand may be "K-major" (the reduction index is the contiguous one in HBM, e.g. an
// activation matrix (rows, channels) as A, an nn.Linear weight (out, in) as B) or "MN-major"
// (the M / N index is contiguous: the transposed reads of the backward GEMMs), so no operand is
// ever transposed in memory:
//      forward  y  = x  . W^T        A = x  K-major,  B = W  (out,in)   K-major
//      HS layer P  = fm . W          A = fm K-major,  B = W  (in,out)   MN-major
//      dgrad    dx = dy . W          A = dy K-major,  B = W  (out,in)   MN-major
//      wgrad    dW = dy^T . x        A = dy MN-major, B = x             MN-major   (split-K, fp32 out)
//
// Structure (one persistent CTA pair per two SMs, warp-specialised, no __syncthreads in the loop):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (128-byte swizzle) into a STAGES-deep ring,
//            completion counted on mbarriers (`full`), slots released by tcgen05.commit (`empty`).
//   warp 1   MMA issuer (leader CTA only): tcgen05.mma.cta_group::2 kind::f16, M = 256 (two CTAs x
//            128 rows) x N = BN x K = 16, accumulators in TMEM, two accumulator buffers so the
//            MMAs of tile i+1 overlap the epilogue of tile i.
//   warps 2-5 epilogue: tcgen05.ld (thread = accumulator row) -> + bias -> bf16 / fp32 ->
//            swizzled shared staging -> TMA store; optionally per-tile column sums and sums of
//            squares of the stored values (BatchNorm batch statistics: the stand-alone reduction
//            pass over the activation disappears), written as per-row-block partials that
//            bn.cu's finalize adds in a fixed order (deterministic, no float atomics).
// With CTAS = 1 the same code runs single-CTA (M = 128); used for small problems.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace hsp {
namespace gemm {

constexpr int BM = 128;          // accumulator rows per CTA (TMEM lanes)
constexpr int BK = 64;           // reduction elements per pipeline stage (128 bytes of bf16)
constexpr int UK = 16;           // reduction elements per tcgen05.mma (kind::f16)
constexpr int THREADS = 192;     // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int EPI_THREADS = 128;
constexpr int OUT_STAGE_BYTES = BM * 128;   // staging tile: 128 rows x 128 bytes (64 bf16 / 32 fp32)
constexpr int SMEM_LIMIT = 227 * 1024;

template <int BN, int CTAS>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = BN / CTAS;               // rows of the B tile this CTA stages
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PART_BYTES = 4 * BN * 2 * (int)sizeof(float);   // [row group][{sum,sq}][BN]
  static constexpr int FIXED = 2 * OUT_STAGE_BYTES + PART_BYTES + 2 * BN * (int)sizeof(float) + 512 /*barriers*/ +
                               1024 /*alignment slack*/;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = FIXED + STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;        // power of two: BN in {32,64,128,256}
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  // default (.release.cta) semantics: a cluster-scope release would drain this thread's in-flight TMA loads
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// TMA: 2-D tiled tensor load global -> shared, bytes counted on an mbarrier.  With two CTAs per MMA the
// barrier is the LEADER's (cluster address with the CTA-rank bit cleared), the data lands in the own CTA.
template <int CTAS>
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  if (CTAS == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (layout type 2), descriptor version 1.
//   K-major  : rows of 128 bytes (64 bf16 of K), 8-row groups SBO = 1024 bytes apart; LBO unused.
//   MN-major : blocks of [64 k-rows][64 mn elements = 128 bytes]; 8-k groups SBO = 1024 bytes apart,
//              64-element MN blocks LBO = 64 rows * 128 bytes = 8192 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((mn_major ? (BK * 128) >> 4 : 1) & 0x3fff) << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor kind::f16: D = F32, A = B = BF16, majors per operand, M x N.
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CTAS>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (CTAS == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when every MMA issued so far by this thread has retired (both CTAs' copies of the barrier).
template <int CTAS>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (CTAS == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Params {
  int M, N, K;          // problem: D[M,N] = A[M,K] . B[N,K]^T
  int a_mn, b_mn;       // operand majors (0 = K-major, 1 = MN-major)
  int splits;           // split-K factor (partials in planes of the 3-D output map)
  int kb_per_split;     // 64-element k-blocks per split
  const float* bias;    // (N) or NULL
  const float* bias_rows;   // (ceil(M / rows_per_group), N) or NULL: a second bias, shared by groups of consecutive rows
  int rows_per_group;
  float* stats;         // (ceil(M/128), 2, N) per-row-block column sums / sums of squares, or NULL
  int relu;             // epilogue: max(0, .) after the biases
  const float* c_in;    // (M, ldc) fp32 or NULL: residual added to the product (fp32 output, splits == 1) — a dgrad
  int ldc;              // whose result joins an existing gradient needs no separate add pass
  int debug;            // diagnostics (hsp_gemm_debug): 1 = no staging writes / stores, 2 = no MMAs, 4 = no TMA loads
};

// ---------------------------------------------------------------- the kernel
template <int BN, int CTAS, bool OUT_F32, bool EXTRA>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const Params p) {
  using C = Cfg<BN, CTAS>;
  constexpr int STAGES = C::STAGES;
  constexpr int OUT_COLS = OUT_F32 ? 32 : 64;            // columns per staging tile (128 bytes)
  constexpr int CHUNKS = BN / OUT_COLS;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                                  // [STAGES][A_BYTES]
  unsigned char* sB = sA + STAGES * C::A_BYTES;              // [STAGES][B_BYTES]
  unsigned char* sO = sB + STAGES * C::B_BYTES;              // [2][OUT_STAGE_BYTES]
  float* s_part = reinterpret_cast<float*>(sO + 2 * OUT_STAGE_BYTES);   // [4][2][BN]
  float* s_bias = s_part + 4 * 2 * BN;                       // [2][BN] (double-buffered across tiles)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 2 * BN);
  uint64_t* full = bars;                                     // [STAGES]
  uint64_t* empty = bars + STAGES;                           // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;                       // [2] accumulator ready
  uint64_t* tempty = bars + 2 * STAGES + 2;                  // [2] accumulator drained
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CTAS == 1 ? 0u : cluster_rank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full + i, CTAS);      // one arrive per CTA's producer (+ the transaction bytes)
      mbar_init(empty + i, 1);        // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull + i, 1);
      mbar_init(tempty + i, 4 * CTAS);   // one arrive per epilogue warp of every CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                   "r"(C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                   "r"(C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // ---- work decomposition (identical in every role): item w -> (row block, column block, k split)
  const int tiles_m = (p.M + BM * CTAS - 1) / (BM * CTAS);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int total_kb = (p.K + BK - 1) / BK;
  // Work item w -> (split, tile): SPLIT-MAJOR, so the CTA pairs that run at the same time work on (nearly) all
  // output tiles of ONE k-slice and share its A / B slabs through L2 — every operand byte leaves HBM about once.
  // (split fastest re-read the B operand of the 3584 x 1296 x 131584 wgrad ~9 times: 4.8 GB of DRAM traffic for
  // 1.4 GB of algorithmic bytes, ncu r2s.)
  const int n_tiles_mn = tiles_m * tiles_n;
  const int n_work = n_tiles_mn * p.splits;
  const int n_clusters = gridDim.x / CTAS;
  const int cluster_id = blockIdx.x / CTAS;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int split = w / n_tiles_mn, t = w % n_tiles_mn;
        const int tn = t % tiles_n, tm = t / tiles_n;
        const int m0 = (tm * CTAS + (int)rank) * BM;
        const int n0 = tn * BN + (int)rank * C::B_ROWS;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          if (p.debug & 4) {
            if (CTAS == 1 || leader) mbar_arrive(full + stage); else mbar_arrive_cluster(full + stage, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (CTAS == 1) mbar_expect_tx(full + stage, C::STAGE_BYTES);
          unsigned char* a = sA + stage * C::A_BYTES;
          unsigned char* b = sB + stage * C::B_BYTES;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d<CTAS>(a, &tmA, k0, m0, full + stage);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d<CTAS>(a + i * (BK * 128), &tmA, m0 + i * 64, k0, full + stage);
          }
          if (!p.b_mn) {
            tma_load_2d<CTAS>(b, &tmB, k0, n0, full + stage);
          } else {
#pragma unroll
            for (int i = 0; i < C::B_ROWS / 64; ++i)
              tma_load_2d<CTAS>(b + i * (BK * 128), &tmB, n0 + i * 64, k0, full + stage);
          }
          if (CTAS == 2) {
            if (leader) mbar_expect_tx(full + stage, 2 * C::STAGE_BYTES);
            else mbar_arrive_cluster(full + stage, 0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc(BM * CTAS, BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t a_adv = p.a_mn ? (UK * 128) >> 4 : (UK * 2) >> 4;   // descriptor start-address step per K = 16
      const uint32_t b_adv = p.b_mn ? (UK * 128) >> 4 : (UK * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters, ++it) {
        const int split = w / n_tiles_mn;
        const int kb0 = split * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int buf = it & 1;
        mbar_wait(tempty + buf, ((it >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * C::A_BYTES), p.a_mn != 0);
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * C::B_BYTES), p.b_mn != 0);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            if (!(p.debug & 2)) umma_bf16<CTAS>(d_tmem, ad + (uint64_t)(k * a_adv), bd + (uint64_t)(k * b_adv), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit<CTAS>(empty + stage);                    // slot free once these MMAs retire
          if (kb == kb1 - 1) umma_commit<CTAS>(tfull + buf);   // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue ===================================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;                // accumulator row of this thread
    const int etid = tid - 64;                    // 0..127
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool do_stats = p.stats != nullptr && !OUT_F32;
    constexpr int NV = OUT_COLS;                  // accumulator values per thread per chunk
    auto load_bias = [&](int w, int slot) {       // bias slice of work item w -> s_bias[slot]
      const int n0w = ((w % n_tiles_mn) % tiles_n) * BN;
      for (int c = etid; c < BN; c += EPI_THREADS)
        s_bias[slot * BN + c] = (p.bias && n0w + c < p.N) ? __ldg(p.bias + n0w + c) : 0.f;
    };
    auto load_acc = [&](uint32_t taddr, uint32_t* dst) {
      tmem_ld32(taddr, *reinterpret_cast<uint32_t (*)[32]>(dst));
      if (NV == 64) tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t (*)[32]>(dst + 32));
    };
    if (cluster_id < n_work) load_bias(cluster_id, 0);
    int it = 0;
    uint32_t ost = 0;                             // staging-buffer counter
    for (int w = cluster_id; w < n_work; w += n_clusters, ++it) {
      const int split = w / n_tiles_mn, t = w % n_tiles_mn;
      const int tn = t % tiles_n, tm = t / tiles_n;
      const int mblk = tm * CTAS + (int)rank;
      const int m0 = mblk * BM, n0 = tn * BN;
      const int buf = it & 1;
      epi_barrier();                              // s_bias[it & 1] written; previous tile's s_bias / s_part readers done
      if (w + n_clusters < n_work) load_bias(w + n_clusters, (it + 1) & 1);
      const float* bs_tile = s_bias + (it & 1) * BN;
      mbar_wait(tfull + buf, (it >> 1) & 1);
      tc_fence_after();
      const bool full_rows = m0 + BM <= p.M;      // warp-uniform: only the last row block masks rows
      const bool row_ok = m0 + row < p.M;
      uint32_t v[2][NV];
      load_acc(t_lane + (uint32_t)(buf * BN), v[0]);
#pragma unroll
      for (int ch = 0; ch < CHUNKS; ++ch, ++ost) {
        unsigned char* so = sO + (ost & 1) * OUT_STAGE_BYTES;
        uint32_t* vc = v[ch & 1];
        tmem_ld_wait();                           // chunk ch is in registers
        if (ch + 1 < CHUNKS) {
          load_acc(t_lane + (uint32_t)(buf * BN + (ch + 1) * OUT_COLS), v[(ch + 1) & 1]);   // in flight during the conversion
        } else {                                  // every TMEM read of this accumulator has completed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 1) mbar_arrive(tempty + buf); else mbar_arrive_cluster(tempty + buf, 0);
          }
        }
        if (p.debug & 1) continue;
        if (!full_rows && !row_ok) {
#pragma unroll
          for (int e = 0; e < NV; ++e) vc[e] = 0u;
        }
        // + bias, convert, write this thread's row into the 128-byte-swizzled staging tile
        const float4* bs4 = reinterpret_cast<const float4*>(bs_tile + ch * OUT_COLS);
        const bool add_bias = p.bias != nullptr && (full_rows || row_ok);
        // EXTRA instantiations only (kept out of the plain kernel: with four epilogue warps per SM every
        // instruction of this loop is on the critical path of the output-bound shapes).  Per-row-group bias
        // (e.g. a per-object term broadcast over the object's points) for this thread's row, then ReLU.
        const float* brow = nullptr;
        if (EXTRA && p.bias_rows != nullptr && (full_rows || row_ok))
          brow = p.bias_rows + (size_t)((m0 + row) / p.rows_per_group) * p.N + n0 + ch * OUT_COLS;
        const bool brow_vec = n0 + (ch + 1) * OUT_COLS <= p.N;     // whole chunk inside the matrix: vector loads
        const float* crow = nullptr;                               // residual row segment of this thread (fp32 output)
        if (EXTRA && OUT_F32 && p.c_in != nullptr && (full_rows || row_ok))
          crow = p.c_in + (size_t)(m0 + row) * p.ldc + n0 + ch * OUT_COLS;
        const float lo = (EXTRA && p.relu) ? 0.f : -INFINITY;
        const uint32_t srow = smem_u32(so) + row * 128;
        if (OUT_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {           // 8 x 16 bytes = 32 floats
            float4 o = make_float4(__uint_as_float(vc[4 * j]), __uint_as_float(vc[4 * j + 1]),
                                   __uint_as_float(vc[4 * j + 2]), __uint_as_float(vc[4 * j + 3]));
            if (add_bias) {
              const float4 b4 = bs4[j];
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            if (EXTRA) {
              if (brow && (brow_vec || n0 + ch * OUT_COLS + 4 * j < p.N)) {      // N % 8 == 0: groups are all in or out
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(brow) + j);
                o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
              }
              if (crow && n0 + ch * OUT_COLS + 4 * j < p.N) {                    // N % 4 == 0 (checked by the host)
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(crow) + j);
                o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w;
              }
              o.x = fmaxf(o.x, lo); o.y = fmaxf(o.y, lo); o.z = fmaxf(o.z, lo); o.w = fmaxf(o.w, lo);
            }
            st_shared_v4(srow + ((j ^ (row & 7)) << 4), __float_as_uint(o.x), __float_as_uint(o.y),
                         __float_as_uint(o.z), __float_as_uint(o.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {           // 8 x 16 bytes = 64 bf16
            float2 f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              f[e] = make_float2(__uint_as_float(vc[8 * j + 2 * e]), __uint_as_float(vc[8 * j + 2 * e + 1]));
            if (add_bias) {
              const float4 b0 = bs4[2 * j], b1 = bs4[2 * j + 1];
              f[0] = __fadd2_rn(f[0], make_float2(b0.x, b0.y));
              f[1] = __fadd2_rn(f[1], make_float2(b0.z, b0.w));
              f[2] = __fadd2_rn(f[2], make_float2(b1.x, b1.y));
              f[3] = __fadd2_rn(f[3], make_float2(b1.z, b1.w));
            }
            if (EXTRA) {
              if (brow && (brow_vec || n0 + ch * OUT_COLS + 8 * j < p.N)) {      // N % 8 == 0: groups are all in or out
                const float4 r0 = __ldg(reinterpret_cast<const float4*>(brow) + 2 * j);
                const float4 r1 = __ldg(reinterpret_cast<const float4*>(brow) + 2 * j + 1);
                f[0] = __fadd2_rn(f[0], make_float2(r0.x, r0.y));
                f[1] = __fadd2_rn(f[1], make_float2(r0.z, r0.w));
                f[2] = __fadd2_rn(f[2], make_float2(r1.x, r1.y));
                f[3] = __fadd2_rn(f[3], make_float2(r1.z, r1.w));
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) { f[e].x = fmaxf(f[e].x, lo); f[e].y = fmaxf(f[e].y, lo); }
            }
            st_shared_v4(srow + ((j ^ (row & 7)) << 4), pack_bf16(f[0].x, f[0].y), pack_bf16(f[1].x, f[1].y),
                         pack_bf16(f[2].x, f[2].y), pack_bf16(f[3].x, f[3].y));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the TMA engine
        // every earlier store has finished READING shared memory, in particular the one issued one chunk
        // ago from the other staging buffer, which the next chunk overwrites after this barrier
        if (etid == 0) tma_store_wait_read<0>();
        epi_barrier();
        if (etid == 0) tma_store_3d(&tmO, so, n0 + ch * OUT_COLS, m0, split);
        if (do_stats) {
          // column sums / sums of squares of the STORED bf16 values: lane = column pair, over the warp's own 32 rows
          float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
          const uint32_t so_u32 = smem_u32(so);
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const int rr = q * 32 + r;
            const uint32_t wv = ld_shared_u32(so_u32 + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
            const float2 f2 = make_float2(__uint_as_float(wv << 16), __uint_as_float(wv & 0xffff0000u));
            s2 = __fadd2_rn(s2, f2);
            q2 = __ffma2_rn(f2, f2, q2);
          }
          float* pp = s_part + (q * 2) * BN + ch * OUT_COLS + 2 * lane;
          *reinterpret_cast<float2*>(pp) = s2;
          *reinterpret_cast<float2*>(pp + BN) = q2;
        }
      }
      if (do_stats) {
        epi_barrier();
        for (int c = etid; c < BN; c += EPI_THREADS) {
          if (n0 + c < p.N && m0 < p.M) {      // (the last CTA pair may own a row block entirely past M)
            float s = 0.f, sq = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) { s += s_part[(g * 2) * BN + c]; sq += s_part[(g * 2 + 1) * BN + c]; }
            float* dst = p.stats + ((size_t)mblk * 2) * p.N + n0 + c;
            dst[0] = s;
            dst[p.N] = sq;
          }
        }
      }
    }
    if (etid == 0) tma_store_wait_read<0>();
  }

  // ---- teardown
  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync_all();
  if (warp == 1) {
    if (CTAS == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// Row-major matrix (rows, cols) with leading dimension ld (elements); box = (box_cols, box_rows); planes of
// `rows * ld` elements in a third dimension (split-K partials) when planes > 1.
static bool make_map(CUtensorMap* m, const void* base, bool f32, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_cols, uint32_t box_rows, uint64_t planes) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const uint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {cols, rows, planes};
  cuuint64_t gstr[2] = {ld * es, rows * ld * es};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t rank = planes > 0 ? 3 : 2;
  return enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base),
             gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int CTAS, bool OUT_F32, bool EXTRA>
static int launch(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tO, const Params& p, int sms,
                  cudaStream_t st) {
  using C = Cfg<BN, CTAS>;
  auto kern = gemm_tc_kernel<BN, CTAS, OUT_F32, EXTRA>;
  static bool configured = false;      // per instantiation
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return HSP_ELAUNCH;
    configured = true;
  }
  const int tiles_m = (p.M + BM * CTAS - 1) / (BM * CTAS), tiles_n = (p.N + BN - 1) / BN;
  const int n_work = tiles_m * tiles_n * p.splits;
  int clusters = sms / CTAS;
  if (clusters > n_work) clusters = n_work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CTAS);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, tA, tB, tO, p) != cudaSuccess) return HSP_ELAUNCH;
  return HSP_OK;
}

}  // namespace gemm
}  // namespace hsp

static int g_gemm_debug = 0;
extern "C" int hsp_gemm_debug(int flags) {
  const int old = g_gemm_debug;
  g_gemm_debug = flags;
  return old;
}

extern "C" int hsp_gemm_bf16_splits(int M, int N, int K, int out_f32) {
  // Split the reduction when the output has too few tiles to fill the 74 CTA pairs, or when the tile count
  // leaves the last wave mostly empty (wgrad: short, very deep problems).  Cost model: waves x k-blocks per
  // work item (+ a per-item epilogue overhead); the smallest cost wins, ties go to fewer splits.
  using namespace hsp::gemm;
  if (!out_f32) return 1;
  const long tiles = (long)((M + 255) / 256) * ((N + 255) / 256);
  const long total_kb = (K + BK - 1) / BK;
  long best = -1;
  int best_s = 1;
  for (int s = 1; s <= 64; ++s) {
    const long kb_per = (total_kb + s - 1) / s;
    if (s > 1 && ((long)(s - 1) * kb_per >= total_kb || kb_per < 8)) continue;   // every split owns >= 8 k-blocks
    const long waves = (tiles * s + 73) / 74;
    const long cost = waves * (kb_per + 12);
    if (best < 0 || cost < best) { best = cost; best_s = s; }
  }
  return best_s;
}

extern "C" int hsp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M,
                             int N, int K, const float* bias, const float* bias_rows, int rows_per_group, int relu,
                             void* out, int ldo, int out_f32, int splits, float* stats, int tile_n, int ctas,
                             void* stream) {
  return hsp_gemm_bf16_acc(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, bias, bias_rows, rows_per_group, relu,
                           nullptr, 0, out, ldo, out_f32, splits, stats, tile_n, ctas, stream);
}

extern "C" int hsp_gemm_bf16_acc(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M,
                                 int N, int K, const float* bias, const float* bias_rows, int rows_per_group,
                                 int relu, const float* c_in, int ldc, void* out, int ldo, int out_f32, int splits,
                                 float* stats, int tile_n, int ctas, void* stream) {
  using namespace hsp;
  using namespace hsp::gemm;
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0 || splits < 1) return HSP_EINVAL;
  if (c_in && (!out_f32 || splits != 1 || (N % 4) != 0 || ldc < N || (ldc % 4) != 0 || ((uintptr_t)c_in % 16) != 0))
    return HSP_EINVAL;
  const int oes = out_f32 ? 4 : 2;
  if (((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)out % 16) || (lda * 2) % 16 || (ldb * 2) % 16 ||
      (ldo * oes) % 16)
    return HSP_EINVAL;
  if (lda < (a_mn_major ? M : K) || ldb < (b_mn_major ? N : K) || ldo < N) return HSP_EINVAL;
  if (stats && (out_f32 || splits != 1)) return HSP_EINVAL;
  if (splits > 1 && (!out_f32 || bias || bias_rows || relu)) return HSP_EINVAL;
  if (bias_rows && (rows_per_group <= 0 || (N % 8) != 0 || ((uintptr_t)bias_rows % 16) != 0)) return HSP_EINVAL;
  const int total_kb = (K + BK - 1) / BK;
  const int kb_per = (total_kb + splits - 1) / splits;
  if ((splits - 1) * kb_per >= total_kb && splits > 1) return HSP_EINVAL;
  // tile selection: two CTAs per MMA and the widest N tile unless the problem is small
  if (ctas != 1 && ctas != 2) ctas = M > 128 ? 2 : 1;
  if (tile_n != 64 && tile_n != 128 && tile_n != 256) tile_n = N > 128 ? 256 : (N > 64 ? 128 : 64);
  if (ctas == 2 && tile_n == 64 && b_mn_major) tile_n = 128;   // an MN-major B box is 64 columns: >= 64 per CTA
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  CUtensorMap tA, tB, tO;
  const int b_rows = tile_n / ctas;
  bool ok = true;
  // A: K-major = (M rows, K cols), box (64 k, 128 m); MN-major = (K rows, M cols), box (64 m, 64 k)
  ok &= a_mn_major ? make_map(&tA, A, false, K, M, lda, 64, BK, 0) : make_map(&tA, A, false, M, K, lda, BK, BM, 0);
  ok &= b_mn_major ? make_map(&tB, B, false, K, N, ldb, 64, BK, 0) : make_map(&tB, B, false, N, K, ldb, BK, b_rows, 0);
  ok &= make_map(&tO, out, out_f32 != 0, M, N, ldo, out_f32 ? 32 : 64, BM, splits);
  if (!ok) return HSP_ELAUNCH;

  Params p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn_major ? 1 : 0; p.b_mn = b_mn_major ? 1 : 0;
  p.splits = splits; p.kb_per_split = kb_per;
  p.bias = bias; p.stats = stats;
  p.bias_rows = bias_rows; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.relu = relu ? 1 : 0;
  p.c_in = c_in; p.ldc = ldc;
  p.debug = g_gemm_debug;
  cudaStream_t st = (cudaStream_t)stream;
  const bool extra = bias_rows != nullptr || relu != 0 || c_in != nullptr;   // the epilogue variant with the row-group bias / ReLU / residual code
#define HSP_GEMM_CASE(BN_, CT_)                                                          \
  if (tile_n == BN_ && ctas == CT_)                                                      \
    return out_f32 ? (extra ? launch<BN_, CT_, true, true>(tA, tB, tO, p, sms, st)      \
                            : launch<BN_, CT_, true, false>(tA, tB, tO, p, sms, st))    \
                   : (extra ? launch<BN_, CT_, false, true>(tA, tB, tO, p, sms, st)     \
                            : launch<BN_, CT_, false, false>(tA, tB, tO, p, sms, st))
  HSP_GEMM_CASE(256, 2);
  HSP_GEMM_CASE(128, 2);
  HSP_GEMM_CASE(64, 2);
  HSP_GEMM_CASE(256, 1);
  HSP_GEMM_CASE(128, 1);
  HSP_GEMM_CASE(64, 1);
#undef HSP_GEMM_CASE
  return HSP_EINVAL;
}
