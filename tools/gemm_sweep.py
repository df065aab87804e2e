"""K6 timing matrix: tile / CTA-pair / debug-flag variants of hsp_gemm_bf16 on the B=128 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hspose_b200.ops as ops
from hspose_b200 import _lib

dev = torch.device("cuda")
lib = _lib.load()


def t(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def case(M, N, K, a_mn, b_mn, f32, splits):
    A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).to(torch.bfloat16)
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).to(torch.bfloat16)
    odt = torch.float32 if f32 else torch.bfloat16
    out = torch.empty((splits, M, N) if splits > 1 else (M, N), dtype=odt, device=dev)
    print(f"--- M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} f32={f32} splits={splits}", flush=True)
    for ctas in CTAS:
        for tile_n in TILES:
            for stats in ((False, True) if not f32 and splits == 1 else (False,)):
                for dbg in DBG:
                    lib.hsp_gemm_debug(dbg)
                    try:
                        ms = t(lambda: ops.gemm_bf16(A, B, a_mn, b_mn, out=out, out_dtype=odt, splits=splits,
                                                     stats=stats, tile_n=tile_n, ctas=ctas))
                    finally:
                        lib.hsp_gemm_debug(0)
                    print(f"ctas={ctas} tile_n={tile_n} stats={int(stats)} debug={dbg}: {ms:.4f} ms  "
                          f"{2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
CTAS = [int(x) for x in os.environ.get("CTAS", "1,2").split(",")]
TILES = [int(x) for x in os.environ.get("TILES", "256,128").split(",")]
DBG = [int(x) for x in os.environ.get("DBG", "0,1,2,4,3,6,7").split(",")]
if which in ("all", "fwd"):
    case(131584, 1024, 1296, False, False, False, 1)
if which in ("all", "p"):
    case(131584, 1024, 128, False, True, False, 1)
if which in ("all", "wgrad"):
    case(1024, 1296, 131584, True, True, True, 3)
