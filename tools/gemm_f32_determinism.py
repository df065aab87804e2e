"""Is the fp32-output K6 GEMM bit-reproducible and correct on the dgrad shapes of the step?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hspose_b200.ops as ops

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K, b_mn) in [(4112, 128, 1024, False), (4112, 128, 128, True), (131584, 128, 1024, False),
                        (131584, 128, 128, True), (1028, 256, 2048, False), (4, 128, 128, True), (4112, 1296, 3584, True)]:
    A = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    Bm = (torch.randn(K, N, device=dev, generator=g) if b_mn else torch.randn(N, K, device=dev, generator=g)).to(torch.bfloat16)
    ref = A.float() @ (Bm.float() if b_mn else Bm.float().t())
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        first = None
        nd = 0
        for r in range(12):
            junk = torch.randn(1 << 22, device=dev)          # move the allocator around between runs
            o = ops.gemm_bf16(A, Bm, b_mn=b_mn, out_dtype=dt)
            torch.cuda.synchronize()
            if first is None:
                first = o.clone()
            elif not torch.equal(first, o):
                nd += 1
            del junk
        err = (first.float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"M={M} N={N} K={K} b_mn={int(b_mn)} out={str(dt)[6:]:9s} non-identical reruns: {nd}/11  max rel err {err:.2e}")
