"""K6 (hsp_gemm_bf16: tcgen05.mma + TMA + TMEM, csrc/gemm_tc.cu) against a plain fp32 torch matmul of
the SAME bf16-rounded operands, through the C ABI.  Tolerance: the fp32 accumulation order differs from
torch's, and a bf16 output is rounded once more: |err| <= 2^-8 * max|ref| (bf16 out), 1e-4 * max|ref| (fp32 out).
Shapes cover the dense stages of the reference (PoseR.py:27-34, FaceRecon.py:38-68, gcn3d.py:171) incl. the
ragged tails (N = 1028 points per object -> M % 128 != 0, K = 1296 -> K % 64 != 0, 30 / 3 output channels)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, bias):
    r = A.float() @ B.float().t()
    return r + bias if bias is not None else r


def _mk(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) * 0.5).cuda().to(torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    return A, B, bias


@pytest.mark.parametrize("M,N,K,tile_n,ctas", [
    (256, 128, 128, 128, 1), (300, 200, 328, 64, 1), (2056, 1024, 1296, 256, 1),
    (2056, 1024, 1296, 256, 2), (1028, 512, 1296, 128, 2), (1028, 104, 304, 64, 2),
    (4112, 256, 1024, 0, 0), (1028, 32, 128, 0, 0), (130, 8, 256, 0, 0)])
def test_gemm_nt_bias_and_bn_partials(cuda, M, N, K, tile_n, ctas):
    import hspose_b200.ops as ops
    A, B, bias = _mk(M, N, K, 1)
    out, st = ops.gemm_bf16(A, B, bias=bias, stats=True, tile_n=tile_n, ctas=ctas)
    ref = _ref(A, B, bias)
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert (out.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    # BatchNorm partials = column sums / sums of squares of the values AS STORED, per 128-row block
    y = out.float()
    assert st.shape == ((M + 127) // 128, 2, N)
    for blk in (0, st.shape[0] - 1):
        rows = y[blk * 128:(blk + 1) * 128]
        assert torch.allclose(st[blk, 0], rows.sum(0), rtol=1e-4, atol=1e-3)
        assert torch.allclose(st[blk, 1], (rows * rows).sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("ctas", [1, 2])
def test_gemm_transposed_operands(cuda, a_mn, b_mn, ctas):
    """MN-major reads: dgrad (B = W (out,in) read along `in`), the HS layers' (in,out) weights, wgrad."""
    import hspose_b200.ops as ops
    M, N, K = 1032, 520, 1296      # an MN-major operand's row pitch (M or N elements) must be 16-byte aligned
    A, B, bias = _mk(M, N, K, 2)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out = ops.gemm_bf16(a, b, a_mn, b_mn, bias=bias, out_dtype=torch.float32, tile_n=256, ctas=ctas)
    ref = _ref(A, B, bias)
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("splits", [1, 2, 5])
def test_gemm_wgrad_split_k(cuda, splits):
    """dW = dY^T X over M = 8 x 1028 rows: split-K planes summed in order (deterministic)."""
    import hspose_b200.ops as ops
    rows, n_out, n_in = 8224, 1024, 1296
    g = torch.Generator().manual_seed(3)
    dY = (torch.randn(rows, n_out, generator=g) * 0.1).cuda().to(torch.bfloat16)
    X = (torch.randn(rows, n_in, generator=g) * 0.5).cuda().to(torch.bfloat16)
    dW = ops.gemm_bf16(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
    ref = dY.float().t() @ X.float()
    assert dW.shape == (n_out, n_in)
    assert (dW - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    again = ops.gemm_bf16(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
    assert torch.equal(dW, again)          # bit-reproducible


def test_gemm_column_slices_and_errors(cuda):
    """Operands / outputs that are column slices of wider buffers (the concatenated-weights node)."""
    import hspose_b200.ops as ops
    from hspose_b200._lib import HSPoseLibraryError
    g = torch.Generator().manual_seed(4)
    wide = (torch.randn(1028, 3584, generator=g) * 0.3).cuda().to(torch.bfloat16)
    W = (torch.randn(1296, 1024, generator=g) * 0.1).cuda().to(torch.bfloat16)     # (K_out? no: (N=1296, K=1024))
    A = wide[:, 1024:2048]
    out = ops.gemm_bf16(A, W)
    ref = A.float() @ W.float().t()
    assert (out.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    with pytest.raises(HSPoseLibraryError):
        ops.gemm_bf16(A.cpu(), W.cpu())                       # no CPU path
    with pytest.raises(TypeError):
        ops.gemm_bf16(A.float(), W)
    with pytest.raises(ValueError):
        ops.gemm_bf16(A, W[:, :512])


@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K,rpg,ctas", [(2056, 520, 264, 1028, 2), (1028, 1288, 328, 257, 1), (300, 64, 128, 100, 0)])
def test_gemm_row_group_bias_and_relu(cuda, out_dtype, M, N, K, rpg, ctas):
    """The EXTRA epilogue (its own kernel instantiation): per-row-group bias (FaceRecon.py:88-91's per-object
    one-hot / global-feature term broadcast over the object's points) + ReLU, ragged M and N tails."""
    import hspose_b200.ops as ops
    A, B, bias = _mk(M, N, K, 5)
    G = (M + rpg - 1) // rpg
    br = torch.randn(G, N, generator=torch.Generator().manual_seed(6)).cuda()
    ref = _ref(A, B, bias) + br.repeat_interleave(rpg, 0)[:M]
    for relu in (False, True):
        out = ops.gemm_bf16(A, B, bias=bias, bias_rows=br, rows_per_group=rpg, relu=relu, out_dtype=out_dtype, ctas=ctas)
        r = ref.clamp_min(0) if relu else ref
        tol = (2 ** -8 if out_dtype == torch.bfloat16 else 1e-4) * ref.abs().max().item()
        assert (out.float() - r).abs().max().item() <= tol
    if out_dtype == torch.bfloat16:      # BatchNorm partials see the biased values
        out, st = ops.gemm_bf16(A, B, bias=bias, bias_rows=br, rows_per_group=rpg, stats=True, ctas=ctas)
        assert torch.allclose(st[0, 0], out.float()[:128].sum(0), rtol=1e-4, atol=1e-3)
    plain = ops.gemm_bf16(A, B, relu=True, out_dtype=out_dtype)      # ReLU alone
    assert (plain.float() - (A.float() @ B.float().t()).clamp_min(0)).abs().max().item() <= tol


def test_linear_bn_relu_node_matches_torch(cuda):
    """The fused Linear -> BatchNorm(train) -> ReLU autograd node (K6 + K6b) vs torch modules on the
    same bf16-rounded operands: forward, running statistics and all gradients."""
    import hspose_b200.ops as ops
    g = torch.Generator().manual_seed(5)
    M, K, N = 2056, 1296, 512
    x = (torch.randn(M, K, generator=g) * 0.5).cuda().to(torch.bfloat16).requires_grad_()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().requires_grad_()
    b = (torch.randn(N, generator=g) * 0.1).cuda().requires_grad_()
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).cuda().requires_grad_()
    beta = (0.1 * torch.randn(N, generator=g)).cuda().requires_grad_()
    rm, rv = torch.zeros(N, device="cuda"), torch.ones(N, device="cuda")
    z = ops.linear_bn_relu(x, W, b, gamma, beta, rm, rv)
    gz = torch.randn(M, N, generator=g).cuda().to(torch.bfloat16)
    z.backward(gz)
    # reference: fp32 math on the bf16-rounded operands
    xr = x.detach().float().requires_grad_()
    Wr = W.detach().to(torch.bfloat16).float().requires_grad_()
    br, gr, ber = (t.detach().clone().requires_grad_() for t in (b, gamma, beta))
    rm2, rv2 = torch.zeros(N, device="cuda"), torch.ones(N, device="cuda")
    y = xr @ Wr.t() + br
    y = y + (y.to(torch.bfloat16).float() - y).detach()      # the GEMM output is stored in bf16 (straight-through)
    zr = torch.relu(torch.nn.functional.batch_norm(y, rm2, rv2, gr, ber, True, 0.1, 1e-5))
    assert (z.float() - zr).abs().max().item() <= 2 ** -7 * zr.abs().max().item()
    assert torch.allclose(rm, rm2, atol=1e-4) and torch.allclose(rv, rv2, rtol=1e-3, atol=1e-4)
    zr.backward(gz.float())

    def rel(a, r):
        return ((a.float() - r).norm() / r.norm()).item()
    # dY is stored in bf16 after the BatchNorm backward's cancellation (dz - mean - xhat * cov): measured
    # 1.3e-2 relative L2 on dX, 3e-3 on dW (which averages 2056 rows)
    assert rel(x.grad, xr.grad) < 2e-2 and rel(W.grad, Wr.grad) < 1e-2
    assert rel(b.grad, br.grad) < 1e-2 or br.grad.norm().item() < 1e-3      # db ~ 0 through BatchNorm
    assert rel(gamma.grad, gr.grad) < 1e-2 and rel(beta.grad, ber.grad) < 1e-2


@pytest.mark.parametrize("M,N,K,b_mn", [(4112, 128, 128, True), (1028, 256, 256, True), (300, 512, 64, False),
                                        (131584, 128, 128, True)])
def test_gemm_residual_input(cuda, M, N, K, b_mn):
    """hsp_gemm_bf16_acc: out = c_in + A . B^T in the epilogue (the pass-through gradient of a residual connection
    joins the dgrad GEMM), ragged M, every tile width; c_in is not modified."""
    import hspose_b200.ops as ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device=cuda, generator=g).to(torch.bfloat16)
    B = (torch.randn(K, N, device=cuda, generator=g) if b_mn else torch.randn(N, K, device=cuda, generator=g)).to(torch.bfloat16)
    C = torch.randn(M, N, device=cuda, generator=g) * 3
    keep = C.clone()
    ref = C + A.float() @ (B.float() if b_mn else B.float().t())
    out = ops.gemm_bf16(A, B, b_mn=b_mn, out_dtype=torch.float32, c_in=C)
    assert torch.equal(C, keep)
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    plain = ops.gemm_bf16(A, B, b_mn=b_mn, out_dtype=torch.float32)
    assert torch.equal(out, plain + C) or (out - (plain + C)).abs().max().item() <= 1e-6 * ref.abs().max().item()
    assert torch.equal(out, ops.gemm_bf16(A, B, b_mn=b_mn, out_dtype=torch.float32, c_in=C))   # reproducible
