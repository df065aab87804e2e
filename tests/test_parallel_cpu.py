"""CPU, gloo, world_size 2: the data-parallel host logic (hs-pose_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from hspose_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    flat = parallel.FlatGradients(net.parameters())
    g = torch.Generator().manual_seed(7)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 2, generator=g), "tag": "t"}
    mine = parallel.shard_batch(batch, rank, world)
    assert mine["x"].shape[0] == 4 and mine["tag"] == "t"
    flat.zero()
    torch.nn.functional.mse_loss(net(mine["x"]), mine["y"]).backward()
    flat.all_reduce_mean()
    # single-process gradient of the whole batch (equal shard sizes -> mean of shard grads)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    ref.load_state_dict(net.state_dict())
    torch.nn.functional.mse_loss(ref(batch["x"]), batch["y"]).backward()
    # every parameter starts on a 16-byte boundary of the flat buffer; the padding stays zero
    ok = all(o % 4 == 0 for o in flat.offsets) and flat.flat.numel() % 4 == 0
    ok = ok and all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(net.parameters(), ref.parameters()))
    ok = ok and abs(float(flat.flat.sum()) - float(sum(q.grad.sum() for q in ref.parameters()))) < 1e-5
    # views stay attached after zero()
    flat.zero()
    ok = ok and all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())
    ok = ok and float(sum(p.grad.abs().sum() for p in net.parameters())) == 0.0
    # same permutation on every rank after seed_all
    parallel.seed_all(123)
    perm = torch.randperm(1028)[:257]
    gathered = [torch.empty_like(perm) for _ in range(world)]
    dist.all_gather(gathered, perm)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_equals_big_batch():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_clip_matches_torch():
    import sys
    from hspose_b200 import parallel
    net = torch.nn.Linear(10, 10)
    flat = parallel.FlatGradients(net.parameters())
    net(torch.ones(3, 10)).sum().backward()
    ref = [p.grad.clone() for p in net.parameters()]
    n1 = flat.clip_(0.5)
    n2 = torch.nn.utils.clip_grad_norm_([torch.nn.Parameter(torch.zeros_like(r)) for r in ref], 0.5)
    tot = torch.sqrt(sum((r ** 2).sum() for r in ref))
    assert torch.allclose(n1, tot)
    for p, r in zip(net.parameters(), ref):
        assert torch.allclose(p.grad, r * min(1.0, 0.5 / (tot.item() + 1e-6)), atol=1e-6)


def test_flat_parameter_adam_equals_per_tensor_adam():
    """FlatGradients.flatten_params: parameters become views of one flat buffer (values, names and
    state_dict unchanged) and Adam over that single flat parameter takes exactly the per-tensor step."""
    import copy
    from hspose_b200 import parallel
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    net_b = copy.deepcopy(net_a)
    sd0 = {k: v.clone() for k, v in net_b.state_dict().items()}
    flat = parallel.FlatGradients(net_b.parameters())
    fp = flat.flatten_params()
    for k, v in net_b.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    opt_a = torch.optim.Adam(net_a.parameters(), lr=1e-2)
    opt_b = torch.optim.Adam([fp], lr=1e-2)
    x = torch.randn(16, 7)
    for _ in range(3):
        opt_a.zero_grad()
        net_a(x).square().mean().backward()
        opt_a.step()
        flat.zero()
        net_b(x).square().mean().backward()
        opt_b.step()
    for (n, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        assert torch.allclose(pa, pb, atol=1e-7), n


def test_flat_buffers_alignment_and_bf16_parameter_shadow():
    """FlatGradients: every parameter starts on a 32-byte boundary of the flat fp32 buffers (= 16 bytes in the bf16
    shadow: TMA descriptors of the tensor-core GEMMs need that), flatten_params keeps values and identities, and
    the shadow refreshed by ONE cast serves any view of a parameter bit-for-bit like a per-tensor cast would."""
    import torch
    from hspose_b200 import ops, parallel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv1d(5, 7, 1), torch.nn.BatchNorm1d(7), torch.nn.Linear(3, 11))
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    flat = parallel.FlatGradients(net.parameters())
    assert all(off % 8 == 0 for off in flat.offsets)
    flat.flatten_params()
    for n, p in net.named_parameters():
        assert torch.equal(p, before[n]) and p.data_ptr() % 32 == flat.flat_param.data_ptr() % 32
    shadow = flat.refresh_shadow()
    assert shadow.dtype == torch.bfloat16 and shadow.numel() == flat.flat_param.numel()
    w = net[0].weight.squeeze(-1)[:, 1:4]                 # a strided view of a parameter
    with ops.weight_shadow(flat.flat_param, shadow):
        served = ops._as_gemm_operand(net[2].weight)      # contiguous (11, 3) -> padded copy (pitch not 16 B)
        assert torch.equal(served.float(), net[2].weight.detach().to(torch.bfloat16).float())
        off = (w.data_ptr() - flat.flat_param.data_ptr()) // 4
        view = shadow.as_strided(w.shape, w.stride(), off)
        assert torch.equal(view.float(), w.detach().to(torch.bfloat16).float())
    assert ops._shadow is None                            # scoped: nothing leaks out of the step
    with torch.no_grad():
        net[2].weight.add_(1.0)
    assert not torch.equal(shadow.as_strided(net[2].weight.shape, net[2].weight.stride(),
                                             (net[2].weight.data_ptr() - flat.flat_param.data_ptr()) // 4).float(),
                           net[2].weight.detach().to(torch.bfloat16).float())     # stale until refreshed ...
    flat.refresh_shadow()
    assert torch.equal(shadow.as_strided(net[2].weight.shape, net[2].weight.stride(),
                                         (net[2].weight.data_ptr() - flat.flat_param.data_ptr()) // 4).float(),
                       net[2].weight.detach().to(torch.bfloat16).float())         # ... which the engine does every step
