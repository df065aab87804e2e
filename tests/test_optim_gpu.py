"""K9 (csrc/optim.cu, hspose_b200/optim.py): clip + Ranger / Adam over the flat buffer.
Ranger: against goldens of the REFERENCE's optimiser (tools/torch_utils/solver/ranger2020.py, 8 steps incl. a
Lookahead step and two clipped steps; tests/golden/make_golden.py::golden_optim).  Adam: against torch.optim.Adam."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPES = [(8, 16), (16,), (4, 8, 1), (3, 24), (5,), (32, 40)]


def _setup(golden, cuda):
    from hspose_b200 import parallel
    g = golden("optim")
    params = [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).to(cuda)) for i in range(len(SHAPES))]
    flat = parallel.FlatGradients(params)
    flat.flatten_params()
    return g, params, flat


def test_ranger_matches_reference_optimiser(cuda, golden):
    from hspose_b200.optim import FlatOptimizer
    g, params, flat = _setup(golden, cuda)
    opt = FlatOptimizer(flat, kind="ranger", lr=1e-2, clip=5.0)
    for step in range(1, 9):
        for i, p in enumerate(params):
            p.grad.copy_(torch.from_numpy(g[f"g{step}_{i}"]).to(cuda))
        norm = opt.step()
        assert abs(norm.item() - float(g[f"norm{step}"])) <= 1e-5 * float(g[f"norm{step}"])
        if step in (1, 5, 6, 8):
            for i, p in enumerate(params):
                ref = torch.from_numpy(g[f"p{step}_{i}"])
                err = (p.detach().cpu() - ref).abs().max().item()
                assert err <= 2e-6, (step, i, err)        # fp32 round-off of a handful of operations
    assert opt.step_count.item() == 8


def test_adam_matches_torch_adam(cuda, golden):
    from hspose_b200.optim import FlatOptimizer
    g, params, flat = _setup(golden, cuda)
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    ref = torch.optim.Adam(ref_params, lr=1e-2)
    opt = FlatOptimizer(flat, kind="adam", lr=1e-2, clip=5.0)
    for step in range(1, 6):
        for i, (p, q) in enumerate(zip(params, ref_params)):
            gr = torch.from_numpy(g[f"g{step}_{i}"]).to(cuda)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
        ref.step()
        opt.step()
    for p, q in zip(params, ref_params):
        assert (p - q).abs().max().item() <= 2e-6
    opt.set_lr(0.0)                                       # the scheduler hook: lr lives on the device
    before = [p.detach().clone() for p in params]
    opt.step()
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, params))
