"""N > 1 on real hardware: the NCCL-reduced flat gradient of the REAL model equals the mean of the
per-shard gradients (per-replica BatchNorm, shared pooling permutation) — SURVEY.md §8(e).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("amp", [0, 1])
def test_nccl_reduced_gradient_equals_mean_of_shard_gradients(cuda, amp):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, HSP_DIST_AMP=str(amp))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_grad_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_GRAD_CHECK" in res.stdout
