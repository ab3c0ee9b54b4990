"""Data-parallel parity: tools/dp_parity.py under torchrun — the gradients a rank holds after the bucketed,
overlapped all-reduce must equal the average of the per-shard gradients computed without any collective.
With >= 2 GPUs: NCCL, one rank per GPU.  On a single-GPU box (the driver's round-end run): the SAME staged / bucketed
path with two ranks sharing GPU 0 over gloo (NCCL refuses two ranks on one device), so the path is never skipped.
The host-side bucketing logic alone is also covered on the CPU by tests/test_host_cpu.py (two gloo ranks)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n, backend, port):
    env = dict(os.environ, DP_PARITY_BACKEND=backend)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    print(line)
    assert out["ok"] and out["world"] == n, out
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_gradients_equal_the_shard_average():
    _run(min(torch.cuda.device_count(), 4), "nccl", 29541)


def test_dp_bucketed_path_two_ranks_on_one_gpu():
    out = _run(2, "gloo", 29543)
    assert out["grad_rel_l2_whole"] < 1e-4
