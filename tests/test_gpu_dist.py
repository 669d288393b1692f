"""Multi-GPU tests (skipped on a one-GPU box): the peer-memory gradient exchange of csrc/comm.cu (mvae_comm_*) against NCCL
and a float64 sum, and the data-parallel step against single-GPU steps on the concatenated shards.  Both are the torchrun
scripts under scripts/, run here on two GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, env=None, args=()):
    e = dict(os.environ, **(env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_two_gpus():
    out = _torchrun("comm_check.py", 29561, env={"STEP": "0"})
    assert "comm_check ok" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_gradients_two_gpus():
    """2 ranks x B against 1 rank over both shards: the exchanged gradient is the mean of the per-replica gradients."""
    out = _torchrun("dp_check.py", 29562, env={"B": "32"}, args=("cfg1",))
    assert "max err" in out
