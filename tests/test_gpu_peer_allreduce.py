"""Multi-GPU: the peer-memory gradient exchange of the data-parallel step (needs >= 2 GPUs on the
box; skipped otherwise -- bench.py's multi-rank runs exercise the same path and self test)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_exchange_matches_nccl_route():
    worker = os.path.join(os.path.dirname(__file__), "peer_worker.py")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", worker]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PEER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
