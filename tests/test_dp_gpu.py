"""Data-parallel correctness on real GPUs over NCCL (needs >= 2 GPUs on the box; the 1-GPU tier skips it -- bench.py
runs the same dp.replica_check inside every N > 1 run and reports it as ``dp_check`` in its JSON line)."""
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
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_replicas_stay_identical_and_staged_exchange_is_the_mean():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "dp_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
