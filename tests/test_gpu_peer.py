"""GPU test (needs two GPUs on the box; skipped otherwise): the GOP-boundary hand-off of config 4 by peer-memory copies -- rank 0's receive buffers
mapped into the producers (svtb200_peer_* of the C ABI), pictures moved device to device over NVLink, torch.distributed carrying only the byte counts."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_peer_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(HERE, "peer_exchange_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert "PEER_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
