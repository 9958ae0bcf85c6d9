"""Multi-GPU functional parity (needs >= 2 GPUs on the box; skipped otherwise): `sharding.distributed_sliding_window_matching`
(every rank holds both frames and runs a contiguous block of the window list, src/same.py:507-593 per block) gathers exactly the
frame the single-process `sliding_window_matching` returns, and the device-side neighbour exchange of border cells delivers the
same rows over NCCL as over gloo.  The check itself is tools/check_distributed.py, launched under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_sliding_window_matching_equals_single_process(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpus()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_distributed.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "distributed check ok" in r.stdout and "halo exchange ok" in r.stdout and "row-sharded section ok" in r.stdout
