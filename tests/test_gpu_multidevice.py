"""Two (or more) real devices: the peer exchange across GPUs against the same compartments exchanged inside one GPU.
Skipped on a box with one GPU (the in-process tests of tests/test_gpu_migrants.py cover the kernels there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_peer_exchange_across_devices_equals_in_process_exchange(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29640 + world), os.path.join(here, "mp_exchange_worker.py"), "200000", "5"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTIDEVICE OK" in p.stdout, p.stdout[-3000:]
