"""The multi-PROCESS machinery of the distributed transforms on a box with ONE GPU: two ranks share device 0
(tests/same_gpu_worker.py).  tests/test_gpu_dist.py is the one-rank-per-GPU version and needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.no_emu]
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_processes_one_gpu_p2p_transport():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tests", "same_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if "exclusive" in (r.stdout + r.stderr).lower() and "SAMEGPU" not in r.stdout:
        pytest.skip("the device is in exclusive-process mode: two contexts cannot share it")
    assert r.returncode == 0 and "SAMEGPU PASSED" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]
