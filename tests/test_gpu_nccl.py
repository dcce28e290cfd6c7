"""-m gpu: the real NCCL transport (one process per GPU).  Needs >= 2 visible GPUs; skipped on a 1-GPU box,
where tests/test_gpu_multirank.py covers the same code above the transport through the loopback group."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_nccl_parity(lib_built, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "nccl_parity_main.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "NCCL_PARITY_RESULT OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
