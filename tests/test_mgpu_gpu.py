"""Multi-GPU parity inside the GPU suite: torchrun over 2 / 4 / 8 ranks (as many as the box has GPUs) runs
tools/mgpu_check.py -- the sharded step with NCCL halo exchange must equal the single-domain step bit for bit on every
rank's owned cells, including a case where the two boundary families interact across a rank boundary."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_sharded_step_equals_single_domain(ib, nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs, {torch.cuda.device_count()} visible")
    port = 29600 + nranks + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
