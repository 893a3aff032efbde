"""Multi-GPU parity: the sharded solve on 2 B200s must meet the same acceptance bars as the
single-GPU one and reproduce its eigenvalues.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise.
The worker is tests/dist_worker.py, launched the way the driver launches bench.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_solve_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29517 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout
