"""Multi-GPU parity of the `-n` solvers on the bisection + inverse-iteration path (option "select_method" = 2): every
rank bisects all eigenvalues (replicated, bit-identical) and runs inverse iteration for the clusters that touch its
column slab.  Needs >= 2 visible GPUs; skipped otherwise.  Worker: tests/dist_worker.py --cases select."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_select_by_bisection_matches_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29547 + world), os.path.join(ROOT, "tests", "dist_worker.py"), "--cases",
           "select", "--select-method", "2"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout
