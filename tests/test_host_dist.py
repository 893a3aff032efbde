"""Host-side logic of the multi-GPU path on CPU: the column-slab layout (Python mirror vs the C++ of
csrc/layout.h through the host-check library) and the torch.distributed plumbing with world_size 2 on gloo
(unique-id hand-over, slab gather, max-over-ranks)."""
import ctypes
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from eigenkernel_b200 import dist as ekdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "eigenkernel_b200", "libekb200_hostcheck.so")


def test_slab_bounds_mirror_matches_c():
    hc = ctypes.CDLL(HC)
    for ncols in (0, 1, 30, 127, 128, 129, 400, 6554, 8192, 32768, 65536):
        for P in (1, 2, 3, 4, 8):
            for gran in (64, 128):
                out = (ctypes.c_longlong * (P + 1))()
                hc.ekb200_host_slab_bounds(ctypes.c_longlong(ncols), P, gran, out)
                assert list(out) == ekdist.slab_bounds(ncols, P, gran)


def test_slab_bounds_cover_and_align():
    for ncols in (1, 30, 400, 6554, 32768):
        for P in (1, 2, 4, 8):
            b = ekdist.slab_bounds(ncols, P)
            assert b[0] == 0 and b[-1] == ncols and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 128 == 0 for x in b[:-1] if x < ncols)
            assert sum(ekdist.local_slab(ncols, P, r)[1] for r in range(P)) == ncols


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from eigenkernel_b200 import dist as ekdist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = ekdist.exchange_unique_id(lambda: bytes(range(128)))
    assert uid == bytes(range(128)), uid
    n, ncols = 37, 300
    full = np.arange(n * ncols, dtype=np.float64).reshape(n, ncols)
    c0, kc = ekdist.local_slab(ncols, world, rank)
    got = ekdist.gather_columns(np.asfortranarray(full[:, c0:c0 + kc]), ncols)
    if rank == 0:
        assert np.array_equal(got, full)
    else:
        assert got is None
    t = ekdist.max_over_ranks(1.0 + rank)
    assert t == float(world)
    dist.barrier()
    if rank == 0:
        print("HOST_DIST_OK")
    dist.destroy_process_group()
""")


def test_gloo_world2_plumbing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_DIST_OK" in r.stdout, r.stdout[-3000:]
