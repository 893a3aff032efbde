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


# ---- the caller's block-cyclic view of the eigenvectors (option "out_block", rank-per-GPU Fortran mode)
def _numroc_ref(n, nb, iproc, isrc, nprocs):
    """ScaLAPACK TOOLS/numroc.f, restated line by line."""
    mydist = (nprocs + iproc - isrc) % nprocs
    nblocks = n // nb
    out = (nblocks // nprocs) * nb
    extra = nblocks % nprocs
    if mydist < extra:
        out += nb
    elif mydist == extra:
        out += n % nb
    return out


@pytest.mark.parametrize("n,n_vec,P", [(30, 30, 2), (400, 400, 4), (65536, 6554, 8), (32768, 32768, 8), (1000, 7, 8),
                                       (3000, 700, 4), (2000, 1999, 2)])
def test_out_block_layout_matches_setup_distributed_matrix(n, n_vec, P):
    """fortran/solver_b200.f90 allocates blacs%Vectors with the reference's setup_distributed_matrix, which clamps the
    block size to max(min(rows/nprow, cols/npcol), 1) (distribute_matrix.f90:114-120); the library must deliver exactly
    the numroc columns of THAT descriptor (blocks r, r + P, ...), not its own 128-granular slabs: (30, 2) gives NB 15
    (slab 128), (6554, 8) gives 819 (slab 896)."""
    import ctypes

    lib = ctypes.CDLL(HC)
    for f in (lib.ekb200_host_block_clamp, lib.ekb200_host_numroc, lib.ekb200_host_cyclic_global_col):
        f.restype = ctypes.c_longlong
    LL = ctypes.c_longlong
    slab = ekdist.slab_bounds(n_vec, P)[1]                          # what the glue asks setup_distributed_matrix for
    nb = lib.ekb200_host_block_clamp(LL(n), LL(n_vec), LL(max(slab, 1)), 1, P)
    assert nb == min(max(slab, 1), max(min(n, n_vec // P), 1))
    assert {(30, 2): 15, (400, 4): 100, (6554, 8): 819, (32768, 8): 4096}.get((n_vec, P), nb) == nb
    owned = []
    for r in range(P):
        nloc = lib.ekb200_host_numroc(LL(n_vec), LL(nb), r, P)
        assert nloc == _numroc_ref(n_vec, nb, r, 0, P)
        cols = [lib.ekb200_host_cyclic_global_col(LL(lc), LL(nb), P, r) for lc in range(nloc)]
        # ScaLAPACK's INDXL2G for every local column
        assert cols == [((lc // nb) * P + r) * nb + lc % nb for lc in range(nloc)]
        assert all(0 <= c < n_vec for c in cols)
        owned += cols
    assert sorted(owned) == list(range(n_vec))               # every column delivered exactly once
