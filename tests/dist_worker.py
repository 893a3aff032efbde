"""torchrun worker of the multi-GPU parity tests (one rank per B200):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tests/dist_worker.py [--cases small|full|select] [--select-method 0|1|2]

Every rank solves the same synthetic problem through the SAME C-ABI entry points as the single-GPU path, with
the context attached to the NCCL ranks; rank 0 then checks (a) BASELINE.json's acceptance metrics against the
oracle's host matrices, (b) the eigenvalues against a single-GPU solve of the same context-free library,
(c) that every rank returned bit-identical eigenvalues (replicated stages are deterministic).
Prints one line `DIST_CHECK_OK cases=<k>` on success; any failure raises."""
from __future__ import annotations

import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    # a hung collective must not burn GPU time: dump every thread's stack and exit after the deadline
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("EKB200_TEST_DEADLINE", "240")), exit=True)
    import torch
    import torch.distributed as dist

    from eigenkernel_b200 import dist as ekdist
    from eigenkernel_b200.device import Context
    from oracle import lapack_twin as lt

    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="small")
    ap.add_argument("--select-method", type=int, default=0, help="0 auto | 1 divide and conquer | 2 bisection + inverse iteration")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = Context(local)
    r2, w2 = ekdist.attach(ctx)
    assert (r2, w2) == (rank, world)
    nr, rk = ctypes.c_int(), ctypes.c_int()
    ctx.lib.ekb200_comm_info(ctx.h, ctypes.byref(nr), ctypes.byref(rk))
    assert (nr.value, rk.value) == (world, rank)
    solo = Context(local)  # same GPU, no communicator: the single-GPU answer
    ctx.set_option("select_method", args.select_method)
    solo.set_option("select_method", args.select_method)

    # (n, nev, generalized, seed)
    cases = [(700, 700, True, 11), (1000, 1000, False, 12), (2048, 2048, True, 13), (1536, 200, True, 14),
             (1536, 300, False, 15)]
    if args.cases == "full":
        cases += [(4096, 4096, True, 16), (8192, 8192, True, 20240601)]
    if args.cases == "select":  # the -n solvers only: slab borders inside the requested range, odd widths
        cases = [(1536, 200, True, 14), (1536, 300, False, 15), (3000, 700, False, 17), (2000, 1999, True, 18),
                 (900, 129, False, 19)]
    done = 0
    for n, nev, gen, seed in cases:
        if rank == 0:
            print(f"[dist_check] case n={n} nev={nev} gen={gen}", flush=True)
        ld = (n + 7) // 8 * 8
        c0, kc = ekdist.local_slab(nev, world, rank)
        cc0, ckc = ctypes.c_int64(), ctypes.c_int64()
        ctx.lib.ekb200_comm_slab(ctx.h, nev, ctypes.byref(cc0), ctypes.byref(ckc))
        assert (cc0.value, ckc.value) == (c0, kc)

        def solve(c):
            dA, dB, dZ, dw = c.alloc(ld * n * 8), c.alloc(ld * n * 8), c.alloc(ld * n * 8), c.alloc((n + 8) * 8)
            c.call("ekb200_fill_synthetic", n, seed, 1.0, 0, 0.0, dA, ld)
            c.call("ekb200_fill_synthetic", n, seed + 1, float(n), 1, 2.0, dB, ld)
            if gen:
                info = c.call("ekb200_sygvd_dev", n, nev, dA, ld, dB, ld, dw, dZ, ld)
            else:
                info = c.call("ekb200_syevd_dev", n, nev, dA, ld, dw, dZ, ld)
            assert info == 0, info
            return dA, dB, dZ, dw

        dA, dB, dZ, dw = solve(ctx)
        ctx.call("ekb200_comm_allgather_slabs", n, nev, dZ, ld)
        w = np.zeros(n)
        ctx.call("ekb200_d2h", w.ctypes.data, dw, n * 8)
        X = np.zeros((n, nev), order="F")
        ctx.call("ekb200_d2h_matrix", X.ctypes.data, n, dZ, ld, n, nev)
        for p in (dA, dB, dZ, dw):
            ctx.free(p)
        # (c) replicated eigenvalues are bit-identical on every rank
        wt = torch.from_numpy(w.copy()).cuda()
        parts = [torch.empty_like(wt) for _ in range(world)]
        dist.all_gather(parts, wt)
        for p in parts:
            assert torch.equal(p.view(torch.int64), parts[0].view(torch.int64)), "eigenvalues differ between ranks"
        # host-pointer entry point: replicated host matrices in, LOCAL piece out
        A, B = lt.synthetic_pair(n, seed)
        wl, Xl = np.zeros(n), np.zeros((n, max(kc, 1)), order="F")
        if gen:
            info = ctx.call("ekb200_sygvd", n, nev, A.ctypes.data, n, B.ctypes.data, n, wl.ctypes.data,
                            Xl.ctypes.data, n)
        else:
            info = ctx.call("ekb200_syevd", n, nev, A.ctypes.data, n, wl.ctypes.data, Xl.ctypes.data, n)
        assert info == 0
        assert np.array_equal(wl, w), "host and device entry points disagree on eigenvalues"
        if kc > 0:
            assert np.array_equal(Xl[:, :kc], X[:, c0:c0 + kc]), "local piece != slab of the device result"
        # the caller's ScaLAPACK view (rank-per-GPU Fortran mode): with option "out_block" = the NB that the reference's
        # setup_distributed_matrix ends up with (clamped to floor(nev / P), distribute_matrix.f90:114-120) the host entry
        # point delivers the numroc columns of that 1 x P block-cyclic descriptor: blocks rank, rank + P, ...
        nb = max(min(n, nev // world), 1)
        nb = min(nb, max(ekdist.slab_bounds(nev, world)[1], 1))
        ctx.set_option("out_block", nb)
        nl = ctypes.c_int64()
        ctx.lib.ekb200_comm_local_cols(ctx.h, nev, ctypes.byref(nl))
        nblk = nev // nb
        want = (nblk // world) * nb + (nb if rank < nblk % world else (nev % nb if rank == nblk % world else 0))
        assert nl.value == want, (nl.value, want)
        wc, Xc = np.zeros(n), np.zeros((n, max(nl.value, 1)), order="F")
        if gen:
            info = ctx.call("ekb200_sygvd", n, nev, A.ctypes.data, n, B.ctypes.data, n, wc.ctypes.data, Xc.ctypes.data, n)
        else:
            info = ctx.call("ekb200_syevd", n, nev, A.ctypes.data, n, wc.ctypes.data, Xc.ctypes.data, n)
        assert info == 0 and np.array_equal(wc, w)
        gcols = [((lc // nb) * world + rank) * nb + lc % nb for lc in range(nl.value)]
        assert np.array_equal(Xc[:, :nl.value], X[:, gcols]), "block-cyclic local piece != columns of the device result"
        ctx.set_option("out_block", 0)
        # device-side checks through the host entry points: replicated COO in, local piece of X in (perturbed, so
        # that the metrics are well above rounding noise and can be compared tightly)
        Xl = np.asfortranarray(Xl + 1e-9 * np.random.default_rng(1000 + rank).standard_normal(Xl.shape))
        i, j = np.tril_indices(n)
        ij = np.ascontiguousarray(np.stack([i + 1, j + 1], axis=1).astype(np.int32))
        vA, vB = np.ascontiguousarray(A[i, j]), np.ascontiguousarray(B[i, j])
        nnz = len(i)
        an, ave, mx, orth = (ctypes.c_double() for _ in range(4))
        ipr = np.zeros(nev)
        nnzB = nnz if gen else 0
        ctx.call("ekb200_eval_residual_norm", n, nev, nev, nnz, ij.ctypes.data, vA.ctypes.data, nnzB,
                 ij.ctypes.data if gen else None, vB.ctypes.data if gen else None, wl.ctypes.data, Xl.ctypes.data, n,
                 ctypes.byref(an), ctypes.byref(ave), ctypes.byref(mx))
        ctx.call("ekb200_eval_orthogonality", n, nev, 1, nev, nnzB, ij.ctypes.data if gen else None,
                 vB.ctypes.data if gen else None, Xl.ctypes.data, n, ctypes.byref(orth))
        ctx.call("ekb200_get_ipratios", n, nev, nnzB, ij.ctypes.data if gen else None, vB.ctypes.data if gen else None,
                 Xl.ctypes.data, n, ipr.ctypes.data)
        full = ekdist.gather_columns(Xl[:, :kc], nev)
        if rank == 0:
            assert np.max(np.abs(full - X)) <= 1e-8
            dA, dB, dZ1, dw1 = solve(solo)
            w1 = np.zeros(n)
            solo.call("ekb200_d2h", w1.ctypes.data, dw1, n * 8)
            for p in (dA, dB, dZ1, dw1):
                solo.free(p)
            Bm = B if gen else None
            r = lt.residual_metrics(A, w[:nev], X, Bm)
            o = lt.orthogonality_metrics(X, Bm)
            scale = np.abs(w1).max()
            dw_rel = float(np.max(np.abs(w - w1)) / scale)
            print(f"[dist_check] P={world} n={n} nev={nev} gen={gen}: res={r['res_max_over_A']:.2e} "
                  f"orth={o['orth_fro']:.2e} dlambda_vs_1gpu={dw_rel:.2e} collectives="
                  f"{ctx.lib.ekb200_num_collectives(ctx.h)}", flush=True)
            rp = lt.residual_metrics(A, w[:nev], full, Bm)
            vo = lt.orthogonality_metrics(full, Bm)["verifier_orthogonality"]
            assert abs(an.value - rp["A_norm"]) <= 1e-13 * rp["A_norm"]
            assert abs(mx.value - rp["res_max_over_A"]) <= 1e-5 * rp["res_max_over_A"], (mx.value, rp)
            assert abs(ave.value - rp["res_avg_over_A"]) <= 1e-5 * rp["res_avg_over_A"]
            assert abs(orth.value - vo) <= 1e-5 * vo, (orth.value, vo)
            assert np.max(np.abs(ipr - lt.ipratios(full, Bm)) / lt.ipratios(full, Bm)) <= 1e-10
            assert r["res_max_over_A"] <= 1e-12 * n, r
            assert o["orth_fro"] <= 1e-12 * n, o
            assert dw_rel <= 1e-12
            assert np.all(np.diff(w) >= 0)
        dist.barrier()
        done += 1
    if rank == 0:
        print(f"DIST_CHECK_OK cases={done} world={world}", flush=True)
    solo.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
