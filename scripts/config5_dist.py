"""BASELINE.json config 5 on P B200s (one rank per GPU, torchrun): selecting solve (general_scalapack_select-style, -n;
reference src/solver_scalapack_select.f90:14-69) on the synthetic STANDARD problem n = 65536 (seed 20240603), lowest
6554 eigenpairs, eigenvector columns sharded over the ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
      scripts/config5_dist.py [--n 65536] [--k 6554] [--solo]

Prints one JSON line (rank 0): seconds (device, max over ranks), stage times, which tridiagonal eigensolver the auto-switch
picked, the reference's residual / orthogonality metrics computed on the device over all k pairs, bit-identity of the
eigenvalues across ranks and (with --solo) the difference to a single-GPU solve of the same problem on rank 0."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("EKB200_TEST_DEADLINE", "600")), exit=True)
    import torch
    import torch.distributed as dist

    from eigenkernel_b200 import dist as ekdist
    from eigenkernel_b200.device import Context

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--k", type=int, default=6554)
    ap.add_argument("--seed", type=int, default=20240603)
    ap.add_argument("--solo", action="store_true")
    ap.add_argument("--select-method", type=int, default=0)
    args = ap.parse_args()
    n, k = args.n, args.k
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    if world > 1:
        ekdist.attach(ctx)
    ctx.set_option("select_method", args.select_method)
    ld = (n + 7) // 8 * 8
    dA, dZ, dw = ctx.alloc(ld * n * 8), ctx.alloc(ld * k * 8), ctx.alloc((n + 8) * 8)

    def fill():
        ctx.call("ekb200_fill_synthetic", n, args.seed, 1.0, 0, 0.0, dA, ld)

    def barrier():
        ctx.sync()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    fill()
    barrier()
    ctx.clear_events()
    sec = ctypes.c_double()
    ctx.call("ekb200_timer_start")
    info = ctx.call("ekb200_syevd_dev", n, k, dA, ld, dw, dZ, ld)
    ctx.call("ekb200_timer_stop", ctypes.byref(sec))
    barrier()
    seconds = sec.value
    if world > 1:
        t = torch.tensor([seconds], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    events = {name: s for name, s, _ in ctx.events()}
    w = np.zeros(n)
    ctx.call("ekb200_d2h", w.ctypes.data, dw, n * 8)
    same_bits = True
    if world > 1:
        wt = torch.from_numpy(w.copy()).cuda()
        parts = [torch.empty_like(wt) for _ in range(world)]
        dist.all_gather(parts, wt)
        same_bits = all(torch.equal(p.view(torch.int64), parts[0].view(torch.int64)) for p in parts)
    # acceptance on the device: regenerate A, make all k columns available on every rank, the reference's two metrics
    fill()
    if world > 1:
        ctx.call("ekb200_comm_allgather_slabs", n, k, dZ, ld)
    an, ave, mx, o, g = (ctypes.c_double() for _ in range(5))
    ctx.call("ekb200_eval_residual_norm_dev", n, k, dA, ld, None, 0, dw, dZ, ld, ctypes.byref(an), ctypes.byref(ave),
             ctypes.byref(mx))
    ctx.call("ekb200_eval_b_orthonormality_dev", n, 1, k, dZ, ld, None, 0, ctypes.byref(o), ctypes.byref(g))
    out = {"what": "config 5", "n": n, "k": k, "n_gpus": world, "info": info, "seconds": seconds, "events": events,
           "tridiagonal_eigensolver": "bisection + inverse iteration" if "eigen_solver_b200:stebz_stein" in events
           else "divide and conquer", "eigenvalues_bit_identical_across_ranks": bool(same_bits),
           "ascending": bool(np.all(np.diff(w[:k]) >= 0)), "A_norm_fro": an.value, "residual_max_over_A": mx.value,
           "residual_ave_over_A": ave.value, "orthogonality_verifier": o.value, "xtx_minus_identity_fro": g.value,
           "tolerance": 1e-12 * n, "collectives": int(ctx.lib.ekb200_num_collectives(ctx.h))}
    for p in (dA, dZ, dw):
        ctx.free(p)
    ctx.set_option("cache_device_memory", 0)
    barrier()
    if args.solo and rank == 0 and world > 1:
        solo = Context(local)
        solo.set_option("select_method", args.select_method)
        a, z, wd = solo.alloc(ld * n * 8), solo.alloc(ld * k * 8), solo.alloc((n + 8) * 8)
        solo.call("ekb200_fill_synthetic", n, args.seed, 1.0, 0, 0.0, a, ld)
        solo.call("ekb200_timer_start")
        i1 = solo.call("ekb200_syevd_dev", n, k, a, ld, wd, z, ld)
        s1 = ctypes.c_double()
        solo.call("ekb200_timer_stop", ctypes.byref(s1))
        w1 = np.zeros(n)
        solo.call("ekb200_d2h", w1.ctypes.data, wd, n * 8)
        solo.close()
        out["solo"] = {"info": i1, "seconds": s1.value,
                       "max_dlambda_over_max_lambda": float(np.abs(w[:k] - w1[:k]).max() / np.abs(w1).max())}
    if rank == 0:
        out["pass"] = bool(info == 0 and mx.value <= 1e-12 * n and g.value <= 1e-12 * n and same_bits)
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
