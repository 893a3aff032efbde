#!/usr/bin/env python
"""bench.py -- generalized dense FP64 eigensolve A x = lambda B x (all eigenpairs), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 32768]

A "step" is one full solve of the synthetic generalized problem of SURVEY.md 8(d) (counter-hash A, B =
2 I + U/n, seed 20240602): device fill of A and B (the setup_matrices stage) + ekb200_sygvd_dev.  `value` is
canonical TFLOP/s (7 n^3 FLOPs per solve / device seconds, inputs generated in HBM); `e2e` is the same metric
through the reference-facing host-pointer entry point ekb200_sygvd with pinned HOST buffers (H2D of A and B,
D2H of eigenvalues and eigenvectors inside the timed region).

Multi-GPU: one rank per B200 (torchrun); ONE problem is solved by all ranks together ("strong" scaling): the
reduction to standard form, the dense-to-band trailing updates, the top D&C merge, both back-transformations and
the final triangular solve are sharded (NCCL all-gathers / panel exchanges inside libekb200.so), Cholesky, bulge
chasing and the lower D&C levels are replicated -- see DESIGN.md 6.  `value` = canonical FLOPs of the one solve /
max-over-ranks device seconds.

--impl reference times the reference's CPU path: the reference itself (Fortran + MPI + ScaLAPACK) cannot be
built in this image, so the arm runs the oracle port (serial-LAPACK twins of the reference's call sequence,
OpenBLAS with all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generalized_eigensolve_fp64_tflops"
UNIT = "TFLOP/s"
SEED = 20240602


_REAL_STDOUT = None


def quiet_stdout() -> None:
    """Rank 0's stdout must carry exactly ONE JSON line: route everything libraries print to fd 1 (NCCL's version
    banner, for one) to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def canonical_flops(n: int) -> float:
    return 7.0 * float(n) ** 3  # n^3/3 + n^3 + 4n^3/3 + 4n^3/3 + 2n^3 + n^3


def workload_name(n: int) -> str:
    return f"synthetic generalized A x = lambda B x, n={n}, FP64, all eigenpairs + eigenvectors (seed {SEED})"


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_solve_once(n: int):
    from oracle import lapack_twin as lt

    A, B = lt.synthetic_pair(n, SEED)
    tm: dict = {}
    t0 = time.perf_counter()
    lt.general_scalapack_twin(A, B, tm)
    return time.perf_counter() - t0, tm


def cpu_threads() -> int:
    from oracle import lapack_twin as lt

    want = min(os.cpu_count() or 1, 64)
    lt.set_num_threads(want)
    return lt.get_num_threads()


def cpu_rate_table(sizes, cores: int) -> dict:
    """The oracle port timed once at each n of `sizes` (rates at >= 3 sizes so the trend towards the
    metric's n = 32768 is visible) + a cubic extrapolation of the seconds at n = 32768 from the largest sample,
    LABELLED as extrapolated (the rate still rises slowly with n, so this is an upper bound on the CPU seconds)."""
    rows = []
    for m in sizes:
        dt, tm = cpu_solve_once(m)
        rows.append({"n": m, "seconds": dt, "tflops": canonical_flops(m) / dt / 1e12, "stage_seconds": tm})
    big = rows[-1]
    return {"cores": cores, "sizes": rows,
            "extrapolated_seconds_n32768": big["seconds"] * (32768.0 / big["n"]) ** 3,
            "extrapolation": f"EXTRAPOLATED, not measured: seconds at n={big['n']} x (32768/{big['n']})^3, i.e. the "
                             f"rate measured at n={big['n']} held constant"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    cores = cpu_threads()
    for _ in range(args.warmup):
        cpu_solve_once(min(n, 1024))  # warm the BLAS threads; the sample itself is seconds long
    t = []
    stages = {}
    for _ in range(args.steps):
        dt, tm = cpu_solve_once(n)
        t.append(dt)
        stages = tm
    sec = sum(t) / len(t)
    val = canonical_flops(n) / sec / 1e12
    table = None
    if not args.no_cpu_table:  # once, outside the K timed steps
        table = cpu_rate_table([x for x in args.cpu_sizes if x != n], cores)
        table["sizes"].append({"n": n, "seconds": sec, "tflops": val, "stage_seconds": stages})
        table["sizes"].sort(key=lambda r: r["n"])
        big = table["sizes"][-1]
        table["extrapolated_seconds_n32768"] = big["seconds"] * (32768.0 / big["n"]) ** 3
    sample = (f"oracle port (serial-LAPACK twins dpotrf/dsygst/dsytrd/dstedc/dormtr/dtrtrs, OpenBLAS, {cores} threads) "
              f"on n={n} of the same generator; canonical 7n^3 FLOPs; the Fortran/MPI/ScaLAPACK reference cannot be "
              f"built in this image")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n), "sample_n": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stage_seconds": stages, "rate_table": table},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="ekb200_clocks_", suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            w = [x.strip() for x in line.split(",")]
            if len(w) < 7:
                continue
            try:
                sm.append(float(w[0]))
                mx.append(float(w[1]))
            except ValueError:
                continue
            for nm, flag in zip(names, w[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


# ------------------------------------------------------------------------------------------ acceptance
def acceptance(ctx, n, ld, dA, dB, dZ, dw, fill, world, rank, local, dist, w_last) -> dict:
    """BASELINE.json's acceptance numbers for the LAST timed solve, computed on the device by the library's twins of
    the reference's own checks (verifier.f90:140-199 residual, :279-325 orthogonality; called as main.f90:149-179
    calls them): max_j ||A x_j - lambda_j B x_j||_2 / ||A||_F, the scaled-Gram metric, and || X^T B X - I ||_F.
    With P > 1 ranks the eigenvalues are also compared with a single-GPU solve of the same problem (a second,
    communicator-less context on rank 0)."""
    import numpy as np
    from ctypes import byref, c_double

    fill()  # the solve destroyed A and left L in B: regenerate both from the counter hash
    if world > 1:
        ctx.call("ekb200_comm_allgather_slabs", n, n, dZ, ld)  # the Gram matrix needs every column on every rank
    an, ave, mx, o, g = c_double(), c_double(), c_double(), c_double(), c_double()
    ctx.call("ekb200_eval_residual_norm_dev", n, n, dA, ld, dB, ld, dw, dZ, ld, byref(an), byref(ave), byref(mx))
    ctx.call("ekb200_eval_b_orthonormality_dev", n, 1, n, dZ, ld, dB, ld, byref(o), byref(g))
    tol = 1e-12 * n
    out = {"n": n, "checked_vectors": n, "A_norm_fro": an.value,
           "residual_max_over_A": mx.value, "residual_ave_over_A": ave.value,
           "orthogonality_verifier": o.value, "xtbx_minus_identity_fro": g.value, "tolerance": tol,
           "definitions": "residual = max_j ||A x_j - lambda_j B x_j||_2 / ||A||_F (verifier.f90:179-199, no division "
                          "by ||x_j||; ||x_j||_2 ~ 0.7 here since B ~ 2 I); orthogonality_verifier = Frobenius norm of "
                          "the unit-diagonal-scaled Gram matrix X^T B X with its diagonal zeroed (verifier.f90:310-325)"}
    dl = None
    if world > 1:
        ctx.set_option("cache_device_memory", 0)  # hand the arena's cached blocks back before the solo solve
        ctx.set_option("cache_device_memory", 1)
        if rank == 0:
            from eigenkernel_b200.device import Context

            solo = Context(local)
            a, b, z = solo.alloc(ld * n * 8), solo.alloc(ld * n * 8), solo.alloc(ld * n * 8)
            w1d = solo.alloc((n + 8) * 8)
            solo.call("ekb200_fill_synthetic", n, SEED, 1.0, 0, 0.0, a, ld)
            solo.call("ekb200_fill_synthetic", n, SEED + 1, float(n), 1, 2.0, b, ld)
            info = solo.call("ekb200_sygvd_dev", n, n, a, ld, b, ld, w1d, z, ld)
            assert info == 0, info
            w1 = np.zeros(n)
            solo.call("ekb200_d2h", w1.ctypes.data, w1d, n * 8)
            solo.close()
            scale = float(np.abs(w1).max())
            dl = {"max_abs_over_max_lambda": float(np.abs(w_last - w1).max() / scale),
                  "max_relative": float((np.abs(w_last - w1) / np.maximum(np.abs(w1), 1e-300)).max()),
                  "smallest_abs_lambda": float(np.abs(w1).min()), "tolerance": 1e-10,
                  "note": "max_relative is ill-conditioned for |lambda| << ||A|| (SURVEY 8(d)); the test is on "
                          "max|dlambda| / max|lambda|"}
    out["dlambda_vs_1gpu"] = dl if world > 1 else None
    ok = mx.value <= tol and g.value <= tol and o.value <= tol
    if dl is not None:
        ok = ok and dl["max_abs_over_max_lambda"] <= 1e-10
    out["pass"] = bool(ok)
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # keep rank 0's stdout to the one JSON line: NCCL's version banner / warnings go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    from eigenkernel_b200 import dist as ekdist
    from eigenkernel_b200.device import Context

    n, K, W = args.n, args.steps, args.warmup
    ctx = Context(local)  # fails loudly without libekb200.so / a GPU: there is no CPU fallback
    lib, h = ctx.lib, ctx.h
    tuning = {}
    for kv in filter(None, os.environ.get("EKB200_BENCH_OPTIONS", "").split(",")):  # experiments only: "key=value,..."
        k, v = kv.split("=")
        ctx.set_option(k.strip(), int(v))
        tuning[k.strip()] = int(v)
    if dist is not None:
        ekdist.attach(ctx)  # NCCL communicator of the library over the torchrun ranks
    c0, kc = ekdist.local_slab(n, world, rank)
    ld = (n + 7) // 8 * 8
    dA, dB, dZ = ctx.alloc(ld * n * 8), ctx.alloc(ld * n * 8), ctx.alloc(ld * n * 8)
    dw = ctx.alloc((n + 8) * 8)

    def fill():
        ctx.call("ekb200_fill_synthetic", n, SEED, 1.0, 0, 0.0, dA, ld)
        ctx.call("ekb200_fill_synthetic", n, SEED + 1, float(n), 1, 2.0, dB, ld)

    def step():
        fill()
        info = ctx.call("ekb200_sygvd_dev", n, n, dA, ld, dB, ld, dw, dZ, ld)
        if info != 0:
            raise RuntimeError(f"ekb200_sygvd_dev: info = {info}")

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    for _ in range(W):
        step()
    peak = ctx.fp64_peak()  # after the warm-up solves: the roofline denominator must be taken at load clocks
    # DMMA.8x8x4 and DFMA have the same per-SM rate (128 FLOP/clk); a DMMA reading below the DFMA one is a clock
    # artefact of the probe (a power-capped burst), never a lower peak: the denominator is the larger of the two
    peak["dmma_tflops_raw"] = peak["dmma_tflops"]
    peak["dmma_tflops"] = max(peak["dmma_tflops"], peak["dfma_tflops"])
    # ---- timed region: exactly K steps, CUDA events on the library's stream, max over ranks
    ctx.clear_events()
    ctx.set_option("profile_gemm", 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.ekb200_num_launches(h)
    sec = ctypes.c_double()
    ctx.call("ekb200_timer_start")
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    ctx.call("ekb200_timer_stop", ctypes.byref(sec))
    wall = time.perf_counter() - t0
    barrier()
    launches = lib.ekb200_num_launches(h) - launches0
    collectives = lib.ekb200_num_collectives(h)
    clocks = sampler.stop() if rank == 0 else {}
    NF = 8
    fam_s, fam_w, fam_l = (ctypes.c_double * NF)(), (ctypes.c_double * NF)(), (ctypes.c_int64 * NF)()
    ctx.call("ekb200_kernel_profile", fam_s, fam_w, fam_l)
    ctx.set_option("profile_gemm", 0)
    gs, gf, gl = ctypes.c_double(fam_s[0]), ctypes.c_double(fam_w[0]), ctypes.c_int64(fam_l[0])
    FAM = ["gemm_engine", "panel_qr", "q2_apply", "sb2st", "gemm_batched", "nccl"]
    prof_rows = {}
    for i in range(lib.ekb200_profile_rows(h)):
        st, fm, ps, pw, pl = ctypes.c_char_p(), ctypes.c_int(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        lib.ekb200_profile_row(h, i, ctypes.byref(st), ctypes.byref(fm), ctypes.byref(ps), ctypes.byref(pw),
                               ctypes.byref(pl))
        ent = {"seconds": ps.value / K, "launches": pl.value // K}
        if pw.value > 0 and ps.value > 0:
            ent["tflops_or_tbs"] = pw.value / ps.value / 1e12
        prof_rows.setdefault(st.value.decode(), {})[FAM[fm.value] if fm.value < len(FAM) else str(fm.value)] = ent
    stage = {name: s / K for name, s, rep in ctx.events()}
    merge_flops = lib.ekb200_last_merge_flops(h)
    seconds = max(sec.value, 0.0)
    if dist is not None:
        import torch
        t = torch.tensor([seconds], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    # a quick on-device sanity check of the last solve (properties only; parity lives in tests/)
    wv = np.zeros(n)
    ctx.call("ekb200_d2h", wv.ctypes.data, dw, n * 8)
    assert np.all(np.isfinite(wv)) and np.all(np.diff(wv) >= 0), "eigenvalues not ascending/finite"
    acc = None
    if not args.no_check:
        acc = acceptance(ctx, n, ld, dA, dB, dZ, dw, fill, world, rank, local, dist, wv)
        barrier()

    # ---- e2e: host buffers through the reference-facing entry point
    e2e = None
    need = world * (2 * n * n * 8 + n * max(kc, 1) * 8)
    avail = mem_available_bytes()
    if dist is not None:  # one decision for all ranks
        import torch
        t = torch.tensor([float(avail if avail is not None else 1e18)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        avail = float(t.item())
    e2e_skip = None
    if not args.no_e2e and avail is not None and need * 1.3 > avail:
        e2e_skip = f"host needs {need / 1e9:.0f} GB of pinned buffers, {avail / 1e9:.0f} GB available"
    if not args.no_e2e and e2e_skip is None:
        hp = [ctypes.c_void_p() for _ in range(3)]
        for p, nb in zip(hp, (n * n * 8, n * n * 8, n * max(kc, 1) * 8)):
            ctx.call("ekb200_host_alloc", nb, ctypes.byref(p))
        hw = np.zeros(n)
        fill()
        ctx.call("ekb200_d2h_matrix", hp[0], n, dA, ld, n, n)
        ctx.call("ekb200_d2h_matrix", hp[1], n, dB, ld, n, n)
        for p in (dA, dB, dZ):
            ctx.free(p)
        dA = dB = dZ = None

        def e2e_step():
            info = ctx.call("ekb200_sygvd", n, n, hp[0], n, hp[1], n, hw.ctypes.data, hp[2], n)
            if info != 0:
                raise RuntimeError(f"ekb200_sygvd: info = {info}")

        e2e_step()  # kernels are warm from the device-resident phase; one pass warms the staging path
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            e2e_step()
        e_wall = time.perf_counter() - t0
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([e_wall], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_wall = float(t.item())
        assert np.all(np.isfinite(hw)) and np.all(np.diff(hw) >= 0)
        e2e = {"value": K * canonical_flops(n) / e_wall / 1e12, "unit": UNIT,
               "h2d_bytes_per_step": 2 * n * n * 8, "d2h_bytes_per_step": n * n * 8 + world * n * 8,
               "seconds_per_step": e_wall / K, "host_buffers": "pinned",
               "note": "every rank holds the replicated host A and B (as the reference hands every rank the replicated "
                       "COO) but uploads only its n/P block of columns of each; the blocks are all-gathered over NVLink; "
                       "every rank downloads all eigenvalues and its column slab of the eigenvectors"
                       if world > 1 else "single rank"}
        for p in hp:
            ctx.call("ekb200_host_free", p)

    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family: the DMMA GEMM engine (tensor pipe, FP64)
    # traffic: DRAM bytes of ONE launch of the family's main kernel from the committed `ncu --set full` capture (the
    # 8192^3 product: 2 x 0.54 GB of operands + 0.54 GB of C algorithmically); null when the summary is not there
    traffic, traffic_note = None, None
    try:
        rd = wr = None
        tpath = os.path.join(ROOT, "profiles", "r02_gemm_raster_8192_ncu.txt")
        for ln in open(tpath):
            w_ = ln.split()
            if "dram__bytes_read.sum" in ln:
                rd = float(w_[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[w_[3]]
            if "dram__bytes_write.sum" in ln:
                wr = float(w_[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[w_[3]]
        if rd is not None and wr is not None:
            traffic = rd + wr
            traffic_note = ("dram__bytes_read + dram__bytes_write of one gemm_bulk_kernel launch, 8192^3 FP64, from "
                            "profiles/r02_gemm_raster_8192_ncu.txt (ncu --set full); algorithmic bytes of that launch "
                            "1.61e9 (A, B, C once): with the L2-aware tile order the operands are re-read about 6x "
                            "(15x before it, profiles/r02_gemm_bulk_8192_ncu.txt), at 3 % of the DRAM peak -- the "
                            "kernel is tensor-bound (95 % pipe-active)")
    except Exception:
        pass
    gemm_share = gs.value / max(sec.value, 1e-30)
    roof = {
        "bound": "tensor", "kernel": "ekb::gemm_bulk_kernel<*> + ekb::gemm_kernel<*> (DMMA.8x8x4 engine: TMA-fed warp-specialised kernel "
                                    "and the LDGSTS kernel for batched / split-K / odd shapes)",
        "achieved": gf.value / max(gs.value, 1e-30) / 1e12, "peak": peak["dmma_tflops"], "unit": "TFLOP/s",
        "frac": gf.value / max(gs.value, 1e-30) / 1e12 / peak["dmma_tflops"], "traffic": traffic,
        "traffic_note": traffic_note,
        "launches": int(gl.value), "seconds_in_kernel_per_step": gs.value / K, "share_of_step": gemm_share,
        "peak_source": "FP64 DMMA issue-rate microbenchmark run by this process (ekb200_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 figure",
    }
    # per-stage view (canonical FLOPs / bytes of SURVEY.md 8(d))
    n3 = float(n) ** 3
    can = {"reduce_generalized_b200:potrf": n3 / 3, "reduce_generalized_b200:sygst": n3,
           "eigen_solver_b200:sy2sb": 4 * n3 / 3, "eigen_solver_b200:stedc": merge_flops,
           "eigen_solver_b200:ormtr_sb2st": 2 * n3, "eigen_solver_b200:ormtr_sy2sb": 2 * n3,
           "recovery_generalized_b200": n3}
    stages = {}
    for name, s in stage.items():
        ent = {"seconds": s}
        if name in can and s > 0:
            ent["tflops"] = can[name] / s / 1e12  # canonical FLOPs of the whole stage / this rank's stage seconds
            ent["frac_of_fp64_peak"] = ent["tflops"] / (peak["dmma_tflops"] * world)  # of the aggregate peak
        if name == "eigen_solver_b200:sb2st" and s > 0:
            band = lib.ekb200_get_band(h)
            ent["gbs_effective"] = 12.0 * band * n * n / s / 1e9
        stages[name] = ent

    cpu = None
    if not args.no_cpu:
        cores = cpu_threads()
        dt, tm = cpu_solve_once(args.cpu_n)
        table = None
        if not args.no_cpu_table:
            table = cpu_rate_table([x for x in args.cpu_sizes if x != args.cpu_n], cores)
            table["sizes"].append({"n": args.cpu_n, "seconds": dt, "tflops": canonical_flops(args.cpu_n) / dt / 1e12,
                                   "stage_seconds": tm})
            table["sizes"].sort(key=lambda r: r["n"])
            big = table["sizes"][-1]
            table["extrapolated_seconds_n32768"] = big["seconds"] * (32768.0 / big["n"]) ** 3
        cpu = {"value": canonical_flops(args.cpu_n) / dt / 1e12, "unit": UNIT, "cores": cores, "kind": "port",
               "rate_table": table,
               "sample": f"oracle port (serial-LAPACK twins of the reference's call sequence, OpenBLAS, {cores} "
                         f"threads) on n={args.cpu_n} of the same generator, {dt:.1f} s; the reference itself "
                         f"(Fortran+MPI+ScaLAPACK) cannot be built in this image",
               "stage_seconds": tm}
    line = {
        "metric": METRIC, "value": K * canonical_flops(n) / seconds / 1e12, "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": seconds / K * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "band": lib.ekb200_get_band(h), "canonical_flops_per_step":
                   canonical_flops(n), "l2": "inputs (2 x %.1f GB) exceed the 126 MB L2" % (n * n * 8 / 1e9),
                   "library_options": tuning or "defaults",
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} ranks, one problem: eigenvector column slabs (D&C top merge, Q2, Q1, trtrs), sharded "
                   f"sygst + dense-to-band (NCCL), replicated potrf / bulge chasing / lower D&C levels"},
        "seconds_per_solve": seconds / K, "wall_seconds_per_solve": wall / K,
        "clocks": clocks, "e2e": e2e if e2e is not None else ({"skipped": e2e_skip} if e2e_skip else None),
        "gpu_launches": int(launches), "nccl_collectives": int(collectives), "roofline": roof, "stages": stages,
        "kernel_profile": prof_rows, "fp64_peak_measured": peak, "cpu_baseline": cpu, "acceptance": acc,
    }
    emit(line)
    if acc is not None and not acc["pass"]:
        raise AssertionError(f"acceptance check failed: {acc}")
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--cpu-n", type=int, default=6144, dest="cpu_n")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", dest="no_check",
                    help="skip the residual / orthogonality / eigenvalue acceptance check of the last timed solve")
    ap.add_argument("--cpu-sizes", type=lambda t: [int(x) for x in t.split(",")], default=[4096, 6144, 8192],
                    dest="cpu_sizes", help="orders n at which the CPU oracle port is also timed once (rate table)")
    ap.add_argument("--no-cpu-table", action="store_true", dest="no_cpu_table")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
