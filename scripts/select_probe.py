"""Single-GPU probe of the round-1 additions (timings with CUDA events on the library's stream; also the ncu target):
  stebz   ekb200_stebz_stein on a synthetic tridiagonal matrix of order n, lowest k = n/10 pairs
  select  whole `-n` solve (standard, synthetic) with the D&C path (select_method 1) and with bisection (2)
  inv     generalized solve with the blocked reduction (reduction 0) and the explicit-inverse variant (1)
Usage: python scripts/select_probe.py <what> <n> [<n> ...]"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigenkernel_b200.device import Context  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "stebz"
sizes = [int(x) for x in sys.argv[2:]] or [8192]
ctx = Context(0)
out = []


def timed(fn):
    ctx.call("ekb200_timer_start")
    info = fn()
    sec = ctypes.c_double()
    ctx.call("ekb200_timer_stop", ctypes.byref(sec))
    return info, sec.value


def profile():
    F = 8
    s, w, l = (ctypes.c_double * F)(), (ctypes.c_double * F)(), (ctypes.c_int64 * F)()
    ctx.call("ekb200_kernel_profile", s, w, l)
    return {"bisect_s": s[6], "sturm_steps_per_sweep": w[6], "stein_s": s[7], "stein_bytes": w[7]}


for n in sizes:
    ld = (n + 7) // 8 * 8
    if what == "stebz":
        # tridiagonal with a semicircle-like dense spectrum: the (d, e) of a random symmetric matrix are ~ N(0,1), chi_k
        rng = np.random.default_rng(n)
        d = rng.standard_normal(n)
        e = np.sqrt(rng.chisquare(np.arange(n - 1, 0, -1)) / 1.0) / np.sqrt(2.0) * np.sqrt(2.0 / 3.0)
        k = max(n // 10, 1)
        dd, de, dw, dZ = ctx.from_numpy(d), ctx.from_numpy(e), ctx.matrix(n, 1), ctx.matrix(n, k)
        ctx.set_option("profile_gemm", 1)
        for rep in range(2):
            info, sec = timed(lambda: ctx.call("ekb200_stebz_stein", n, k, dd.ptr, de.ptr, dw.ptr, dZ.ptr, dZ.ld))
            p = profile()
        ctx.set_option("profile_gemm", 0)
        w = dw.download()[:, 0]
        Z = dZ.download()
        Y = d[:, None] * Z
        Y[:-1] += e[:, None] * Z[1:]
        Y[1:] += e[:, None] * Z[:-1]
        res = np.linalg.norm(Y - Z * w[None, :k], axis=0).max() / (np.abs(d).max() + 2 * np.abs(e).max())
        orth = np.linalg.norm(Z.T @ Z - np.eye(k), "fro")
        out.append({"what": what, "n": n, "k": k, "info": info, "seconds": sec, **p, "residual_over_T": res,
                    "orth_fro": orth, "orth_tol": 1e-12 * n})
        for m in (dd, de, dw, dZ):
            m.free()
    elif what == "select":
        k = int(os.environ.get("EKB_SELECT_K", max(n // 10, 1)))
        methods = [int(x) for x in os.environ.get("EKB_SELECT_METHODS", "1,2").split(",")]
        for method in methods:
            ctx.set_option("select_method", method)
            dA, dw, dZ = ctx.matrix(n, n), ctx.matrix(n, 1), ctx.matrix(n, k)
            ctx.call("ekb200_fill_synthetic", n, 20240603, 1.0, 0, 0.0, dA.ptr, dA.ld)
            ctx.clear_events()
            try:
                info, sec = timed(lambda: ctx.call("ekb200_syevd_dev", n, k, dA.ptr, dA.ld, dw.ptr, dZ.ptr, dZ.ld))
            except Exception as exc:  # e.g. the D&C workspaces do not fit at n = 65536: record and go on
                out.append({"what": what, "n": n, "k": k, "select_method": method, "error": str(exc)})
                print(json.dumps(out[-1]), flush=True)
                for m in (dA, dw, dZ):
                    m.free()
                continue
            ev = {name: s for name, s, _ in ctx.events()}
            # check on the device: residual and orthogonality of the k pairs
            ctx.call("ekb200_fill_synthetic", n, 20240603, 1.0, 0, 0.0, dA.ptr, dA.ld)
            a, ave, mx, o = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            ctx.call("ekb200_eval_residual_norm_dev", n, k, dA.ptr, dA.ld, None, 0, dw.ptr, dZ.ptr, dZ.ld,
                     ctypes.byref(a), ctypes.byref(ave), ctypes.byref(mx))
            ctx.call("ekb200_eval_orthogonality_dev", n, 1, k, dZ.ptr, dZ.ld, None, 0, ctypes.byref(o))
            out.append({"what": what, "n": n, "k": k, "select_method": method, "info": info, "seconds": sec, "events": ev,
                        "res_max": mx.value, "orthogonality": o.value, "tol": 1e-12 * n})
            for m in (dA, dw, dZ):
                m.free()
        ctx.set_option("select_method", 0)
    elif what == "inv":
        for red in (0, 1):
            ctx.set_option("reduction", red)
            dA, dB, dw, dZ = ctx.matrix(n, n), ctx.matrix(n, n), ctx.matrix(n, 1), ctx.matrix(n, n)
            for rep in range(2):
                ctx.call("ekb200_fill_synthetic", n, 20240601, 1.0, 0, 0.0, dA.ptr, dA.ld)
                ctx.call("ekb200_fill_synthetic", n, 20240602, float(n), 1, 2.0, dB.ptr, dB.ld)
                ctx.clear_events()
                info, sec = timed(lambda: ctx.call("ekb200_sygvd_dev", n, n, dA.ptr, dA.ld, dB.ptr, dB.ld, dw.ptr, dZ.ptr, dZ.ld))
            ev = {name: s for name, s, _ in ctx.events() if name.startswith("re")}
            out.append({"what": what, "n": n, "reduction": red, "info": info, "seconds": sec, "events": ev})
            for m in (dA, dB, dw, dZ):
                m.free()
        ctx.set_option("reduction", 0)
ctx.close()
for o in out:
    print(json.dumps(o))
