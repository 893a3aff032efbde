"""Single-GPU probe of the bulge-chasing stage alone (ekb200_sb2st on a random symmetric band matrix of order n,
half bandwidth 64): seconds per variant (CUDA events on the library's stream) and a cross-check of the resulting
tridiagonal matrices through their eigenvalues (all n of them by the library's own bisection kernel).
Usage: python scripts/sb2st_probe.py <n> [<n> ...]      (also the ncu target for sb2st_reg_kernel)"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigenkernel_b200.device import Context  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [8192]
# variant:warps[:reflector_warp[:ctas_per_sm]]
variants = [tuple(int(y) for y in x.split(":")) for x in os.environ.get("EKB_SB2ST_VARIANTS", "0:8,1:8,1:16").split(",")]
variants = [tuple(list(v) + [1, 0][len(v) - 2:]) for v in variants]
reps = int(os.environ.get("EKB_SB2ST_REPS", "2"))
ctx = Context(0)
b = 64
for n in sizes:
    rng = np.random.default_rng(n)
    AB = np.zeros((2 * b, n), order="F")
    AB[: b + 1, :] = rng.standard_normal((b + 1, n))
    for d in range(1, b + 1):
        AB[d, n - d:] = 0.0
    ld = (n + 7) // 8 * 8
    ntm = ctx.lib.ekb200_sb2st_max_tasks(ctx.h, n)
    dAB = ctx.alloc(2 * b * n * 8)
    dV2, dTAU = ctx.alloc(ld * n * 8), ctx.alloc(ntm * n * 8)
    dd, de = ctx.alloc((n + 8) * 8), ctx.alloc((n + 8) * 8)
    dw, dZ = ctx.alloc((n + 8) * 8), ctx.alloc(ld * 8)
    ref = None
    for var, warps, rwarp, cps in variants:
        ctx.set_option("sb2st_variant", var)
        ctx.set_option("sb2st_warps", warps)
        ctx.set_option("sb2st_rwarp", rwarp)
        ctx.set_option("sb2st_cps", cps)
        best = 1e30
        for _ in range(reps):
            ctx.call("ekb200_h2d", dAB, AB.ctypes.data, AB.nbytes)
            ctx.call("ekb200_timer_start")
            info = ctx.call("ekb200_sb2st", n, dAB, 2 * b, dV2, ld, dTAU, ntm, dd, de)
            sec = ctypes.c_double()
            ctx.call("ekb200_timer_stop", ctypes.byref(sec))
            best = min(best, sec.value)
        d, e = np.zeros(n), np.zeros(n)
        ctx.call("ekb200_d2h", d.ctypes.data, dd, n * 8)
        ctx.call("ekb200_d2h", e.ctypes.data, de, (n - 1) * 8)
        ctx.call("ekb200_stebz_stein", n, 1, dd, de, dw, dZ, ld)
        w = np.zeros(n)
        ctx.call("ekb200_d2h", w.ctypes.data, dw, n * 8)
        out = {"n": n, "variant": var, "warps": warps, "rwarp": rwarp, "cps": cps, "info": info, "seconds": best,
               "us_per_sweep": best / max(n - 2, 1) * 1e6, "gbs_effective": 12.0 * b * n * n / best / 1e9,
               "trace_err": abs(d.sum() - AB[0].sum()) / np.abs(AB[0]).sum(),
               "fro_err": abs(np.sqrt((d ** 2).sum() + 2 * (e ** 2).sum())
                              - np.sqrt((AB[0] ** 2).sum() + 2 * (AB[1:] ** 2).sum())) / np.sqrt((AB ** 2).sum())}
        if ref is None:
            ref = w
        else:
            out["max_dlambda_vs_first_over_norm"] = float(np.abs(w - ref).max() / np.abs(ref).max())
        print(json.dumps(out), flush=True)
    for p in (dAB, dV2, dTAU, dd, de, dw, dZ):
        ctx.free(p)
ctx.close()
