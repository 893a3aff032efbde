"""Single-GPU driver for ncu captures: runs each hot kernel family once at order n (default 8192) through the
C-ABI stage entry points -- dense-to-band (panel QR + engine GEMMs), bulge chasing, Q2, Q1 -- and one n^3 DGEMM.
Usage under ncu: ncu --set full -k regex:<kernel> -c 1 -o gpurun_out/<name> python scripts/ncu_target.py 8192"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigenkernel_b200.device import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
what = sys.argv[2] if len(sys.argv) > 2 else "all"
ctx = Context(0)
b = 64
ld = (n + 7) // 8 * 8
if what in ("all", "gemm"):
    A, B, C = ctx.matrix(n, n), ctx.matrix(n, n), ctx.matrix(n, n)
    ctx.call("ekb200_fill_synthetic", n, 1, 1.0, 0, 0.0, A.ptr, A.ld)
    ctx.call("ekb200_fill_synthetic", n, 2, 1.0, 0, 0.0, B.ptr, B.ld)
    ctx.dgemm("N", "N", 1.0, A, B, 0.0, C)
    ctx.sync()
    for m in (A, B, C):
        m.free()
if what in ("all", "stages"):
    dA = ctx.alloc(ld * n * 8)
    ctx.call("ekb200_fill_synthetic", n, 1, 1.0, 0, 0.0, dA, ld)
    dAB = ctx.alloc(2 * b * n * 8)
    npan = ctx.lib.ekb200_sy2sb_num_panels(ctx.h, n)
    dT = ctx.alloc(b * b * max(npan, 1) * 8)
    ctx.call("ekb200_sy2sb", n, dA, ld, dAB, 2 * b, dT)
    dV2 = ctx.alloc(ld * n * 8)
    ntm = ctx.lib.ekb200_sb2st_max_tasks(ctx.h, n)
    dTAU = ctx.alloc(ntm * n * 8)
    dd, de = ctx.alloc((n + 8) * 8), ctx.alloc((n + 8) * 8)
    ctx.call("ekb200_sb2st", n, dAB, 2 * b, dV2, ld, dTAU, ntm, dd, de)
    dZ = ctx.alloc(ld * n * 8)
    ctx.call("ekb200_fill_synthetic", n, 7, 1.0, 0, 0.0, dZ, ld)
    ctx.call("ekb200_apply_q2", n, n, dV2, ld, dTAU, ntm, dZ, ld)
    ctx.call("ekb200_apply_q1", n, n, dA, ld, dT, dZ, ld)
    ctx.sync()
ctx.close()
print("ncu_target done", n, what)
