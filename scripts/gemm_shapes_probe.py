"""Single-GPU probe of the DMMA GEMM engine at the shapes the eigensolve actually uses (rank-2b trailing updates,
N = 64 panels products, WY applications), with and without accumulation into C: TFLOP/s by CUDA events on the library
stream (operands are zero-filled device buffers: timing only).  Usage: python scripts/gemm_shapes_probe.py"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigenkernel_b200.device import Context  # noqa: E402

ctx = Context(0)
print(json.dumps({"peak": ctx.fp64_peak()}), flush=True)
shapes = [
    # (ta, tb, m, n, k, alpha, beta, label)
    ("N", "T", 16000, 16000, 128, -1.0, 1.0, "rank-128 update, C += (SYR2K shape, full square)"),
    ("N", "T", 16000, 16000, 128, 1.0, 0.0, "rank-128 product, C ="),
    ("N", "T", 16000, 16000, 512, -1.0, 1.0, "rank-512 update, C +="),
    ("N", "T", 16000, 16000, 512, 1.0, 0.0, "rank-512 product, C ="),
    ("N", "T", 32000, 32000, 128, -1.0, 1.0, "rank-128 update m=32000, C +="),
    ("N", "N", 16000, 64, 16000, 1.0, 0.0, "panel product N=64 (SYMM shape)"),
    ("N", "N", 32000, 64, 32000, 1.0, 0.0, "panel product N=64, m=32000"),
    ("T", "N", 512, 16384, 16000, 1.0, 0.0, "W = V^T Z (Q1 step 1)"),
    ("N", "N", 16000, 16384, 512, -1.0, 1.0, "Z -= V W (Q1 step 3)"),
    ("N", "N", 8192, 8192, 8192, 1.0, 0.0, "square 8192"),
    ("N", "N", 8192, 8192, 8192, -1.0, 1.0, "square 8192, C +="),
    ("N", "N", 2048, 2048, 2048, -1.0, 1.0, "square 2048, C +="),
    ("N", "N", 1024, 1024, 1024, -1.0, 1.0, "square 1024, C +="),
]
for ta, tb, m, n, k, alpha, beta, label in shapes:
    A = ctx.matrix(k if ta == "T" else m, m if ta == "T" else k)
    B = ctx.matrix(n if tb == "T" else k, k if tb == "T" else n)
    C = ctx.matrix(m, n)
    best = 1e30
    for it in range(4):
        ctx.call("ekb200_timer_start")
        ctx.dgemm(ta, tb, alpha, A, B, beta, C)
        sec = ctypes.c_double()
        ctx.call("ekb200_timer_stop", ctypes.byref(sec))
        if it:
            best = min(best, sec.value)
    print(json.dumps({"shape": [ta, tb, m, n, k], "alpha": alpha, "beta": beta, "label": label, "ms": best * 1e3,
                      "tflops": 2.0 * m * n * k / best / 1e12}), flush=True)
    for M in (A, B, C):
        M.free()
ctx.close()
