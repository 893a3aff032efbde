"""GPU probe: FP64 peak (DMMA / DFMA issue rate) and DGEMM throughput of the engine at a few shapes."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eigenkernel_b200.device import Context

ctx = Context(0)
out = {"peak": ctx.fp64_peak()}
print(out, flush=True)
res = []
for (ta, tb, m, n, k) in [("N","N",8192,8192,8192),("N","T",8192,8192,8192),("T","N",8192,8192,8192),
                          ("N","T",16384,16384,128),("N","N",16384,64,16384),("T","N",64,16384,16384),("N","N",16384,16384,64)]:
    A = ctx.matrix(k if ta=="T" else m, m if ta=="T" else k)
    B = ctx.matrix(n if tb=="T" else k, k if tb=="T" else n)
    C = ctx.matrix(m, n)
    ctx.call("ekb200_fill_synthetic", max(A.m, A.n), 1, 1.0, 0, 0.0, A.ptr, A.ld) if A.m == A.n else None
    for it in range(2):
        ctx.dgemm(ta, tb, 1.0, A, B, 0.0, C)
    ctx.sync()
    reps = 5
    t0 = time.perf_counter()
    for it in range(reps):
        ctx.dgemm(ta, tb, 1.0, A, B, 1.0, C)
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    tf = 2.0 * m * n * k / dt / 1e12
    res.append({"shape": [ta, tb, m, n, k], "ms": dt * 1e3, "tflops": tf})
    print(res[-1], flush=True)
    for d in (A, B, C): d.free()
out["gemm"] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_probe.json", "w"), indent=1)
