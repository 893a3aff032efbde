"""GPU probe: time apply_q2 / apply_q1 alone on an n x n identity-free random Z (stage-level tuning aid)."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eigenkernel_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = Context(0)
b = 64
ld = (n + 7) // 8 * 8
dA = ctx.alloc(ld * n * 8)
ctx.call("ekb200_fill_synthetic", n, 1, 1.0, 0, 0.0, dA, ld)
dAB = ctx.alloc(2 * b * n * 8)
npan = ctx.lib.ekb200_sy2sb_num_panels(ctx.h, n)
dT = ctx.alloc(b * b * max(npan, 1) * 8)
sec = ctypes.c_double()
def timed(name, fn, flops=None):
    ctx.call("ekb200_timer_start"); fn(); ctx.call("ekb200_timer_stop", ctypes.byref(sec))
    msg = f"{name}: {sec.value*1e3:.1f} ms"
    if flops: msg += f"  {flops/sec.value/1e12:.2f} TF/s"
    print(msg, flush=True)
    return sec.value
timed("sy2sb", lambda: ctx.call("ekb200_sy2sb", n, dA, ld, dAB, 2 * b, dT), 4 * n**3 / 3)
dV2 = ctx.alloc(ld * n * 8)
ntm = ctx.lib.ekb200_sb2st_max_tasks(ctx.h, n)
dTAU = ctx.alloc(ntm * n * 8)
dd, de = ctx.alloc((n + 8) * 8), ctx.alloc((n + 8) * 8)
timed("sb2st", lambda: ctx.call("ekb200_sb2st", n, dAB, 2 * b, dV2, ld, dTAU, ntm, dd, de))
dZ = ctx.alloc(ld * n * 8)
import subprocess
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,utilization.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print("smi:", smi())
for kc in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,32,64,112".split(","))]:
    ctx.set_option("q2_kc", kc)
    best = 1e9
    for rep in range(4):
        ctx.call("ekb200_fill_synthetic", n, 7, 1.0, 0, 0.0, dZ, ld)
        ctx.call("ekb200_timer_start"); ctx.call("ekb200_apply_q2", n, n, dV2, ld, dTAU, ntm, dZ, ld); ctx.call("ekb200_timer_stop", ctypes.byref(sec))
        best = min(best, sec.value)
    print(f"apply_q2 kc={kc}: best {best*1e3:.1f} ms  {2*n**3/best/1e12:.2f} TF/s", flush=True)
print("smi:", smi())
best = 1e9
for rep in range(3):
    ctx.call("ekb200_timer_start"); ctx.call("ekb200_apply_q1", n, n, dA, ld, dT, dZ, ld); ctx.call("ekb200_timer_stop", ctypes.byref(sec))
    best = min(best, sec.value)
print(f"apply_q1: best {best*1e3:.1f} ms  {2*n**3/best/1e12:.2f} TF/s", flush=True)
