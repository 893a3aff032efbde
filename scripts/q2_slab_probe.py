"""GPU probe: apply_q2 on a narrow column slab (what one rank of an 8-GPU solve sees) -- library's choice of
columns per warp (kc = 0: 8 per warp for narrow slabs) against the forced 16-per-warp kernel (kc = 64)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigenkernel_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ks = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "2048,4096")]
ctx = Context(0)
b = 64
ld = (n + 7) // 8 * 8
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
sec = ctypes.c_double()
for k in ks:
    dZ = ctx.alloc(ld * k * 8)
    for kc in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0,64".split(","))]:
        ctx.set_option("q2_kc", kc)
        best = 1e9
        for rep in range(3):
            ctx.call("ekb200_fill_synthetic", min(n, k), 7, 1.0, 0, 0.0, dZ, ld)
            ctx.call("ekb200_timer_start")
            ctx.call("ekb200_apply_q2", n, k, dV2, ld, dTAU, ntm, dZ, ld)
            ctx.call("ekb200_timer_stop", ctypes.byref(sec))
            best = min(best, sec.value)
        print(f"apply_q2 n={n} k={k} kc={kc}: best {best * 1e3:.1f} ms  {2 * n * n * k / best / 1e12:.2f} TF/s", flush=True)
    ctx.free(dZ)
