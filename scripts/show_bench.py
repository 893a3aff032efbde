#!/usr/bin/env python
"""Print the stage table and the per-(stage, family) kernel profile of a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
    if d is None:
        print(path, ": no JSON line")
        continue
    print(f"== {path}: n_gpus={d['n_gpus']} value={d['value']:.2f} {d['unit']} s/solve={d['ms_per_step'] / 1e3:.3f} "
          f"launches={d.get('gpu_launches')} nccl={d.get('nccl_collectives')} clocks={d.get('clocks', {}).get('sm_mhz')}")
    e = d.get("e2e")
    if e and "value" in e:
        print(f"   e2e {e['value']:.2f} {e['unit']}  {e['seconds_per_step']:.3f} s")
    for k, v in d.get("stages", {}).items():
        extra = " ".join(f"{a}={b:.3f}" for a, b in v.items() if a != "seconds")
        print(f"   {k:42s} {v['seconds']:8.3f} s  {extra}")
    for st, v in d.get("kernel_profile", {}).items():
        for f, ent in v.items():
            r = f" rate={ent['tflops_or_tbs']:.2f}" if "tflops_or_tbs" in ent else ""
            print(f"      [{st}] {f}: {ent['seconds']:.3f} s, {ent['launches']} launches{r}")
    if d.get("acceptance"):
        a = d["acceptance"]
        print("   acceptance:", {k: a[k] for k in ("residual_max_over_A", "orthogonality_verifier",
                                                   "xtbx_minus_identity_fro", "tolerance", "dlambda_vs_1gpu", "pass")})
    rf = d.get("roofline")
    if rf:
        print(f"   roofline: {rf['achieved']:.2f}/{rf['peak']:.2f} {rf['unit']} frac={rf['frac']:.3f} share={rf.get('share_of_step', 0):.3f}")
