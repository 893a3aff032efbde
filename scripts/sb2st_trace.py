"""Decode the phase timestamps written by libekb200 when EKB200_SB2ST_TRACE=<file> is set (development aid for the
bulge-chasing kernel): per task of one CTA, clock64() of lane 0 of warp 0 (slot 0) and warp 1 (slot 1) at
  0 after the opening barrier | 1 L block done | 2 dependency seen | 3 loads consumed | 4 before the mid barrier |
  5 after it | 6 updates + stores issued | 7 (slot 0) next reflector formed.
Prints the median clocks of every interval over the steady-state tasks.  Usage: python scripts/sb2st_trace.py <file>"""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 2, 8)
ok = (a[:, 1, 0] > 0) & (a[:, 1, 6] > 0)
a = a[ok]
print("tasks traced:", len(a))
names = ["S1", "Ldone", "dep", "loaded", "preS2", "postS2", "updated", "refl"]
for slot in (0, 1):
    d = np.diff(a[:, slot, :7], axis=1)
    print(f"slot {slot} median clocks:", {f"{names[i]}->{names[i + 1]}": int(np.median(d[:, i])) for i in range(6)})
    print(f"slot {slot} p90    clocks:", {f"{names[i]}->{names[i + 1]}": int(np.percentile(d[:, i], 90)) for i in range(6)})
r = a[:, 0, 7]
m = r > 0
print("reflector done - postS2 (slot 1):", int(np.median(r[m] - a[m, 1, 5])), " refl - updated(slot 0):",
      int(np.median(r[m] - a[m, 0, 6])))
nxt = a[1:, 1, 0] - a[:-1, 1, 0]
same = nxt > 0
print("task period (S1 -> next S1), median / p10 / p90:", int(np.median(nxt[same])), int(np.percentile(nxt[same], 10)),
      int(np.percentile(nxt[same], 90)))
print("updated -> next S1 (slot 1):", int(np.median(a[1:, 1, 0][same] - a[:-1, 1, 6][same])))
