#!/bin/bash
# Round 2, GPU call 25: ncu --set full of the Q2 walk (14 thin warps + producer), n = 16384, k = 16384.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:q2_apply_kernel -c 1 -f -o $O/r02_q2_14w python scripts/q2_slab_probe.py 16384 16384 0 > $O/ncu_q2.log 2>&1
echo "ncu rc=$?"; tail -3 $O/ncu_q2.log
ls -la $O/r02_q2_14w.ncu-rep
