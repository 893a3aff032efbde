#!/bin/bash
# Round 2, GPU call 35: config 5 (n = 65536, lowest 6554 pairs) on ONE B200 with the final code.
set -u
mkdir -p gpurun_out
O=gpurun_out
EKB_SELECT_METHODS=0 EKB_SELECT_K=6554 timeout -s KILL 200 python scripts/select_probe.py select 65536 > $O/r02_config5_1gpu_final.json 2> $O/r02_config5_1gpu_final.err
echo "rc=$?"; tail -1 $O/r02_config5_1gpu_final.json | cut -c1-700; tail -2 $O/r02_config5_1gpu_final.err
