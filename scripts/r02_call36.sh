#!/bin/bash
# Round 2, GPU call 36: ncu --set full of the final shared-memory panel QR (one L2 round trip per column), m ~ 16000.
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s KILL 150 $NCU -k regex:panel_qr_smem_kernel --launch-skip 4 -c 1 -o $O/r02_panelqr_smem_final_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_qrs_final.log 2>&1
echo "rc=$?"; tail -2 $O/r02_ncu_qrs_final.log
