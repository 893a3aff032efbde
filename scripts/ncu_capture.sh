#!/bin/bash
# Round-1 ncu evidence (single GPU).  Writes .ncu-rep files and the launch list under gpurun_out/.
set -u
NCU="ncu --set full --clock-control none --import-source on -f"
N=${1:-8192}
timeout 200 $NCU -k regex:gemm_kernel -c 1 -o gpurun_out/r01b_gemm_${N} python scripts/ncu_target.py $N gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 200 $NCU -k regex:sb2st_kernel -c 1 -o gpurun_out/r01b_sb2st_${N} python scripts/ncu_target.py $N stages > gpurun_out/ncu_sb2st.log 2>&1
timeout 200 $NCU -k regex:q2_apply_kernel -c 1 -o gpurun_out/r01b_q2_${N} python scripts/ncu_target.py $N stages > gpurun_out/ncu_q2.log 2>&1
timeout 200 $NCU -k regex:panel_qr_kernel --launch-skip 4 -c 1 -o gpurun_out/r01b_panelqr_${N} python scripts/ncu_target.py $N stages > gpurun_out/ncu_panelqr.log 2>&1
# launch list of the bench command at a size ncu's per-launch overhead allows
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01b_launches_n4096.csv \
  python bench.py --n 4096 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01b_launches_n4096.csv
wc -l gpurun_out/r01b_launches_n4096.csv
