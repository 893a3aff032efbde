#!/bin/bash
# Round 2, GPU call 1 (single B200): config 5 on one GPU, headline app run with checks, ncu of the round-1 kernels
# that were never captured.  Everything lands in gpurun_out/r02_*; ncu runs are profiles only.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > $O/r02_smi.txt
EKB_SELECT_METHODS=0 EKB_SELECT_K=6554 timeout 400 python scripts/select_probe.py select 65536 > $O/r02_select_65536.json 2> $O/r02_select_65536.err
echo "select rc=$?"
timeout 60 python scripts/select_probe.py stebz 65536 > $O/r02_stebz_65536.json 2> $O/r02_stebz_65536.err
mkdir -p $O/r02_app && (cd $O/r02_app && timeout 200 ../../app/bin/ekb200_app -s general_b200 -c -1 -t 1,32768 \
  synthetic:32768:20240602 synthetic:32768:20240603 > stdout.txt 2> stderr.txt; echo "app rc=$?"; tail -12 stdout.txt)
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:sb2st_kernel -c 1 -o $O/r02_sb2st_old_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_sb2st.log 2>&1
timeout 120 $NCU -k regex:"bisect_kernel|stein_kernel" -c 2 -o $O/r02_stebz_32768 python scripts/select_probe.py stebz 32768 > $O/r02_ncu_stebz.log 2>&1
timeout 150 $NCU -k regex:gemm_kernel --launch-skip 40 -c 6 -o $O/r02_sy2sb_gemms_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_sy2sb.log 2>&1
cat $O/r02_select_65536.json $O/r02_stebz_65536.json
tail -n 3 $O/r02_*.err
ls -la $O/*.ncu-rep
