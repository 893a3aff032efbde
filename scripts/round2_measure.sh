#!/bin/bash
# Open measurements of DESIGN.md section 10, one single-GPU gpurun call (about 6 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/round2_measure.sh'
# Everything lands in gpurun_out/r02_*; nothing here is a bench value (ncu runs are profiles only).
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. config 5 on one B200 (memory-lean -n path must be picked by select_method 0 at n = 65536) + full-size probes
timeout 200 python scripts/select_probe.py select 65536 > $O/r02_select_65536.json 2> $O/r02_select_65536.err
timeout 60 python scripts/select_probe.py stebz 65536 > $O/r02_stebz_65536.json 2> $O/r02_stebz_65536.err
timeout 90 python scripts/select_probe.py inv 16384 32768 > $O/r02_inv.json 2> $O/r02_inv.err
# 2. the user-facing command on the headline workload (device-resident checks included)
mkdir -p $O/r02_app && (cd $O/r02_app && timeout 150 ../../app/bin/ekb200_app -s general_b200 -c -1 -t 1,32768 \
  synthetic:32768:20240602 synthetic:32768:20240603 > stdout.txt 2> stderr.txt; tail -8 stdout.txt)
# 3. ncu: the 7-chain bisection kernel and the inverse-iteration kernel at full size, the N = 64 SYMM of dense-to-band
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 120 $NCU -k regex:"bisect_kernel|stein_kernel" -c 2 -o $O/r02_stebz_32768 python scripts/select_probe.py stebz 32768 > $O/r02_ncu_stebz.log 2>&1
timeout 120 $NCU -k regex:gemm_kernel --launch-skip 40 -c 6 -o $O/r02_sy2sb_gemms_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_sy2sb.log 2>&1
cat $O/r02_select_65536.json $O/r02_stebz_65536.json $O/r02_inv.json
tail -n 2 $O/r02_*.err
ls -la $O/*.ncu-rep
