#!/bin/bash
# Round 2, GPU call 10: shared-memory-resident panel QR.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_qrs.json 2> $O/r02_bench_qrs.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_qrs.json 2>&1 | grep -vE "^\s+\[.*(potrf|sygst|stedc|recovery)"; tail -3 $O/r02_bench_qrs.err
EKB_SELECT_METHODS=0 EKB_SELECT_K=6554 timeout 400 python scripts/select_probe.py select 65536 > $O/r02_select_65536_b.json 2> $O/r02_select_65536_b.err
cat $O/r02_select_65536_b.json | cut -c1-700
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s KILL 200 $NCU -k regex:panel_qr_smem_kernel --launch-skip 4 -c 1 -o $O/r02_panelqr_smem_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_qrs.log 2>&1
