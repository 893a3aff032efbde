#!/bin/bash
# Round 2, GPU calls 23-24: Q2 with 14 thin warps (8 columns each) + producer against 7 warps of 16 columns (call 24: dependent DMMAs kept apart).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py -x -q -m gpu 2>&1 | tail -3
for v in 0 112; do
EKB200_BENCH_OPTIONS="q2_kc=$v" timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > $O/r02_bench_q2kc_$v.json 2> $O/bench_q$v.err
echo "q2_kc=$v rc=$?"; python scripts/show_bench.py $O/r02_bench_q2kc_$v.json 2>&1 | grep -E "value=|ormtr_sb2st|acceptance"; tail -2 $O/bench_q$v.err
done
