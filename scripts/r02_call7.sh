#!/bin/bash
# Round 2, GPU call 7: dense-to-band panel look-ahead.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for la in 1 0; do
  EKB200_BENCH_OPTIONS="sy2sb_lookahead=$la" timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_la$la.json 2> $O/r02_bench_la$la.err
  echo "bench lookahead=$la rc=$?"; python scripts/show_bench.py $O/r02_bench_la$la.json 2>&1 | grep -E "==|sy2sb|acceptance"; tail -3 $O/r02_bench_la$la.err
done
