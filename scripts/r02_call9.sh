#!/bin/bash
# Round 2, GPU call 9: look-ahead with residency gate; GEMM with two pipeline geometries.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes3.jsonl 2> $O/r02_gemm_shapes3.err
cat $O/r02_gemm_shapes3.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes3.err
for la in 1 0; do
  EKB200_BENCH_OPTIONS="sy2sb_lookahead=$la" timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_la${la}c.json 2> $O/r02_bench_la${la}c.err
  echo "bench lookahead=$la rc=$?"; python scripts/show_bench.py $O/r02_bench_la${la}c.json 2>&1 | grep -vE "^\s+\[.*(potrf|sygst|stedc|recovery)"; tail -3 $O/r02_bench_la${la}c.err
done
