#!/bin/bash
# Round 2, GPU call 29: in-place trsm leaves (no copy launch), conditional tile order -- tests + bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > $O/r02_bench_inplace_leaf.json 2> $O/bench_ipl.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_inplace_leaf.json 2>&1 | grep -vE "^\s+\[" | head -16
