#!/bin/bash
# Round 2, GPU call 13: TMA-fed GEMM by default (+ symmetric-A panel products), full test suite, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes_bulk3.jsonl 2> $O/r02_gemm_shapes_bulk3.err
grep -E "panel product|square 8192\"" $O/r02_gemm_shapes_bulk3.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes_bulk3.err
timeout -s KILL 500 python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02_bench_bulk3.json 2> $O/r02_bench_bulk3.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_bulk3.json 2>&1; tail -3 $O/r02_bench_bulk3.err
