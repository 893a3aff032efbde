#!/bin/bash
# Round 2, GPU call 22: split-K on the TMA-fed GEMM kernel chosen by the round-count model -- tests, A/B bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_twostage.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes_autosplit.jsonl 2> $O/shapes.err; tail -3 $O/shapes.err
cut -c1-200 $O/r02_gemm_shapes_autosplit.jsonl
for v in 0 1; do
EKB200_BENCH_OPTIONS="gemm_autosplit=$v" timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > $O/r02_bench_autosplit_$v.json 2> $O/bench_$v.err
echo "autosplit=$v rc=$?"; python scripts/show_bench.py $O/r02_bench_autosplit_$v.json 2>&1 | grep -vE "^\s+\[" | head -16
done
