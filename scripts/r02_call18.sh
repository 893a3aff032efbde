#!/bin/bash
# Round 2, GPU call 18: TMA-fed GEMM generalised (batched D&C merges, k tails, unaligned operands).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_twostage.py -x -q 2>&1 | tail -4
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 500 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_bulkb.json 2> $O/r02_bench_bulkb.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_bulkb.json 2>&1 | grep -vE "^\s+\[.*(sb2st|q2)"; tail -3 $O/r02_bench_bulkb.err
