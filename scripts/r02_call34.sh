#!/bin/bash
# Round 2, GPU call 34: sanity of the layout.h refactor of the GEMM host decisions -- stage / two-stage / solve tests, one bench step.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -2
timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r02_bench_layout_refactor.json 2> $O/bench_lr.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_layout_refactor.json 2>&1 | grep -E "value=|roofline"
