#!/bin/bash
# Round 2, GPU call 37: panel QR with three block barriers per column (scalars derived in every thread) -- tests + bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -2
timeout -s KILL 150 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r02_bench_qr_3barriers.json 2> $O/bench_qr3.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_qr_3barriers.json 2>&1 | grep -E "value=|sy2sb\]|acceptance"
