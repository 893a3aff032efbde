#!/bin/bash
# Round 2, GPU call 32: shared-memory panel QR with all partial-dot loads in flight -- tests + bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -2
timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > $O/r02_bench_qr_batched_loads.json 2> $O/bench_qrb.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_qr_batched_loads.json 2>&1 | grep -E "value=|sy2sb|acceptance"
