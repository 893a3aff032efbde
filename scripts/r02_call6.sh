#!/bin/bash
# Round 2, GPU call 6 (2 B200): multi-rank parity with the new kernels, block-cyclic delivery, sharded upload; bench N=2.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
timeout -s KILL 500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zzz_dist_select.py -x -q -s 2>&1 | grep -E "dist_check|DIST_CHECK|passed|failed|Error|error" | tail -30
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > $O/r02_bench_p2_quick.json 2> $O/r02_bench_p2_quick.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_p2_quick.json 2>&1 | tail -32; tail -5 $O/r02_bench_p2_quick.err
