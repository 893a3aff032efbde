#!/bin/bash
# Round 2, GPU call 30 (2 B200): sharded-solve parity and the headline bench on two ranks with the final code.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29652 tests/dist_worker.py 2>&1 | grep -E "dist_check\] P|DIST_CHECK|Error" > $O/r02_dist_check_p2_final.log
echo "dist rc=$?"; tail -4 $O/r02_dist_check_p2_final.log
timeout -s KILL 400 $TR --nproc-per-node 2 --master-port 29751 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > $O/r02_bench_p2_final.json 2> $O/r02_bench_p2_final.err
echo "bench2 rc=$?"; python scripts/show_bench.py $O/r02_bench_p2_final.json 2>&1 | grep -vE "^\s+\["; tail -2 $O/r02_bench_p2_final.err
