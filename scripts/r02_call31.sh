#!/bin/bash
# Round 2, GPU call 31 (4 B200): sharded-solve parity and the headline solve on four ranks with the final code.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --nproc-per-node 4 --master-port 29652 tests/dist_worker.py 2>&1 | grep -E "dist_check\] P|DIST_CHECK|Error" > $O/r02_dist_check_p4_final.log
echo "dist rc=$?"; tail -3 $O/r02_dist_check_p4_final.log
timeout -s KILL 300 $TR --nproc-per-node 4 --master-port 29752 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r02_bench_p4_final2.json 2> $O/r02_bench_p4_final2.err
echo "bench4 rc=$?"; python scripts/show_bench.py $O/r02_bench_p4_final2.json 2>&1 | grep -vE "^\s+\[" | grep -v acceptance; tail -2 $O/r02_bench_p4_final2.err
