#!/bin/bash
# Round 2, GPU call 14 (8 B200): config 5 (n = 65536, lowest 6554 pairs) sharded over 8 ranks, multi-rank parity
# (2 and 4 ranks), the headline bench on 8 ranks with the acceptance block.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29650 scripts/config5_dist.py --solo > $O/r02_config5_8gpu.json 2> $O/r02_config5_8gpu.err
echo "config5 rc=$?"; tail -1 $O/r02_config5_8gpu.json | cut -c1-1500; tail -3 $O/r02_config5_8gpu.err
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29651 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu > $O/r02_bench_p8.json 2> $O/r02_bench_p8.err
echo "bench8 rc=$?"; python scripts/show_bench.py $O/r02_bench_p8.json 2>&1 | tail -40; tail -3 $O/r02_bench_p8.err
timeout -s KILL 500 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zzz_dist_select.py -x -q -s 2>&1 | grep -E "dist_check\] P|DIST_CHECK|passed|failed|Error" > $O/r02_dist_check_p2_p4.log
tail -12 $O/r02_dist_check_p2_p4.log
timeout -s KILL 200 $TR --nproc-per-node 8 --master-port 29652 tests/dist_worker.py 2>&1 | grep -E "dist_check\] P|DIST_CHECK|Error" > $O/r02_dist_check_p8.log
tail -7 $O/r02_dist_check_p8.log
