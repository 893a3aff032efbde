#!/bin/bash
# Round 2, GPU call 21 (8 B200): final numbers on 8 ranks -- headline bench and config 5.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29751 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu > $O/r02_bench_p8_final.json 2> $O/r02_bench_p8_final.err
echo "bench8 rc=$?"; python scripts/show_bench.py $O/r02_bench_p8_final.json 2>&1 | grep -vE "^\s+\["; tail -2 $O/r02_bench_p8_final.err
timeout -s KILL 300 $TR --nproc-per-node 8 --master-port 29750 scripts/config5_dist.py > $O/r02_config5_8gpu_final.json 2> $O/r02_config5_8gpu_final.err
echo "config5 rc=$?"; tail -1 $O/r02_config5_8gpu_final.json | cut -c1-900
timeout -s KILL 300 $TR --nproc-per-node 4 --master-port 29752 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r02_bench_p4_final.json 2> $O/r02_bench_p4_final.err
echo "bench4 rc=$?"; python scripts/show_bench.py $O/r02_bench_p4_final.json 2>&1 | grep -vE "^\s+\["; tail -2 $O/r02_bench_p4_final.err
