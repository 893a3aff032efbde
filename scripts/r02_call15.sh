#!/bin/bash
# Round 2, GPU call 15: Q2 with a producer warp and a ring of three image buffers.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_backtransform.py tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q 2>&1 | tail -4
timeout -s KILL 300 python scripts/q2_slab_probe.py 16384 16384,4096,2048 0,64,80,96,112,128,1004,1008,1012 > $O/r02_q2_slab_ring.txt 2>&1
cat $O/r02_q2_slab_ring.txt | tail -30
timeout -s KILL 500 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_q2ring.json 2> $O/r02_bench_q2ring.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_q2ring.json 2>&1 | grep -E "==|ormtr_sb2st|q2_apply|acceptance"; tail -3 $O/r02_bench_q2ring.err
