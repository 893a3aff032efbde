#!/bin/bash
# Round 2, GPU call 20: bulge chasing with two 9-warp CTAs per SM (112 registers).
set -u
mkdir -p gpurun_out
O=gpurun_out
export EKB_SB2ST_VARIANTS="1:8:1:0,1:8:1:1,1:8:0:0,1:16:1:0"
timeout -s KILL 200 python scripts/sb2st_probe.py 8192 32768 > $O/r02_sb2st_probe5.jsonl 2> $O/r02_sb2st_probe5.err
echo "probe rc=$?"; cut -c1-210 $O/r02_sb2st_probe5.jsonl; tail -3 $O/r02_sb2st_probe5.err
unset EKB_SB2ST_VARIANTS
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_solve.py tests/test_gpu_zz_select.py -x -q 2>&1 | tail -3
timeout -s KILL 500 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_sb2st112.json 2> $O/r02_bench_sb2st112.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_sb2st112.json 2>&1 | grep -E "==|sb2st|acceptance"; tail -3 $O/r02_bench_sb2st112.err
