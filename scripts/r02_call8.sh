#!/bin/bash
# Round 2, GPU call 8: look-ahead with the software grid barrier; GEMM pipeline geometry BK=32 x 3 stages.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q 2>&1 | tail -3
EKB200_BENCH_OPTIONS="sy2sb_lookahead=1" timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_la1b.json 2> $O/r02_bench_la1b.err
echo "bench lookahead=1 rc=$?"; python scripts/show_bench.py $O/r02_bench_la1b.json 2>&1 | grep -E "==|sy2sb|acceptance"; tail -3 $O/r02_bench_la1b.err
EKB200_LIB=$PWD/eigenkernel_b200/libekb200_bk32s3.so timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes_bk32s3.jsonl 2> $O/r02_gemm_shapes_bk32s3.err
cat $O/r02_gemm_shapes_bk32s3.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes_bk32s3.err
