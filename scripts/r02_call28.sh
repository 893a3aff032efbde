#!/bin/bash
# Round 2, GPU call 28: per-shape trace of the engine GEMMs of one n = 32768 solve (EKB200_GEMM_TRACE).
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/gemm_trace.txt
timeout -s KILL 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -2
EKB200_GEMM_TRACE=$O/gemm_trace_all.txt timeout -s KILL 300 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > $O/r02_bench_trace.json 2> $O/bench_trace.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_trace.json 2>&1 | grep -E "value="
python scripts/gemm_trace_summary.py $O/gemm_trace_all.txt 12 > $O/r02_gemm_trace_summary.txt; head -80 $O/r02_gemm_trace_summary.txt
