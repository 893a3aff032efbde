#!/bin/bash
# Round 2, GPU call 4: bulge chasing v3 (batched loads, deferred D stores, pipelined polling), GEMM shape probe.
set -u
mkdir -p gpurun_out
O=gpurun_out
export EKB_SB2ST_VARIANTS="1:8:0:0,1:8:1:0,1:16:0:0,1:16:1:0"
timeout -s KILL 120 python scripts/sb2st_probe.py 2048 8192 > $O/r02_sb2st_probe3_small.jsonl 2> $O/r02_sb2st_probe3_small.err
echo "probe small rc=$?"; cut -c1-200 $O/r02_sb2st_probe3_small.jsonl; tail -3 $O/r02_sb2st_probe3_small.err
timeout -s KILL 120 python scripts/sb2st_probe.py 32768 > $O/r02_sb2st_probe3_32768.jsonl 2> $O/r02_sb2st_probe3_32768.err
echo "probe big rc=$?"; cut -c1-200 $O/r02_sb2st_probe3_32768.jsonl; tail -3 $O/r02_sb2st_probe3_32768.err
for v in 1:8:1:0 1:16:1:0; do
  EKB_SB2ST_VARIANTS=$v EKB_SB2ST_REPS=1 EKB200_SB2ST_TRACE=$O/r02_trace3_${v//:/_}.bin timeout -s KILL 60 python scripts/sb2st_probe.py 8192 > /dev/null 2>&1
  echo "== trace $v"; python scripts/sb2st_trace.py $O/r02_trace3_${v//:/_}.bin
done
unset EKB_SB2ST_VARIANTS
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_backtransform.py tests/test_gpu_stages.py tests/test_gpu_solve.py -x -q 2>&1 | tail -5
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes.jsonl 2> $O/r02_gemm_shapes.err
cat $O/r02_gemm_shapes.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes.err
