#!/bin/bash
# Round 2, GPU call 5: bulge chasing v4 (immediate band offsets, lean reductions), GEMM with batched C loads.
set -u
mkdir -p gpurun_out
O=gpurun_out
export EKB_SB2ST_VARIANTS="1:8:0:0,1:8:1:0,1:16:1:0"
timeout -s KILL 120 python scripts/sb2st_probe.py 2048 8192 32768 > $O/r02_sb2st_probe4.jsonl 2> $O/r02_sb2st_probe4.err
echo "probe rc=$?"; cut -c1-200 $O/r02_sb2st_probe4.jsonl; tail -3 $O/r02_sb2st_probe4.err
for v in 1:8:1:0; do
  EKB_SB2ST_VARIANTS=$v EKB_SB2ST_REPS=1 EKB200_SB2ST_TRACE=$O/r02_trace4_${v//:/_}.bin timeout -s KILL 60 python scripts/sb2st_probe.py 8192 > /dev/null 2>&1
  echo "== trace $v"; python scripts/sb2st_trace.py $O/r02_trace4_${v//:/_}.bin
done
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes2.jsonl 2> $O/r02_gemm_shapes2.err
cat $O/r02_gemm_shapes2.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes2.err
unset EKB_SB2ST_VARIANTS
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_backtransform.py tests/test_gpu_stages.py tests/test_gpu_solve.py -x -q 2>&1 | tail -5
timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_quick3.json 2> $O/r02_bench_quick3.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_quick3.json 2>&1 | tail -30; tail -3 $O/r02_bench_quick3.err
NCU="ncu --set full --clock-control none --import-source on -f"
EKB_SB2ST_VARIANTS=1:8:1:0 EKB_SB2ST_REPS=1 timeout -s KILL 200 $NCU -k regex:sb2st_reg_kernel -c 1 -o $O/r02_sb2st_reg_v4_8192 python scripts/sb2st_probe.py 8192 > $O/r02_ncu_sb2st_v4.log 2>&1
