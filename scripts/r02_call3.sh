#!/bin/bash
# Round 2, GPU call 3: bulge-chasing variants (relaxed polling, dedicated reflector warp, CTAs per SM), phase trace,
# GEMM engine with accumulator-initialised epilogue.
set -u
mkdir -p gpurun_out
O=gpurun_out
export EKB_SB2ST_VARIANTS="1:8:0:0,1:8:0:1,1:8:1:0,1:16:0:0,1:16:1:0"
timeout -s KILL 120 python scripts/sb2st_probe.py 2048 8192 > $O/r02_sb2st_probe2_small.jsonl 2> $O/r02_sb2st_probe2_small.err
echo "probe small rc=$?"; cat $O/r02_sb2st_probe2_small.jsonl; tail -3 $O/r02_sb2st_probe2_small.err
timeout -s KILL 120 python scripts/sb2st_probe.py 32768 > $O/r02_sb2st_probe2_32768.jsonl 2> $O/r02_sb2st_probe2_32768.err
echo "probe big rc=$?"; cat $O/r02_sb2st_probe2_32768.jsonl; tail -3 $O/r02_sb2st_probe2_32768.err
for v in 1:8:0:0 1:8:1:0 1:16:1:0; do
  EKB_SB2ST_VARIANTS=$v EKB_SB2ST_REPS=1 EKB200_SB2ST_TRACE=$O/r02_trace_${v//:/_}.bin timeout -s KILL 60 python scripts/sb2st_probe.py 8192 > /dev/null 2>&1
  echo "== trace $v"; python scripts/sb2st_trace.py $O/r02_trace_${v//:/_}.bin
done
unset EKB_SB2ST_VARIANTS
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_quick2.json 2> $O/r02_bench_quick2.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_quick2.json 2>&1 | tail -30; tail -3 $O/r02_bench_quick2.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s KILL 300 $NCU -k regex:q2_apply_kernel -c 1 -o $O/r02_q2_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_q2.log 2>&1
ls -la $O/*.ncu-rep
