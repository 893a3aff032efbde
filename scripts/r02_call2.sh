#!/bin/bash
# Round 2, GPU call 2: first run of the register-resident bulge-chasing kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 90 python scripts/sb2st_probe.py 700 2048 8192 > $O/r02_sb2st_probe_small.jsonl 2> $O/r02_sb2st_probe_small.err
echo "probe small rc=$?"; cat $O/r02_sb2st_probe_small.jsonl; tail -3 $O/r02_sb2st_probe_small.err
timeout -s KILL 300 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_backtransform.py -x -q 2>&1 | tail -5
timeout -s KILL 120 python scripts/sb2st_probe.py 32768 > $O/r02_sb2st_probe_32768.jsonl 2> $O/r02_sb2st_probe_32768.err
echo "probe big rc=$?"; cat $O/r02_sb2st_probe_32768.jsonl; tail -3 $O/r02_sb2st_probe_32768.err
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
NCU="ncu --set full --clock-control none --import-source on -f"
EKB_SB2ST_VARIANTS=1:8 EKB_SB2ST_REPS=1 timeout -s KILL 200 $NCU -k regex:sb2st_reg_kernel -c 1 -o $O/r02_sb2st_reg_16384 python scripts/sb2st_probe.py 16384 > $O/r02_ncu_sb2st_reg.log 2>&1
timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_quick.json 2> $O/r02_bench_quick.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_quick.json 2>&1 | tail -30; tail -3 $O/r02_bench_quick.err
