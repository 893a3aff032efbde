#!/bin/bash
# Round 2, GPU call 11: TMA-fed warp-specialised GEMM kernel (option gemm_bulk=1): correctness then speed.
set -u
mkdir -p gpurun_out
O=gpurun_out
export EKB200_OPTIONS="gemm_bulk=1"
timeout -s KILL 200 python -m pytest tests/test_gpu_stages.py -x -q 2>&1 | tail -4
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes_bulk2.jsonl 2> $O/r02_gemm_shapes_bulk2.err
cat $O/r02_gemm_shapes_bulk2.jsonl | cut -c1-250; tail -3 $O/r02_gemm_shapes_bulk2.err
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 400 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_bench_bulk2.json 2> $O/r02_bench_bulk2.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_bulk2.json 2>&1 | grep -vE "^\s+\[.*(sb2st|q2)"; tail -3 $O/r02_bench_bulk2.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s KILL 200 $NCU -k regex:gemm_bulk_kernel -c 1 -o $O/r02_gemm_bulk2_8192 python scripts/ncu_target.py 8192 gemm > $O/r02_ncu_bulk2.log 2>&1
