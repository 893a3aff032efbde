#!/bin/bash
# Round 2, GPU call 27: L2-aware tile order in the TMA-fed GEMM kernel -- tests, shape table, ncu of 8192^3, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_twostage.py tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -3
timeout -s KILL 200 python scripts/gemm_shapes_probe.py > $O/r02_gemm_shapes_raster.jsonl 2> $O/shapes.err; tail -3 $O/shapes.err
cut -c1-200 $O/r02_gemm_shapes_raster.jsonl
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -s KILL 200 $NCU -k regex:gemm_bulk_kernel -c 1 -o $O/r02_gemm_raster_8192 python scripts/ncu_target.py 8192 gemm > $O/r02_ncu_raster.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02_ncu_raster.log
timeout -s KILL 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > $O/r02_bench_raster.json 2> $O/bench_raster.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_raster.json 2>&1 | grep -vE "^\s+\[" | head -16
