#!/bin/bash
# Round 2, GPU call 19: ncu captures of the CURRENT kernels on the shapes that dominate the solve; CPU oracle at n = 16384.
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# dense-to-band at n = 16384: the TMA-fed kernel on the N = 64 symmetric panel product and the K = 128 update (panel ~6)
timeout -s KILL 200 $NCU -k regex:gemm_bulk_kernel --launch-skip 30 -c 6 -o $O/r02_sy2sb_bulk_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_sy2sb_bulk.log 2>&1
# Q2 with the producer warp + ring
timeout -s KILL 300 $NCU -k regex:q2_apply_kernel -c 1 -o $O/r02_q2_ring_16384 python scripts/ncu_target.py 16384 stages > $O/r02_ncu_q2_ring.log 2>&1
# the batched D&C merge products (top levels) at n = 8192 through the whole solve
timeout -s KILL 300 $NCU -k regex:gemm_bulk_kernel.*Lb1 --launch-skip 6 -c 2 -o $O/r02_stedc_merge_8192 python bench.py --n 8192 --steps 1 --warmup 0 --no-e2e --no-cpu --no-check > $O/r02_ncu_merge.log 2>&1
ls -la $O/*.ncu-rep | tail -4
# CPU oracle at n = 16384 (once per round, outside the driver's step count)
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 --cpu-n 16384 --no-cpu-table > $O/r02_cpu_n16384.json 2> $O/r02_cpu_n16384.err
cut -c1-900 $O/r02_cpu_n16384.json; tail -2 $O/r02_cpu_n16384.err
