#!/bin/bash
# Round 2, GPU calls 26 and 33: final verification on HEAD -- full GPU test suite, smoke, the default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py > $O/r02_bench_n32768_p1_final3.json 2> $O/bench_final3.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_n32768_p1_final3.json 2>&1 | grep -vE "^\s+\[" | head -24; tail -2 $O/bench_final3.err
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/r02_bench_reference_final3.json 2>> $O/bench_final3.err
echo "ref rc=$?"; cut -c1-400 $O/r02_bench_reference_final3.json
