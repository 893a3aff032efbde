#!/bin/bash
# Round 2, GPU call 16: host entry points with overlapped transfers; Q2 ring; full suite + headline bench with e2e.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py --steps 2 --warmup 1 > $O/r02_bench_final1.json 2> $O/r02_bench_final1.err
echo "bench rc=$?"; python scripts/show_bench.py $O/r02_bench_final1.json 2>&1 | grep -vE "^\s+\["; tail -3 $O/r02_bench_final1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_final1.json') if l.startswith('{')][-1]
print(json.dumps(d['cpu_baseline'])[:900])
print(d['roofline'])
PY
