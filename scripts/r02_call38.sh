#!/bin/bash
# Round 2, GPU call 38: the full GPU test suite and smoke() on the final HEAD.
set -u
timeout -s KILL 110 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -s KILL 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
