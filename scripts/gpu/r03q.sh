#!/bin/bash
# Round 2, GPU call AQ: run-to-run spread of smoke()'s distances from the oracle (float atomics make the scatter order, hence the
# rounding, differ between runs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in $(seq 1 14); do
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "^smoke" | cut -c1-110
done > gpurun_out/r03q_smoke_spread.txt
DSK_NO_GRAPHS=1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "^smoke" | cut -c1-110 >> gpurun_out/r03q_smoke_spread.txt
