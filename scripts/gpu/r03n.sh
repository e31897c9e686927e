#!/bin/bash
# Round 2, GPU call AN: fewer CTAs per SM for the throughput-layout grid kernels (call AM: more is slower).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for a in 2 3; do
  DSK_FLAT_ADJ_CTAS=$a $B --workload gathermove > $O/r03n_gathermove_adj$a.json 2>&1
  DSK_FLAT_FWD_CTAS=$a $B --workload gathermove > $O/r03n_gathermove_fwd$a.json 2>&1
done
