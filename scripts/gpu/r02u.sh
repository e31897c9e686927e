#!/bin/bash
# Round 2, GPU call U: full counting sort every n-th env step only (DSK_RESORT_INTERVAL), with the transposed scatter.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for n in 1 2 4 8; do
  DSK_RESORT_INTERVAL=$n $B --workload gathermove > $O/r02u_gathermove_resort$n.json 2>&1
  DSK_RESORT_INTERVAL=$n $B --workload cutrearrange > $O/r02u_cutrearrange_resort$n.json 2>&1
done
DSK_RESORT_INTERVAL=4 $B --workload liftspread > $O/r02u_liftspread_resort4.json 2>&1
DSK_RESORT_INTERVAL=4 $B --workload gathermove --envs 8 > $O/r02u_gathermove_8env_resort4.json 2>&1
