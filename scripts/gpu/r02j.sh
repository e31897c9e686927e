#!/bin/bash
# Round 2, GPU call J: CTAs per SM of the latency-layout grid kernels (A/B), ncu DRAM traffic of k_g2p2g on GatherMove x64.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for c in 2 3 4 6; do
  DSK_GRID_CTAS_PER_SM=$c $B --workload liftspread > $O/r02j_liftspread_gridctas$c.json 2>&1
  DSK_GRID_CTAS_PER_SM=$c $B --workload gathermove --envs 8 > $O/r02j_gathermove_8env_gridctas$c.json 2>&1
done
PROFILE_ITERS=1 timeout 600 ncu --set full --clock-control none -k regex:"k_g2p2g|k_grid_flat" -s 6 -c 6 \
  -o $O/r02j_ncu_gathermove64_fwd -f python scripts/profile_step.py gathermove 1 64 > $O/r02j_ncu_gathermove64_fwd.log 2>&1
