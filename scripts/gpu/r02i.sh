#!/bin/bash
# Round 2, GPU call I: scatter variant of the plane-split kernels on fragmented doughs (the slow rank of the weak-scaling run,
# the 8-env GatherMove shard), SVD tape on/off, kernel family crossover.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for ts in 0 1; do
  DSK_TS_PL=$ts $B --workload liftspread --env-offset 4 > $O/r02i_liftspread_env4_tspl$ts.json 2>&1
  DSK_TS_PL=$ts $B --workload liftspread --env-offset 5 > $O/r02i_liftspread_env5_tspl$ts.json 2>&1
  DSK_TS_PL=$ts $B --workload gathermove --envs 8 > $O/r02i_gathermove_8env_tspl$ts.json 2>&1
done
DSK_FORCE_BIG=1 $B --workload gathermove --envs 8 > $O/r02i_gathermove_8env_big.json 2>&1
DSK_FORCE_BIG=0 $B --workload gathermove --envs 16 > $O/r02i_gathermove_16env_small.json 2>&1
$B --workload gathermove --envs 16 > $O/r02i_gathermove_16env.json 2>&1
DSK_NO_SVD_TAPE=1 $B --workload gathermove > $O/r02i_gathermove_nosvdtape.json 2>&1
DSK_NO_SVD_TAPE=1 $B --workload sweep:1000000:256 > $O/r02i_sweep1m_nosvdtape.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 8 > $O/r02i_timeline_gathermove_8.txt 2>&1
