#!/bin/bash
# Round 2, GPU call AA: k_permute_rows for single scenes / small batches too (DSK_PERM_SMEM_SMALL=1)?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for v in 0 1; do
  DSK_PERM_SMEM_SMALL=$v $B --workload liftspread > $O/r03a_liftspread_small$v.json 2>&1
  DSK_PERM_SMEM_SMALL=$v $B --workload gathermove --envs 8 > $O/r03a_gathermove_8env_small$v.json 2>&1
  DSK_PERM_SMEM_SMALL=$v $B --workload random_rollout > $O/r03a_random_rollout_small$v.json 2>&1
done
