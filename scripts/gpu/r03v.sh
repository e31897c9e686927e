#!/bin/bash
# Round 2, GPU call AV: kernel-family crossover with the final kernels (8 / 12 envs per GPU in the batched layout?).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
DSK_FORCE_BIG=1 $B --workload gathermove --envs 8 > $O/r03v_gathermove_8env_big.json 2>&1
$B --workload gathermove --envs 8 > $O/r03v_gathermove_8env.json 2>&1
$B --workload gathermove --envs 12 > $O/r03v_gathermove_12env.json 2>&1
DSK_FORCE_BIG=1 $B --workload gathermove --envs 12 > $O/r03v_gathermove_12env_big.json 2>&1
