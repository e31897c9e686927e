#!/bin/bash
# Round 2, GPU call AW: kernel-family crossover, second part (12 / 16 envs in the single-scene family).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
DSK_FORCE_BIG=0 $B --workload gathermove --envs 12 > $O/r03w_gathermove_12env_small.json 2>&1
DSK_FORCE_BIG=0 $B --workload gathermove --envs 16 > $O/r03w_gathermove_16env_small.json 2>&1
$B --workload gathermove --envs 16 > $O/r03w_gathermove_16env.json 2>&1
