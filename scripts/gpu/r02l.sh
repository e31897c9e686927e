#!/bin/bash
# Round 2, GPU call L: CTA size of the batched particle kernels (the 14 KB-per-warp scatter tile allows 3 CTAs of 128 threads =
# 12 warps per SM, but 5 CTAs of 96 threads = 15 warps).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for b in 128 96 64; do
  DSK_BIG_BLOCK=$b $B --workload gathermove > $O/r02l_gathermove_block$b.json 2>&1
  DSK_BIG_BLOCK=$b $B --workload sweep:1000000:256 > $O/r02l_sweep1m_block$b.json 2>&1
  DSK_BIG_BLOCK=$b $B --workload cutrearrange > $O/r02l_cutrearrange_block$b.json 2>&1
done
