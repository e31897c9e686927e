#!/bin/bash
# Round 2, GPU call C: plane-pass transposed scatter (19 KB / CTA) + parallel tile marking; occupancy sweep of the particle
# kernels (register caps 128 / 80 / 64 = 4 / 6 / 8 CTAs of 128 threads per SM) on the batched and the 1 M-particle configs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for minb in 4 6 8; do
  DSK_BIG_MINB=$minb $B --workload gathermove > $O/r02c_gathermove_minb$minb.json 2>&1
  DSK_BIG_MINB=$minb $B --workload sweep:1000000:256 > $O/r02c_sweep1m_minb$minb.json 2>&1
done
DSK_MINB_G2P2G=8 DSK_MINB_G2P_ADJ=8 DSK_MINB_P2G_ADJ=4 $B --workload gathermove > $O/r02c_gathermove_minb884.json 2>&1
DSK_MINB_G2P2G=8 DSK_MINB_G2P_ADJ=8 DSK_MINB_P2G_ADJ=4 $B --workload sweep:1000000:256 > $O/r02c_sweep1m_minb884.json 2>&1
DSK_MINB_G2P2G=8 DSK_MINB_G2P_ADJ=8 DSK_MINB_P2G_ADJ=6 $B --workload gathermove > $O/r02c_gathermove_minb886.json 2>&1
DSK_MINB_G2P2G=8 DSK_MINB_G2P_ADJ=8 DSK_MINB_P2G_ADJ=6 $B --workload sweep:1000000:256 > $O/r02c_sweep1m_minb886.json 2>&1
DSK_BIG_MINB=8 DSK_BIG_BLOCK=64 $B --workload gathermove > $O/r02c_gathermove_minb8_block64.json 2>&1
DSK_BIG_MINB=8 $B --workload cutrearrange > $O/r02c_cutrearrange_minb8.json 2>&1
DSK_BIG_MINB=8 $B --workload gathermove --envs 8 > $O/r02c_gathermove_8env_minb8.json 2>&1
DSK_FORCE_BIG=1 DSK_BIG_MINB=8 $B --workload gathermove --envs 8 > $O/r02c_gathermove_8env_big_minb8.json 2>&1
DSK_FORCE_BIG=1 DSK_BIG_MINB=8 $B --workload liftspread > $O/r02c_liftspread_big_minb8.json 2>&1
$B --workload liftspread > $O/r02c_liftspread.json 2>&1
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/r02c_pytest.log
ls -la $O | tail -5
