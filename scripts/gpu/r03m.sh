#!/bin/bash
# Round 2, GPU call AM: grid size of the throughput-layout grid kernels (more, shorter CTAs: the block scheduler balances the
# few contact-heavy tiles).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for a in 4 8 16 32; do
  DSK_FLAT_ADJ_CTAS=$a $B --workload gathermove > $O/r03m_gathermove_adj$a.json 2>&1
done
for f in 8 16; do
  DSK_FLAT_FWD_CTAS=$f DSK_FLAT_ADJ_CTAS=16 $B --workload gathermove > $O/r03m_gathermove_adj16_fwd$f.json 2>&1
done
DSK_FLAT_ADJ_CTAS=16 $B --workload cutrearrange > $O/r03m_cutrearrange_adj16.json 2>&1
DSK_FLAT_ADJ_CTAS=16 $B --workload sweep:1000000:256 > $O/r03m_sweep1m_adj16.json 2>&1
