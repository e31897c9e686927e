#!/bin/bash
# Round 2, GPU call T: staged k_p2g_adj with the next wave's rows prefetched into L2 after the CTA's own rows arrived
# (DSK_PREFETCH=0: no prefetch); CutRearrange with the full substep tape (memory-based step-slot rule of bench.py).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for pf in 0 592 296; do
  DSK_PREFETCH=$pf $B --workload gathermove > $O/r02t_gathermove_pf$pf.json 2>&1
  DSK_PREFETCH=$pf $B --workload sweep:1000000:256 > $O/r02t_sweep1m_pf$pf.json 2>&1
done
DSK_STAGE=0 $B --workload cutrearrange > $O/r02t_cutrearrange_stage0.json 2>&1
$B --workload cutrearrange > $O/r02t_cutrearrange.json 2>&1
