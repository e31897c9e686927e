#!/bin/bash
# Round 2, GPU call M: tile-tag loads moved ahead of the scatter (compare after it) and L2 bulk prefetch of the particle rows.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02m_pytest.log 2>&1
tail -3 $O/r02m_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for pf in -1 0 592; do
  DSK_PREFETCH=$pf $B --workload gathermove > $O/r02m_gathermove_pf$pf.json 2>&1
  DSK_PREFETCH=$pf $B --workload sweep:1000000:256 > $O/r02m_sweep1m_pf$pf.json 2>&1
  DSK_PREFETCH=$pf $B --workload cutrearrange > $O/r02m_cutrearrange_pf$pf.json 2>&1
  DSK_PREFETCH=$pf $B --workload liftspread > $O/r02m_liftspread_pf$pf.json 2>&1
done
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r02m_timeline_gathermove_64.txt 2>&1
