#!/bin/bash
# Round 2, GPU call AK: run reduction of the transposed scatter as one unrolled pass over the row (one shuffle per run).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r03k_pytest.log 2>&1
tail -3 $O/r03k_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r03k_gathermove.json 2>&1
$B --workload sweep:1000000:256 > $O/r03k_sweep1m.json 2>&1
$B --workload cutrearrange > $O/r03k_cutrearrange.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r03k_timeline_gathermove_64.txt 2>&1
