#!/bin/bash
# Round 2, GPU call AO: L1 prefetch of the next tile's tool poses in the throughput-layout grid kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r03o_pytest.log 2>&1
tail -3 $O/r03o_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r03o_gathermove.json 2>&1
$B --workload cutrearrange > $O/r03o_cutrearrange.json 2>&1
$B --workload sweep:1000000:256 > $O/r03o_sweep1m.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r03o_timeline_gathermove_64.txt 2>&1
