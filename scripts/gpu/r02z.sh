#!/bin/bash
# Round 2, GPU call Z: k_permute_rows with compile-time rows per CTA (<= 48 KB: 384 CTAs for 64 envs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -x -q > $O/r02z_pytest.log 2>&1
tail -3 $O/r02z_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r02z_gathermove.json 2>&1
$B --workload sweep:1000000:256 > $O/r02z_sweep1m.json 2>&1
$B --workload cutrearrange > $O/r02z_cutrearrange.json 2>&1
$B --workload liftspread > $O/r02z_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r02z_gathermove_8env.json 2>&1
$B --workload random_rollout > $O/r02z_random_rollout.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 --full > $O/r02z_timeline_gathermove_64.txt 2>&1
