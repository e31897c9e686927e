#!/bin/bash
# Round 2, GPU call O: as call N with the claims deduplicated per run of lanes sharing a base tile (sorted order)

cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02o_pytest.log 2>&1
tail -3 $O/r02o_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r02o_gathermove.json 2>&1
$B --workload sweep:1000000:256 > $O/r02o_sweep1m.json 2>&1
$B --workload cutrearrange > $O/r02o_cutrearrange.json 2>&1
$B --workload liftspread > $O/r02o_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r02o_gathermove_8env.json 2>&1
$B --workload random_rollout > $O/r02o_random_rollout.json 2>&1
