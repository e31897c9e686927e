#!/bin/bash
# Round 2, GPU call AF: k_grid_adj_flat parks the contact adjoints, k_grid_adj_tools reduces the pose adjoints on the side branch.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -x -q > $O/r03f_pytest.log 2>&1
tail -3 $O/r03f_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r03f_gathermove.json 2>&1
$B --workload cutrearrange > $O/r03f_cutrearrange.json 2>&1
$B --workload liftspread > $O/r03f_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r03f_gathermove_8env.json 2>&1
$B --workload random_rollout > $O/r03f_random_rollout.json 2>&1
$B --workload liftspread --api gradmodel > $O/r03f_liftspread_gradmodel.json 2>&1
$B --workload gathermove --per-step-calls > $O/r03f_gathermove_per_step_calls.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r03f_timeline_gathermove_64.txt 2>&1
