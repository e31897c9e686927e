#!/bin/bash
# Round 2, GPU call Q: grid kernels load the scalars of their prologue in one round trip; tile list two iterations ahead in the flat kernels

cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02q_pytest.log 2>&1
tail -3 $O/r02q_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r02q_gathermove.json 2>&1
$B --workload sweep:1000000:256 > $O/r02q_sweep1m.json 2>&1
$B --workload cutrearrange > $O/r02q_cutrearrange.json 2>&1
$B --workload liftspread > $O/r02q_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r02q_gathermove_8env.json 2>&1
$B --workload random_rollout > $O/r02q_random_rollout.json 2>&1
