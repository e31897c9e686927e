#!/bin/bash
# Round 2, GPU call V: Jacobi rotations of converged column pairs skipped (svd3.cuh); parity suite + benches.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_horizon.py -m gpu -x -q > $O/r02v_pytest.log 2>&1
tail -3 $O/r02v_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B --workload gathermove > $O/r02v_gathermove.json 2>&1
$B --workload sweep:1000000:256 > $O/r02v_sweep1m.json 2>&1
$B --workload cutrearrange > $O/r02v_cutrearrange.json 2>&1
$B --workload liftspread > $O/r02v_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r02v_gathermove_8env.json 2>&1
$B --workload random_rollout > $O/r02v_random_rollout.json 2>&1
