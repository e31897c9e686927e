#!/bin/bash
# Round 2, GPU call R: k_p2g_adj with its particle rows staged in shared memory by cp.async (DSK_STAGE=0: plain loads).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_planner.py -m gpu -x -q > $O/r02r_pytest.log 2>&1
tail -3 $O/r02r_pytest.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for st in 0 1; do
  DSK_STAGE=$st $B --workload gathermove > $O/r02r_gathermove_stage$st.json 2>&1
  DSK_STAGE=$st $B --workload sweep:1000000:256 > $O/r02r_sweep1m_stage$st.json 2>&1
  DSK_STAGE=$st $B --workload cutrearrange > $O/r02r_cutrearrange_stage$st.json 2>&1
done
