#!/bin/bash
# Round 2, GPU call S: batch split over several engines on concurrent streams; and the parity tests call R missed.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python scripts/exp_split_batch.py gathermove 64 1 2 4 > $O/r02s_split_gathermove.txt 2>&1
timeout 600 python scripts/exp_split_batch.py cutrearrange 32 1 2 4 > $O/r02s_split_cutrearrange.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02s_pytest.log 2>&1
tail -3 $O/r02s_pytest.log
