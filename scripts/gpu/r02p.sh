#!/bin/bash
# Round 2, GPU call P: full in-graph timeline of GatherMove x64 (with and without the two-branch backward), to find why
# k_grid_adj_flat takes 18 us inside the graph and 10 us alone under ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 --full > $O/r02p_timeline_gathermove_64.txt 2>&1
DSK_NO_PIPELINE=1 DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 --full > $O/r02p_timeline_gathermove_64_nopipeline.txt 2>&1
DSK_LIB=timeline python scripts/timeline_step.py cutrearrange 32 --full > $O/r02p_timeline_cutrearrange_32.txt 2>&1
DSK_LIB=timeline python scripts/timeline_step.py liftspread 1 --full > $O/r02p_timeline_liftspread_1.txt 2>&1
