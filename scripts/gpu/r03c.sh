#!/bin/bash
# Round 2, GPU call AC: ncu --set full (source counters) of the grid kernels, the tool kinematics and the final particle
# kernels at GatherMove x64.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROFILE_ITERS=1 timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_grid_adj_flat|k_grid_flat|k_kinematics|k_g2p2g|k_g2p_adj|k_p2g_adj|k_permute_rows" -s 60 -c 14 \
  -o $O/r03c_ncu_gathermove64 -f python scripts/profile_step.py gathermove 1 64 > $O/r03c_ncu_gathermove64.log 2>&1
tail -3 $O/r03c_ncu_gathermove64.log
