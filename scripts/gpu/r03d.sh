#!/bin/bash
# Round 2, GPU call AD: ncu --set full of the grid kernels / kinematics / particle kernels at GatherMove x64 in the fifth env
# step (tools in contact), forward and backward windows.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROFILE_ITERS=1 timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_grid_flat|k_kinematics|k_g2p2g" -s 156 -c 8 \
  -o $O/r03d_ncu_gathermove64_fwd -f python scripts/profile_step.py gathermove 6 64 > $O/r03d_ncu_fwd.log 2>&1
PROFILE_ITERS=1 timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_grid_adj_flat|k_g2p_adj|k_p2g_adj" -s 66 -c 6 \
  -o $O/r03d_ncu_gathermove64_bwd -f python scripts/profile_step.py gathermove 6 64 > $O/r03d_ncu_bwd.log 2>&1
tail -2 $O/r03d_ncu_bwd.log
