#!/bin/bash
# Round 2, GPU call AR: ncu --set full of the particle / grid kernels at CutRearrange x32 (DRAM bytes per launch for
# profiles/traffic.json, counters).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROFILE_ITERS=1 timeout 900 ncu --set full --metrics $M --clock-control none -k regex:"k_grid_flat|k_g2p2g" -s 60 -c 4 \
  -o $O/r03r_ncu_cutrearrange32_fwd -f python scripts/profile_step.py cutrearrange 2 32 > $O/r03r_ncu_fwd.log 2>&1
PROFILE_ITERS=1 timeout 900 ncu --set full --metrics $M --clock-control none -k regex:"k_grid_adj_flat|k_g2p_adj|k_p2g_adj" -s 30 -c 6 \
  -o $O/r03r_ncu_cutrearrange32_bwd -f python scripts/profile_step.py cutrearrange 2 32 > $O/r03r_ncu_bwd.log 2>&1
tail -2 $O/r03r_ncu_bwd.log
