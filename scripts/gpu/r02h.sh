#!/bin/bash
# Round 2, GPU call H: per-rank spread attribution, ncu sections + counters + DRAM traffic of the final kernels, sweep grid.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
DSK_LIB=timeline python scripts/rank_spread.py liftspread 8 > $O/r02h_rank_spread_liftspread.md 2> $O/r02h_rank_spread.err
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROFILE_ITERS=1 timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_g2p2g|k_g2p_adj|k_p2g_adj" -s 34 -c 6 \
  -o $O/r02h_ncu_sweep300k -f python scripts/profile_step.py sweep:300000:128 1 1 > $O/r02h_ncu_sweep300k.log 2>&1
PROFILE_ITERS=1 timeout 600 ncu --set full --metrics $M --clock-control none -k regex:"k_g2p2g|k_g2p_adj|k_p2g_adj|k_grid_flat|k_grid_adj_flat" -s 40 -c 10 \
  -o $O/r02h_ncu_gathermove64 -f python scripts/profile_step.py gathermove 1 64 > $O/r02h_ncu_gathermove64.log 2>&1
PROFILE_ITERS=1 timeout 600 ncu --set full --metrics $M --clock-control none -k regex:"k_g2p2g|k_g2p_adj|k_p2g_adj" -s 100 -c 6 \
  -o $O/r02h_ncu_sweep1m -f python scripts/profile_step.py sweep:1000000:256 1 1 > $O/r02h_ncu_sweep1m.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r02h_launches_gathermove.csv python bench.py --workload gathermove --steps 1 --warmup 1 --no-cpu-baseline > $O/r02h_launches_bench.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
for g in 64 128 256; do for n in 10000 30000 100000 300000 1000000; do
  $B --workload sweep:$n:$g > $O/r02h_sweep_${n}_${g}.json 2> $O/r02h_sweep_${n}_${g}.err
done; done
ls $O | grep r02h | wc -l
