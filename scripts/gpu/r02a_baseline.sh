#!/bin/bash
# Round 2, GPU call A: state of the round-1 kernels on every BASELINE config (fresh post-svd3 numbers), the phase-boundary
# micro-benchmark, and source-level ncu captures of the three particle kernels with the atomic / bank-conflict counters.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02a_gpu.txt
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > $O/r02a_pytest.log
./scripts/micro/sync_chain.bin > $O/r02a_sync_chain.txt 2>&1
python bench.py --workload gathermove --steps 5 --warmup 3 > $O/r02a_bench_gathermove.json 2> $O/r02a_bench_gathermove.err
for wl in liftspread random_rollout cutrearrange sweep:1000000:256 sweep:100000:128; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $O/r02a_bench_${wl//:/_}.json 2> $O/r02a_bench_${wl//:/_}.err
done
python bench.py --workload gathermove --envs 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02a_bench_gathermove_8env.json 2>&1
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum
PROFILE_ITERS=1 timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_g2p2g|k_g2p_adj|k_p2g_adj" -s 30 -c 12 \
  -o $O/r02a_ncu_sweep300k -f python scripts/profile_step.py sweep:300000:128 1 1 > $O/r02a_ncu_sweep300k.log 2>&1
ls -la $O
