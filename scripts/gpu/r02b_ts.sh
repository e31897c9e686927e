#!/bin/bash
# Round 2, GPU call B: transposed shared-memory scatter + 2-rsqrt packed SVD + factorised gathers.  Parity suite, bench of
# every config, A/B against the shuffle butterfly (DSK_TS=0), source-level ncu capture with the bank-conflict counters.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/r02b_pytest.log
for wl in gathermove liftspread cutrearrange sweep:1000000:256 sweep:100000:128 random_rollout; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_${wl//:/_}.json 2> $O/r02b_bench_${wl//:/_}.err
done
python bench.py --workload gathermove --envs 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_gathermove_8env.json 2>&1
for wl in gathermove liftspread sweep:1000000:256; do
  DSK_TS=0 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_${wl//:/_}_butterfly.json 2> $O/r02b_bench_${wl//:/_}_butterfly.err
done
DSK_BIG_BLOCK=64 python bench.py --workload gathermove --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_gathermove_block64.json 2>&1
DSK_BIG_MINB=3 python bench.py --workload gathermove --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_gathermove_minb3.json 2>&1
DSK_BIG_BLOCK=64 python bench.py --workload sweep:1000000:256 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_sweep_1000000_256_block64.json 2>&1
DSK_FORCE_BIG=1 python bench.py --workload liftspread --steps 5 --warmup 3 --no-cpu-baseline > $O/r02b_bench_liftspread_big.json 2>&1
M=lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__inst_executed_op_global_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROFILE_ITERS=1 timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:"k_g2p2g|k_g2p_adj|k_p2g_adj" -s 34 -c 6 \
  -o $O/r02b_ncu_sweep300k -f python scripts/profile_step.py sweep:300000:128 1 1 > $O/r02b_ncu_sweep300k.log 2>&1
ls -la $O | tail -30
