#!/bin/bash
# Round 2, GPU call D: single-pass transposed scatter + parallel tile marking + guarded Jacobi rotation: full GPU suite
# (with the 50-step horizon test and the parity log), bench of every config.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40) > $O/r02d_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in gathermove liftspread cutrearrange sweep:1000000:256 random_rollout; do
  $B --workload $wl > $O/r02d_bench_${wl//:/_}.json 2> $O/r02d_bench_${wl//:/_}.err
done
DSK_BIG_MINB=6 $B --workload gathermove > $O/r02d_bench_gathermove_minb6.json 2>&1
DSK_MINB_G2P2G=6 DSK_MINB_G2P_ADJ=6 $B --workload sweep:1000000:256 > $O/r02d_bench_sweep_1000000_256_minb664.json 2>&1
$B --workload gathermove --envs 8 > $O/r02d_bench_gathermove_8env.json 2>&1
ls -la $O | tail -5
