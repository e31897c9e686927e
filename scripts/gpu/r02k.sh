#!/bin/bash
# Round 2, GPU call K: full GPU suite with Chopsticks / MLP / all round-2 changes, smoke, default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02k_smoke.log 2>&1
python bench.py > $O/r02k_bench_default.json 2> $O/r02k_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02k_bench_reference_arm.json 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in liftspread cutrearrange sweep:1000000:256; do
  $B --workload $wl > $O/r02k_bench_${wl//:/_}.json 2> $O/r02k_bench_${wl//:/_}.err
done
python bench.py --workload random_rollout --steps 5 --warmup 3 > $O/r02k_bench_random_rollout.json 2>&1
