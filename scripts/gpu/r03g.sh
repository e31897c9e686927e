#!/bin/bash
# Round 2, GPU call AG: final validation of the second half of round 2 (full GPU suite, smoke, default bench line, reference arm, the other workloads, launch list).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r03g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r03g_smoke.log 2>&1
python bench.py > $O/r03g_bench_default.json 2> $O/r03g_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r03g_bench_reference_arm.json 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in liftspread cutrearrange sweep:1000000:256; do
  $B --workload $wl > $O/r03g_bench_${wl//:/_}.json 2> $O/r03g_bench_${wl//:/_}.err
done
python bench.py --workload random_rollout --steps 5 --warmup 3 > $O/r03g_bench_random_rollout.json 2>&1
$B --workload gathermove --envs 8 > $O/r03g_bench_gathermove_8env.json 2>&1
$B --workload liftspread --api gradmodel > $O/r03g_bench_liftspread_gradmodel.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r03g_launches_gathermove.csv python bench.py --workload gathermove --steps 1 --warmup 1 --no-cpu-baseline > $O/r03g_launches_bench.log 2>&1
