#!/bin/bash
# Round 2, GPU call AT: k_kinematics with the collision queries of an env shared by 8 CTAs when it is on the critical path
# (per-step entry points): full GPU suite, the per-step call paths, default line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r03t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r03t_smoke.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
for p in 1 8; do
  DSK_KIN_PARTS=$p $B --workload gathermove --per-step-calls > $O/r03t_gathermove_per_step_calls_parts$p.json 2>&1
  DSK_KIN_PARTS=$p $B --workload liftspread --api gradmodel > $O/r03t_liftspread_gradmodel_parts$p.json 2>&1
  DSK_KIN_PARTS=$p $B --workload liftspread --per-step-calls > $O/r03t_liftspread_per_step_calls_parts$p.json 2>&1
done
python bench.py > $O/r03t_bench_default.json 2> $O/r03t_bench_default.err
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r03t_timeline_gathermove_64.txt 2>&1
