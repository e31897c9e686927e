#!/bin/bash
# Round 2, GPU call AH: eager path with the split adjoint grid kernels (what bench.py profiles): full GPU suite again, default
# bench line, LiftSpread / 8-env lines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r03h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r03h_smoke.log 2>&1
python bench.py > $O/r03h_bench_default.json 2> $O/r03h_bench_default.err
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --workload liftspread > $O/r03h_bench_liftspread.json 2>&1
$B --workload gathermove --envs 8 > $O/r03h_bench_gathermove_8env.json 2>&1
$B --workload cutrearrange > $O/r03h_bench_cutrearrange.json 2>&1
