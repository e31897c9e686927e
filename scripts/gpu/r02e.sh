#!/bin/bash
# Round 2, GPU call E: tool adjoints deferred to a side stream (A/B), the reference-API path (GradModel autograd loop),
# device-side planner, in-graph timelines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py tests/test_gpu_aux.py -m gpu -q 2>&1 | tail -15) > $O/r02e_pytest.log
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "Rope or without_fast_math or svd" 2>&1 | tail -8) >> $O/r02e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in gathermove liftspread cutrearrange; do
  $B --workload $wl > $O/r02e_bench_${wl}.json 2> $O/r02e_bench_${wl}.err
  DSK_TOOL_DEFER=0 $B --workload $wl > $O/r02e_bench_${wl}_nodefer.json 2>&1
done
$B --workload gathermove --envs 8 > $O/r02e_bench_gathermove_8env.json 2>&1
$B --workload liftspread --api gradmodel > $O/r02e_bench_liftspread_api_gradmodel.json 2> $O/r02e_bench_liftspread_api_gradmodel.err
$B --workload gathermove --envs 1 --api gradmodel > $O/r02e_bench_gathermove_api_gradmodel.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r02e_timeline_gathermove_64.txt 2>&1
DSK_LIB=timeline python scripts/timeline_step.py liftspread 1 > $O/r02e_timeline_liftspread_1.txt 2>&1
ls -la $O | tail -5
