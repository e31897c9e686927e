#!/bin/bash
# Round 2, GPU call G: software-pipelined flat grid kernels, defer heuristic; quick parity check + bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batched_layout or fine_grained or golden" 2>&1 | tail -5) > $O/r02g_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in gathermove liftspread cutrearrange; do
  $B --workload $wl > $O/r02g_bench_${wl}.json 2> $O/r02g_bench_${wl}.err
done
$B --workload gathermove --envs 8 > $O/r02g_bench_gathermove_8env.json 2>&1
DSK_LIB=timeline python scripts/timeline_step.py gathermove 64 > $O/r02g_timeline_gathermove_64.txt 2>&1
