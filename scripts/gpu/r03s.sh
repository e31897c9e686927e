#!/bin/bash
# Round 2, GPU call AS: configs[4] sweep grid with the final kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
for g in 64 128 256; do for n in 10000 30000 100000 300000 1000000; do
  $B --workload sweep:$n:$g > $O/r03s_sweep_${n}_${g}.json 2> $O/r03s_sweep_${n}_${g}.err
done; done
ls $O | grep r03s | wc -l
