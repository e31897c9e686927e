#!/bin/bash
# Round 2 (second half): default line (GatherMove x64 sharded) on N GPUs.  usage: r03j_multigpu_gathermove.sh N
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r03i_bench_gathermove_n$N.json 2> gpurun_out/r03i_bench_gathermove_n$N.err
tail -c 300 gpurun_out/r03i_bench_gathermove_n$N.err
