#!/bin/bash
# Round 2 (second half), multi-GPU call: BASELINE configs[2] (GatherMove x64 sharded 64/N per GPU, strong) and configs[3] (CutRearrange,
# 32 start/goal pairs per GPU = 256 on 8 GPUs, weak) on N GPUs of one box.  usage: r03i_multigpu.sh N
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r03i_gpus_n$N.txt
run() {  # workload, port
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 \
    bench.py --gpus $N --steps 5 --warmup 3 --workload $1 > $O/r03i_bench_$1_n$N.json 2> $O/r03i_bench_$1_n$N.err
  tail -c 600 $O/r03i_bench_$1_n$N.err
}
run gathermove 29511
run cutrearrange 29512
run liftspread 29513
ls -la $O | tail -4
