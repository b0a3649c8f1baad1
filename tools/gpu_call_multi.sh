#!/bin/bash
# usage (on the GPU box): bash tools/gpu_call_multi.sh <tag> <n_gpus> [steps] [warmup]
tag=${1:-x}; n=${2:-2}; steps=${3:-5}; warm=${4:-3}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup $warm ) > gpurun_out/bench_n${n}_$tag.log 2> gpurun_out/bench_n${n}_$tag.err
tail -c 6000 gpurun_out/bench_n${n}_$tag.log; tail -15 gpurun_out/bench_n${n}_$tag.err
