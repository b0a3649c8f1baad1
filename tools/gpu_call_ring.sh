#!/bin/bash
# usage: bash tools/gpu_call_ring.sh <tag> <n_gpus> <n0> <n1> <passes>
tag=${1:-x}; n=${2:-2}; n0=${3:-8192}; n1=${4:-4096}; np=${5:-200}
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tools/slab_multi_gpu.py $n0 $n1 $np ) > gpurun_out/ring_n${n}_$tag.log 2> gpurun_out/ring_n${n}_$tag.err
grep '^{' gpurun_out/ring_n${n}_$tag.log | cut -c1-1500; tail -5 gpurun_out/ring_n${n}_$tag.err
