#!/bin/bash
# usage (on the GPU box, from the repo root): bash tools/gpu_call.sh <tag> [pytest|bench|all]
tag=${1:-x}; what=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$tag.txt 2>&1
if [ "$what" = "pytest" ] || [ "$what" = "all" ]; then
  ( time timeout 600 python -m pytest tests -m gpu -x -q --durations=15 --timeout 150 ) > gpurun_out/pytest_$tag.log 2>&1
  tail -30 gpurun_out/pytest_$tag.log
fi
if [ "$what" = "bench" ] || [ "$what" = "all" ]; then
  ( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
  tail -c 3000 gpurun_out/bench_$tag.log; tail -5 gpurun_out/bench_$tag.err
fi
