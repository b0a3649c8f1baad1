#!/bin/bash
# compute-sanitizer memcheck over the kernels touched this session (small shapes)
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q --timeout 900 -k "bulk2d_matches or bulk3d_matches or tile2d_matches or ring2d_matches or multichain_grid or kstate" > gpurun_out/memcheck_$tag.log 2>&1
echo rc=$?; tail -8 gpurun_out/memcheck_$tag.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kstate.py -x -q --timeout 500 > gpurun_out/memcheck_kstate_$tag.log 2>&1
echo rc=$?; tail -4 gpurun_out/memcheck_kstate_$tag.log
