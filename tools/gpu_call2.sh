#!/bin/bash
# pytest (with per-test timeout) + the drop-in API throughput
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=8 --timeout 150 ) > gpurun_out/pytest_$tag.log 2>&1
tail -15 gpurun_out/pytest_$tag.log
( timeout 300 python tools/api_run_bench.py 4096 4096 10000 100 ) > gpurun_out/api_run_$tag.log 2>&1
tail -5 gpurun_out/api_run_$tag.log
( USE_NLIST=0 timeout 300 python tools/api_run_bench.py 4096 4096 10000 100 ) > gpurun_out/api_run_nonlist_$tag.log 2>&1
tail -3 gpurun_out/api_run_nonlist_$tag.log
