#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
( timeout 900 python bench.py ) > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
grep -o '"value": [0-9.e+]*\|"frac": [0-9.]*' gpurun_out/bench_$tag.log | head -8 | tr '\n' ' '; tail -2 gpurun_out/bench_$tag.err
