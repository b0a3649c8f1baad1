#!/bin/bash
# final verification at HEAD: full GPU suite, smoke(), N = 1 bench
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/pytest_$tag.log 2>&1; tail -6 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 600 python bench.py ) > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 1500 gpurun_out/bench_$tag.log; tail -5 gpurun_out/bench_$tag.err
