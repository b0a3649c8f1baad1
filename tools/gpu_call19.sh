#!/bin/bash
# full GPU suite + N=1 bench + the A/B table of the chained launches
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/pytest_$tag.log 2>&1; tail -6 gpurun_out/pytest_$tag.log
( time timeout 600 python bench.py ) > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
tail -c 6000 gpurun_out/bench_$tag.log; tail -5 gpurun_out/bench_$tag.err
timeout 400 python tools/sweep_variants.py pdl_3d,pdl_2d_sweep8,pdl_2d_one,pdl_3d_256,pdl_2d_16k,2d_one 2>&1 | tee gpurun_out/chain_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-14s %-18s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')"
