#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_$tag.log 2>&1
tail -6 gpurun_out/pytest_$tag.log
python tools/sweep_variants.py 2d_slab8,2d_sweep8 auto 2>&1 | tee gpurun_out/sweep2d_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-10s %-14s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))"
