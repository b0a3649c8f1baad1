#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 400 python tools/sweep_variants.py pdl_tile_grid,pdl_tile_512,2d_grid 2>&1 | tee gpurun_out/sweep_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-14s %-18s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')"
