#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
python tools/sweep_variants.py 3d_js,3d 2>&1 | tee gpurun_out/sweep3djs_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-10s %-14s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))"
