#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "tile or auto_selection or two_streams or rollback or host_formats or drop_in" ) 2>&1 | tail -6
timeout 300 python tools/sweep_variants.py pdl_tile_512,tile_nt_256 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-14s %-18s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')" | grep "auto \|nt=" 
