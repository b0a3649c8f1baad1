#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
python tools/diag_case.py 16 2>&1 | grep -v "\[\] samples differ in \[\]" | tail -20
python tools/auto_vs_generic.py 2>&1 | tee gpurun_out/auto_vs_generic_$tag.json | grep -v "true" | tail -18
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "tile or grid or multichain or variant or golden or full_size or rollback or drop_in or run_management" 2>&1 | tail -4
timeout 400 python tools/sweep_variants.py pdl_tile_512,pdl_tile_768,2d_tile 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-14s %-18s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')"
