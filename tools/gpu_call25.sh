#!/bin/bash
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python tools/sweep_variants.py mid_ring_256sq,2d_ringsync_small 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-18s %-10s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')"
