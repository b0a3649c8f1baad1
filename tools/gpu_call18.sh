#!/bin/bash
# programmatic dependent launch / chained half-sweeps: parity tests, then A/B
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "chained or bulk or full_size or 3d or golden or rollback" 2>&1 | tail -5
timeout 400 python tools/sweep_variants.py pdl_3d,pdl_2d_sweep8,pdl_2d_one,pdl_3d_256,pdl_2d_16k,chain_ns_2d,chain_ns_3d 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('  %-14s %-18s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))
    except Exception: print(l, end='')" | tee gpurun_out/pdl_$tag.txt
