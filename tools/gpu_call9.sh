#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
python tools/sweep_variants.py 2d_ring 2>&1 | tee gpurun_out/sweepring_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-10s %-14s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))"
timeout 300 ncu --set full --clock-control none -k regex:k_ring2d -s 1 -c 1 -f -o gpurun_out/ring2d_$tag python tools/profile_target.py 2d 300 auto > gpurun_out/ncu_ring2d_$tag.log 2>&1
python tools/ncu_brief.py gpurun_out/ring2d_$tag.ncu-rep > gpurun_out/ncu_ring2d_$tag.txt 2>&1; rm -f gpurun_out/ring2d_$tag.ncu-rep
grep -v "^$" gpurun_out/ncu_ring2d_$tag.txt | head -40
