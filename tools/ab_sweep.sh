#!/bin/bash
# A/B run of the library variants built by tools/ab_build.py (on the GPU box)
# usage: tools/ab_sweep.sh "2d_sweep8,2d_big,3d" name1 name2 ...
cases=$1; shift
for v in "$@"; do
  echo "== $v"
  CMG_LIB_PATH=casmcode_monte_b200/_variants/lib_$v.so python tools/sweep_variants.py "$cases" auto | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-10s %-8s sp=%d  %.4g att/s  %.2f us' % (d['case'], d['variant'], d['sample_period'], d['attempts_per_s'], d['us_per_halfsweep']))"
done
