#!/bin/bash
# where does SemiGrandCanonicalCalculator.run spend its time? kernel list (ncu, serialised) of one API run + plain timings
tag=${1:-x}
mkdir -p gpurun_out
python tools/api_run_bench.py 4096 4096 10000 100 2>&1 | grep '^{' | tee gpurun_out/api_run_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_api_$tag.csv python tools/api_run_bench.py 4096 4096 3000 100 > gpurun_out/api_under_ncu_$tag.log 2>&1
python - $tag <<'P'
import csv, collections, sys
rows = [r for r in csv.reader(open('gpurun_out/launches_api_%s.csv' % sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:60]
    tot[name] += float(r[-1]); cnt[name] += 1
for k, v in tot.most_common(25):
    print('%-62s n=%5d total %.3f ms avg %.1f us' % (k, cnt[k], v * 1e-6, v * 1e-3 / cnt[k]))
P
