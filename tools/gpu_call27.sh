#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
for c in config4full config4; do
  timeout 600 python tools/run_baseline_configs.py $c > gpurun_out/${c}_$tag.json 2> gpurun_out/${c}_$tag.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/${c}_$tag.json')); print('$c', d['kernel'], d['n_chains'], '%.4g'%d['attempts_per_s'], d['seconds_measured'], d['seconds_device_statistics_all_chains'])"; tail -2 gpurun_out/${c}_$tag.err
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "grid or multichain or variant or sharding or drop_in" 2>&1 | tail -4
