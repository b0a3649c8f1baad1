#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
echo "== waiting loop"; CMG_OVERLAP_CHECKS=0 OVERLAP=0 python tools/api_run_bench.py 4096 4096 10000 100 2>&1 | grep '^{' | tee gpurun_out/api_run_$tag.log | cut -c1-200
echo "== overlapped"; OVERLAP=1 python tools/api_run_bench.py 4096 4096 10000 100 2>&1 | grep '^{' | tee gpurun_out/api_run_overlap_$tag.log | cut -c1-200
python tools/check_call_bench.py > gpurun_out/check_call_$tag.json 2>&1; python -c "
import json
for r in json.load(open('gpurun_out/check_call_$tag.json')): print(r['n_samples'], r['case'], round(r['us_per_check'],1))"
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
