#!/bin/bash
# records at HEAD after the chained launches: BASELINE configs 2-4 on one GPU, the drop-in run()
tag=${1:-x}
mkdir -p gpurun_out
for c in config3 config2seq config4; do
  timeout 600 python tools/run_baseline_configs.py $c > gpurun_out/${c}_$tag.json 2> gpurun_out/${c}_$tag.err
  tail -c 1500 gpurun_out/${c}_$tag.json; tail -3 gpurun_out/${c}_$tag.err
done
timeout 300 python tools/api_run_bench.py > gpurun_out/api_run_$tag.log 2>&1; tail -5 gpurun_out/api_run_$tag.log
