#!/bin/bash
# final records at HEAD: BASELINE configs 2 (sequential form), 3, 4 on one GPU + the ncu captures of gpu_call4.sh
tag=${1:-x}
mkdir -p gpurun_out
for c in config2seq config3 config4; do
  timeout 600 python tools/run_baseline_configs.py $c > gpurun_out/${c}_$tag.json 2> gpurun_out/${c}_$tag.err
  tail -c 400 gpurun_out/${c}_$tag.json; echo; tail -2 gpurun_out/${c}_$tag.err
done
bash tools/gpu_call4.sh $tag
