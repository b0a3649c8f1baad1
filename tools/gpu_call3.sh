#!/bin/bash
# config 3 (all six conditions) + ncu of the 3-d kernel at HEAD
tag=${1:-x}
mkdir -p gpurun_out
( timeout 600 python tools/run_baseline_configs.py config3 ) > gpurun_out/config3_$tag.json 2> gpurun_out/config3_$tag.err
tail -c 1500 gpurun_out/config3_$tag.json; tail -3 gpurun_out/config3_$tag.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_halfsweep_bulk3d -s 8 -c 2 -f -o gpurun_out/bulk3d_$tag python tools/profile_target.py 3d 6 auto > gpurun_out/ncu3d_$tag.log 2>&1
tail -3 gpurun_out/ncu3d_$tag.log
