#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile2d -s 3 -c 1 -f -o gpurun_out/tile_mid_$tag python tools/profile_target.py mid 30 tile2d > gpurun_out/ncu_tile_mid_$tag.log 2>&1
tail -2 gpurun_out/ncu_tile_mid_$tag.log
python tools/ncu_brief.py gpurun_out/tile_mid_$tag.ncu-rep > gpurun_out/ncu_tile_mid_$tag.txt 2>&1
cat gpurun_out/ncu_tile_mid_$tag.txt | head -50
rm -f gpurun_out/tile_mid_$tag.ncu-rep
