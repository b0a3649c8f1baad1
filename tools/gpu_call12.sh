#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
for rep in 1 2; do bash tools/ab_sweep.sh "2d_ring,2d_sweep8,2d_slab8,3d" ${@:2}; done 2>&1 | tee gpurun_out/ab_ties_$tag.txt | grep -v "rp=32\|rp=256\|tile2d\|js=\|ns=" 
