#!/bin/bash
# 3-d kernel iteration: parity tests that touch the 3-d kernel, throughput, ncu of the kernel as built
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "3d or golden or full_plane or bulk3d or kernels_match" --timeout 150 ) > gpurun_out/pytest3d_$tag.log 2>&1
tail -5 gpurun_out/pytest3d_$tag.log
python tools/sweep_variants.py 3d auto 2>&1 | tee gpurun_out/sweep3d_$tag.json | cut -c1-250
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_halfsweep_bulk3d -s 6 -c 2 -f -o gpurun_out/bulk3d_$tag python tools/profile_target.py 3d 6 bulk3d > gpurun_out/ncu3d_$tag.log 2>&1
python tools/ncu_brief.py gpurun_out/bulk3d_$tag.ncu-rep > gpurun_out/ncu_bulk3d_$tag.txt 2>&1
grep -v "^$" gpurun_out/ncu_bulk3d_$tag.txt | head -40
