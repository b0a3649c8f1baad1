#!/bin/bash
# 3-d kernel iteration: parity tests that touch the 3-d kernels, throughput of both, ncu of tma3d
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "3d or golden or full_plane or kernels_match" --timeout 150 ) > gpurun_out/pytest3d_$tag.log 2>&1
tail -15 gpurun_out/pytest3d_$tag.log
python tools/sweep_variants.py 3d_tma 2>&1 | tee gpurun_out/sweep3d_$tag.json | cut -c1-250
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_halfsweep_tma3d -s 6 -c 2 -f -o gpurun_out/tma3d_$tag python tools/profile_target.py 3d 6 tma3d > gpurun_out/ncu3d_$tag.log 2>&1
python tools/ncu_brief.py gpurun_out/tma3d_$tag.ncu-rep > gpurun_out/ncu_tma3d_$tag.txt 2>&1
grep -v "^$" gpurun_out/ncu_tma3d_$tag.txt | head -80
