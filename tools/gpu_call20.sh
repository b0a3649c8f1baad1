#!/bin/bash
# ncu --set full captures at HEAD (round 2, chained launches): ring2d, bulk2d on 8 x 4096^2, bulk3d on 512^3; launch list of the bench
tag=${1:-x}
mkdir -p gpurun_out
cap() {  # name kernel-regex skip count case passes variant
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1_$tag python tools/profile_target.py $5 $6 $7 > gpurun_out/ncu_$1_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$1_$tag.log
  python tools/ncu_brief.py gpurun_out/$1_$tag.ncu-rep > gpurun_out/ncu_$1_$tag.txt 2>&1
  [ "$1" = ring2d ] || rm -f gpurun_out/$1_$tag.ncu-rep
}
cap ring2d k_ring2d 1 1 2d 300 auto
cap bulk2d_sweep8 k_halfsweep_bulk2d 8 2 sweep8 6 bulk2d
cap bulk3d k_halfsweep_bulk3d 8 2 3d 6 auto
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$tag.log 2>&1 )
tail -1 gpurun_out/launches_bench_$tag.csv | cut -c1-200
