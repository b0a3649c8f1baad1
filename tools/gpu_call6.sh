#!/bin/bash
# 2-d streaming kernel iteration: all GPU tests, sweep of the streaming shapes, N=1 bench
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 ) > gpurun_out/pytest_$tag.log 2>&1
tail -6 gpurun_out/pytest_$tag.log
python tools/sweep_variants.py 2d_sweep8,2d_big,2d_slab8,2d_slab2 2>&1 | tee gpurun_out/sweep2d_$tag.json | cut -c1-220
( timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
grep -o '"value": [0-9.e+]*\|"streaming_8_lattices": {[^}]*}\|"decomposed_lattice_on_one_gpu": {[^}]*}' gpurun_out/bench_$tag.log | cut -c1-300
tail -3 gpurun_out/bench_$tag.err
