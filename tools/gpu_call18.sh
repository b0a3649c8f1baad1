#!/bin/bash
# programmatic dependent launch of the half-sweep kernels: parity tests, then A/B
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
timeout 300 python tools/sweep_variants.py pdl_3d,pdl_2d_sweep8,pdl_2d_one,pdl_3d_256 2>&1 | tee gpurun_out/pdl_$tag.json
