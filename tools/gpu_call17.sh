#!/bin/bash
# ring2d with neighbour-wise waits: parity tests, then A/B against the CTA barrier build
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ring or smem or resident or variants or full_size" --timeout 300 2>&1 | tail -5
( echo "== neighbour-wise"; timeout 300 python tools/sweep_variants.py 2d_ringsync,2d_ringsync_small,2d_ringsync_8192
  echo "== cta barrier"; CMG_LIB_PATH=casmcode_monte_b200/_variants/lib_ctasync.so timeout 300 python tools/sweep_variants.py 2d_ringsync,2d_ringsync_small,2d_ringsync_8192 ) 2>&1 | tee gpurun_out/ringsync_$tag.json
