#!/bin/bash
# k_tile2d single-tile with neighbour-warp waits: parity tests, then A/B against the CTA-barrier build
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "tile or grid or multichain or golden or smem or drop_in or variants" 2>&1 | tail -5
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print("  %-14s %-18s sp=%d  %.4g att/s  %.2f us" % (d["case"], d["variant"], d["sample_period"], d["attempts_per_s"], d["us_per_halfsweep"]))
    except Exception: print(l, end="")'
( echo "== neighbour-warp waits"; timeout 300 python tools/sweep_variants.py 2d_gridtile,2d_grid | python -c "$fmt"
  echo "== cta barrier"; CMG_LIB_PATH=casmcode_monte_b200/_variants/lib_tilecta.so timeout 300 python tools/sweep_variants.py 2d_gridtile,2d_grid | python -c "$fmt" ) 2>&1 | tee gpurun_out/tilesync_$tag.txt
