#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "neighbour_warp" 2>&1 | tail -8
