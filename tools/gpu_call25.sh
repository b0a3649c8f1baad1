#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_drop_in_api.py -m gpu -x -q --timeout 300 -k "overlapped or checkerboard_mode_matches" ) 2>&1 | tail -25
