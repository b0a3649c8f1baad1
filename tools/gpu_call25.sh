#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_two_streams.py -m gpu -x -q --timeout 300 ) 2>&1 | tail -8
