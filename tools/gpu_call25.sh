#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "rollback or host_formats" ) 2>&1 | tail -12
