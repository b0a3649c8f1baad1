#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "slab or ring" --timeout 200 ) > gpurun_out/pytest_$tag.log 2>&1
tail -30 gpurun_out/pytest_$tag.log
