#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "kstate" --timeout 200 2>&1 | tail -4
python tools/kstate_rate.py 2>&1 | tee gpurun_out/kstate_rate_$tag.json
