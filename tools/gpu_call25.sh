#!/bin/bash
( time timeout 900 python -m pytest tests/test_gpu_run_management.py -m gpu -x -q --timeout 300 ) 2>&1 | tail -8
