#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
echo "== CTA-barrier build"
CMG_LIB_PATH=casmcode_monte_b200/_variants/lib_ctabar.so timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "test_tile2d_matches_oracle or test_ring2d_matches_oracle" > gpurun_out/racecheck_cta_$tag.log 2>&1
echo rc=$?; grep -c "Race reported" gpurun_out/racecheck_cta_$tag.log; tail -2 gpurun_out/racecheck_cta_$tag.log
echo "== tma3d (mbarrier full/empty ring, round-2 kernel) in the product build"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "test_tma3d_matches_oracle" > gpurun_out/racecheck_tma_$tag.log 2>&1
echo rc=$?; grep -c "Race reported" gpurun_out/racecheck_tma_$tag.log; tail -2 gpurun_out/racecheck_tma_$tag.log
