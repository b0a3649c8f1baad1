"""Does a completion check (second stream) overlap the next block of passes (main stream)?
usage (GPU box): python tools/overlap_probe.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import casmcode_monte_b200 as cm

Q = (cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION)
out = []
for variant in ("ring2d", "bulk2d", "tile2d"):
    for own_stream in (False, True):
        lat = cm.IsingLatticeGPU([4096, 4096], J=0.1)
        if own_stream:
            st = torch.cuda.Stream()
            lat.set_stream(st.cuda_stream)
        lat.set_conditions(2800.0, 0.0)
        lat.seed_philox(1)
        lat.set_kernel_variant(variant)
        lat.run_passes(8000, cm.MODE_CHECKERBOARD, 1)
        lat.series_check(Q, [1e-12, 1e-12])
        lat.sync()
        n = lat.n_samples

        def timed(f, reps=5):
            lat.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                f()
            lat.sync()
            return (time.perf_counter() - t0) / reps * 1e3

        t_block = timed(lambda: lat.run_passes(100, cm.MODE_CHECKERBOARD, 1))
        t_check = timed(lambda: lat.series_check(Q, [1e-12, 1e-12], count=n))

        def both():
            lat.mark()
            lat.run_passes(100, cm.MODE_CHECKERBOARD, 1)
            lat.series_check(Q, [1e-12, 1e-12], count=lat.n_samples - 100)

        t_both = timed(both)
        out.append({"variant": variant, "own_stream": own_stream, "ms_block_100_passes": t_block, "ms_check": t_check,
                    "ms_mark_block_check": t_both, "overlap": t_both < 0.85 * (t_block + t_check)})
        lat.close()
print(json.dumps(out, indent=1))
