"""Throughput of the k-state coloured sweep (3 species, 4096^2 and 256^3). usage (GPU box): python tools/kstate_rate.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import casmcode_monte_b200 as cm

V3 = np.array([[0.0, 0.03, 0.05], [0.03, 0.0, 0.02], [0.05, 0.02, 0.0]])
for shape in ([4096, 4096], [256, 256, 256]):
    lat = cm.IsingLatticeGPU(shape)
    lat.kstate_set_model(V3)
    lat.kstate_set_conditions(2500.0, [0.0, 0.01, -0.01])
    lat.seed_philox(1)
    n = int(np.prod(shape))
    lat.kstate_upload(np.random.default_rng(1).integers(0, 3, size=n).astype(np.int32))
    lat.kstate_run_passes(5, cm.MODE_CHECKERBOARD, 0)
    lat.sync()
    t0 = time.perf_counter()
    lat.kstate_run_passes(40, cm.MODE_CHECKERBOARD, 0)
    lat.sync()
    dt = time.perf_counter() - t0
    print(json.dumps({"shape": shape, "species": 3, "attempts_per_s": n * 40 / dt}), flush=True)
    lat.close()
