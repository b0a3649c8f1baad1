"""Per-call cost of short run_passes calls: ring2d vs tile2d on one 4096^2 lattice."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU

stream = torch.cuda.Stream()
lat = IsingLatticeGPU([4096, 4096], J=0.1)
lat.set_stream(stream.cuda_stream)
lat.set_conditions(2633.0, 0.0)
lat.seed_philox(1)
lat.randomize(5, 0.5)
for variant in ("ring2d", "tile2d"):
    lat.set_kernel_variant(variant)
    for k in (1, 2, 3, 4, 6, 8, 16):
        for _ in range(5):
            lat.run_passes(k, MODE_CHECKERBOARD, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 60
        for _ in range(n):
            lat.run_passes(k, MODE_CHECKERBOARD, 1)
        e1.record(stream)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / n
        print(json.dumps({"variant": variant, "passes_per_call": k, "us_per_call": t * 1e6, "attempts_per_s": 4096 * 4096 * k / t}), flush=True)
