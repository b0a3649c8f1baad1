"""What the tie path costs: the hot loops are branch-free except for threshold ties, which only occur
for acceptance probabilities strictly between 0 and 1 -- at T = 1 K every table entry is 0 or 1.
usage (GPU box): python tools/tie_cost.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from casmcode_monte_b200 import MODE_CHECKERBOARD, IsingLatticeGPU

stream = torch.cuda.Stream()
for shape, chains, variant, n_passes in (([4096, 4096], 1, "ring2d", 512), ([4096, 4096], 8, "bulk2d", 48), ([512, 512, 512], 1, "bulk3d", 20)):
    for T in (2633.0 if len(shape) == 2 else 5235.0, 1.0):
        lat = IsingLatticeGPU(shape, n_chains=chains, J=0.1)
        lat.set_stream(stream.cuda_stream)
        lat.set_conditions(T, 0.0)
        lat.seed_philox(1)
        lat.randomize(5, 0.5)
        lat.set_kernel_variant(variant)
        for sp in (0, 1):
            lat.run_passes(10, MODE_CHECKERBOARD, sp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            lat.run_passes(n_passes, MODE_CHECKERBOARD, sp)
            e1.record(stream)
            torch.cuda.synchronize()
            n = chains
            for s in shape:
                n *= s
            print(json.dumps({"shape": shape, "chains": chains, "variant": variant, "T": T, "sample_period": sp,
                              "attempts_per_s": n * n_passes / (e0.elapsed_time(e1) * 1e-3)}), flush=True)
        lat.close()
