"""Two contexts on two streams sharing the GPU (chained streaming launches, tiles with halos, the
resident kernel): each must end where it ends when it has the GPU to itself.
usage (GPU box): python tools/two_streams.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import casmcode_monte_b200 as cm


def make(shape, chains, variant, stream, seed):
    lat = cm.IsingLatticeGPU(shape, n_chains=chains, J=0.1)
    lat.set_stream(stream.cuda_stream)
    for ch in range(chains):
        lat.set_conditions((2633.0 if len(shape) == 2 else 5235.0) + 5.0 * ch, 0.001 * ch, chain=ch)
        lat.randomize(seed + ch, 0.5, chain=ch)
    lat.seed_philox(seed)
    lat.set_kernel_variant(variant)
    return lat


def result(lat, chains):
    lat.sync()
    return [lat.download(ch) for ch in range(chains)], [lat.samples_sb(ch) for ch in range(chains)]


pairs = [(([4096, 4096], 4, "bulk2d"), ([512, 512, 128], 1, "bulk3d")), (([1024, 1024], 8, "tile2d"), ([4096, 2048], 2, "bulk2d")),
         (([4096, 4096], 1, "ring2d"), ([1024, 1024], 4, "tile2d")), (([2048, 2048], 8, "bulk2d"), ([2048, 2048], 8, "bulk2d"))]
bad = 0
for (sa, ca, va), (sb, cb, vb) in pairs:
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    alone = []
    for shape, chains, variant, stream, seed in ((sa, ca, va, s1, 5), (sb, cb, vb, s2, 9)):
        lat = make(shape, chains, variant, stream, seed)
        for _ in range(6):
            lat.run_passes(8, cm.MODE_CHECKERBOARD, 2)
        alone.append(result(lat, chains))
        lat.close()
    la, lb = make(sa, ca, va, s1, 5), make(sb, cb, vb, s2, 9)
    for _ in range(6):  # interleaved enqueues: the two streams' kernels share the GPU
        la.run_passes(8, cm.MODE_CHECKERBOARD, 2)
        lb.run_passes(8, cm.MODE_CHECKERBOARD, 2)
    together = [result(la, ca), result(lb, cb)]
    la.close()
    lb.close()
    same = True
    for (oa, sa_), (ob, sb_) in zip(alone, together):
        same &= all(np.array_equal(x, y) for x, y in zip(oa, ob))
        same &= all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(sa_, sb_))
    bad += 0 if same else 1
    print(json.dumps({"a": [sa, ca, va], "b": [sb, cb, vb], "identical_to_running_alone": bool(same)}), flush=True)
print("FAILED" if bad else "all identical")
sys.exit(1 if bad else 0)
