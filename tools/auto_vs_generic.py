"""Robustness sweep: for an assortment of shapes and chain counts, the kernel `auto` selects (with its
launch form: resident, chained, programmatic dependent, plain) must leave the same occupation, samples
and counters as the generic byte kernel.  usage (GPU box): python tools/auto_vs_generic.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import casmcode_monte_b200 as cm

cases = [([2048, 2048], 1, 24), ([8192, 8192], 1, 6), ([1024, 1024], 16, 12), ([512, 512, 64], 1, 10), ([128, 128, 128], 2, 12),
         ([64, 64], 512, 30), ([100, 100], 3, 10), ([4096, 64], 1, 20), ([512, 512], 1, 20), ([768, 2048], 2, 12),
         ([256, 256], 300, 20), ([96, 34], 5, 10), ([2048, 64, 16], 1, 10), ([16384, 2048], 1, 6), ([32, 6], 1, 9)]
if len(sys.argv) > 1 and sys.argv[1] == "more":  # a second assortment: many chains on the resident kernel, ragged sizes, long thin lattices
    cases = [([1024, 256], 40, 12), ([2048, 512], 7, 10), ([4096, 512], 3, 9), ([8192, 64], 64, 8), ([1024, 4096], 2, 9),
             ([64, 4096], 6, 12), ([192, 300], 9, 10), ([320, 64], 100, 16), ([1024, 130], 1, 14), ([8192, 2050], 1, 6),
             ([256, 32, 64], 5, 10), ([512, 64, 64], 3, 8), ([1024, 32, 32], 2, 8), ([64, 64, 64], 9, 10), ([4096, 4096], 2, 6),
             ([128, 100], 148, 14), ([448, 448], 4, 10), ([2, 2], 3, 5), ([6, 4, 2], 2, 6)]
bad = 0
for shape, chains, n_passes in cases:
    out = {}
    for variant in ("auto", "generic"):
        lat = cm.IsingLatticeGPU(shape, n_chains=chains, J=0.1)
        for ch in range(chains):
            lat.set_conditions((2633.0 if len(shape) == 2 else 5235.0) + 7.0 * ch, 0.002 * (ch % 5), chain=ch)
            lat.randomize(11 + ch, 0.5, chain=ch)
        lat.seed_philox(2024)
        lat.set_kernel_variant(variant)
        lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 2)
        lat.run_passes(3, cm.MODE_CHECKERBOARD, 1)
        lat.sync()
        pick = [0, chains // 2, chains - 1]
        out[variant] = (lat.kernel_variant, [lat.download(ch) for ch in pick], [lat.samples_sb(ch) for ch in pick], [lat.counters(ch) for ch in pick])
        lat.close()
    a, g = out["auto"], out["generic"]
    same = all(np.array_equal(x, y) for x, y in zip(a[1], g[1])) and all(
        np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a[2], g[2])) and a[3] == g[3]
    bad += 0 if same else 1
    print(json.dumps({"shape": shape, "chains": chains, "passes": n_passes + 3, "auto_kernel": a[0], "identical_to_generic": bool(same)}), flush=True)
print("FAILED" if bad else "all identical")
sys.exit(1 if bad else 0)
