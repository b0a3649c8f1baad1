"""Cost of one completion check on the device series (cmg_series_check) as a function of the
series length, for a series that never equilibrates (the scan runs to the end) and one that does.
usage (GPU box): python tools/check_call_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casmcode_monte_b200 as cm

lat = cm.IsingLatticeGPU([1024, 1024], J=0.1)
lat.set_conditions(2800.0, 0.0)
lat.seed_philox(1)
lat.run_passes(20000, cm.MODE_CHECKERBOARD, 1)
lat.sync()
out = []
for n in (100, 1000, 5000, 10000, 20000):
    for prec, label in ((1e-12, "never equilibrates"), (1e-3, "equilibrates at once")):
        lat.series_check((cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION), [prec, prec], count=n)
        t0 = time.perf_counter()
        for _ in range(20):
            r = lat.series_check((cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION), [prec, prec], count=n)
        dt = (time.perf_counter() - t0) / 20
        out.append({"n_samples": n, "case": label, "us_per_check": dt * 1e6, "n_equil": r["n_equil"], "n_stats": r["n_stats"]})
print(json.dumps(out, indent=1))
