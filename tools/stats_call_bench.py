"""Cost of one statistics call through the C ABI (series of n doubles from the host)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from casmcode_monte_b200.lattice import host_series_stats, host_series_equilibration
rng = np.random.default_rng(0)
for n in (1000, 10000, 100000):
    x = np.cumsum(rng.normal(size=n)) * 0.01 + rng.normal(size=n)
    host_series_stats(x); host_series_equilibration(x, 1e-3)
    t0 = time.perf_counter()
    for _ in range(50):
        st = host_series_stats(x)
    t1 = time.perf_counter()
    for _ in range(50):
        host_series_equilibration(x, 1e-3)
    t2 = time.perf_counter()
    print(json.dumps({"n": n, "stats_us": (t1 - t0) / 50 * 1e6, "equil_us": (t2 - t1) / 50 * 1e6, "k_star": st["k_star"]}), flush=True)
