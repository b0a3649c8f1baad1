"""Fuzz of the device statistics against the oracle: random series of assorted lengths and shapes
(constant, drifting, stepping, heavy-tailed, tiny, near the chunk sizes of the equilibration scan),
equilibration verdict and index exact, k* exact, mean 1e-12, precision 1e-10 (or 1e-14 of the mean).
usage (GPU box): python tools/stats_fuzz.py [n_cases] [seed]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np

import _monte_oracle as oracle
from casmcode_monte_b200.lattice import host_series_equilibration, host_series_stats

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
special = [1, 2, 3, 4, 5, 7, 8, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4096, 4097]
bad = 0
for it in range(n_cases):
    n = int(special[it % len(special)] if it < 3 * len(special) else rng.integers(1, 6000))
    kind = it % 7
    if kind == 0:
        x = rng.normal(size=n)
    elif kind == 1:
        x = np.cumsum(rng.normal(size=n)) * 0.05 + rng.normal(size=n)
    elif kind == 2:
        x = np.concatenate([np.linspace(3, 0, n // 5), np.zeros(n - n // 5)]) + rng.normal(scale=0.01, size=n)
    elif kind == 3:
        x = np.full(n, float(rng.normal())) * (1.0 + (rng.random(n) < 0.01) * 1e-9)
    elif kind == 4:
        x = rng.standard_cauchy(size=n)
    elif kind == 5:
        x = np.where(np.arange(n) < rng.integers(0, n + 1), 1.0, -1.0) + rng.normal(scale=0.1, size=n)
    else:
        x = (rng.integers(0, 3, size=n)).astype(float) * 1e-3 + 0.5
    x = np.ascontiguousarray(x, dtype=np.float64)
    prec = float(10.0 ** rng.uniform(-4, 0))
    ok = True
    try:
        eq_dev, eq_ref = host_series_equilibration(x, prec), oracle.default_equilibration_check(x, abs=prec)
        # The last step of the check walks to the first sample on the other side of the mean of
        # the rest (EquilibrationCheck.cc:107-112).  When a sample EQUALS that mean to the last
        # bits (discrete-valued series) the comparison is decided by the rounding of the partition
        # sums, which the reference leaves to Eigen's reduction order, the oracle takes in sequence
        # and the device reduces as a tree: verdicts must agree, the index only when no sample ties.
        tie = False
        for s0 in range(0, min(len(x), max(eq_dev[1], eq_ref[1]) + 1)):  # the walk starts at or before both answers
            m_rest = float(np.mean(x[s0:]))
            tie = tie or bool(np.any(np.abs(x[s0:] - m_rest) <= 1e-12 * max(1.0, abs(m_rest))))
        ok &= (eq_dev[0] == eq_ref[0]) if tie else (eq_dev == eq_ref)
        st = host_series_stats(x)
        mean, p = oracle.basic_statistics(x)
        f, k = oracle.autocorrelation_factor(x)
        ok &= st["k_star"] == k
        ok &= math.isclose(st["mean"], mean, rel_tol=1e-12, abs_tol=1e-15 * float(np.max(np.abs(x))) + 1e-300)  # (a mean that cancels to ~0)
        if math.isfinite(p) and p < 1e300:
            # (a nearly constant series has a variance of 1e-20 * mean^2: the order of the sums
            # decides its last digits, in the reference's Eigen reductions as much as here)
            ok &= math.isclose(st["calculated_precision"], p, rel_tol=1e-10, abs_tol=1e-14 * abs(mean) + 1e-300)
        else:
            ok &= not (st["calculated_precision"] < 1e300)
    except Exception as e:  # noqa: BLE001
        ok = False
        print("exception", it, n, kind, repr(e)[:200])
    if not ok:
        bad += 1
        print("MISMATCH case", it, "n", n, "kind", kind, "prec", prec, eq_dev, eq_ref,
              "device", host_series_stats(x), "oracle", oracle.basic_statistics(x), oracle.autocorrelation_factor(x), flush=True)
print("cases", n_cases, "bad", bad)
sys.exit(1 if bad else 0)
