"""Regenerates tests/golden/ising_sgc_golden.json from the CPU oracle
(oracle/monte_oracle.hh).  The reference itself cannot be built in this image
(DESIGN.md section 1), so these vectors do not come from libcasm-monte: they
freeze what the restatement -- pinned to the reference's known answers by
tests/test_oracle_known_answers.py -- produces, so that neither the oracle nor
the device path can drift unnoticed between rounds.

  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import _monte_oracle as orc  # noqa: E402

J = 0.1


def occ_of(shape, seed):
    n = int(np.prod(shape))
    return np.random.default_rng(seed).choice(np.array([-1, 1], dtype=np.int32), size=n)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexf(x):
    return [float(v).hex() for v in np.atleast_1d(x)]


def main():
    out = {"J": J, "checkerboard": [], "serial": [], "statistics": []}
    for shape, T, mu, seed, n_passes, period in [
        ([64, 48], 2633.0, 0.05, 0xC0FFEE, 6, 2),
        ([1024, 128], 2633.0, 0.02, 424242, 5, 1),
        ([100, 100], 2000.0, 0.0, 7, 4, 1),
        ([32, 10, 8], 5235.0, 0.05, 99, 3, 1),
    ]:
        occ = occ_of(shape, 1000 + len(out["checkerboard"]))
        r = orc.checkerboard_run(shape, occ, J, T, mu, seed, 0, 0, n_passes, period)
        out["checkerboard"].append({
            "shape": shape, "T": T, "mu": mu, "philox_seed": seed, "n_passes": n_passes, "sample_period": period,
            "occ_seed": 1000 + len(out["checkerboard"]),
            "occupation_sha256": digest(r["occupation"].astype(np.int32)),
            "S": [int(v) for v in r["S"]], "B": [int(v) for v in r["B"]],
            "n_accept": int(r["n_accept"]),
            "potential_energy": hexf(r["potential_energy"]), "param_composition": hexf(r["param_composition"]),
        })
    for shape, T, mu, seed, max_count in [([25, 25], 2000.0, 0.0, 12345, 20), ([10, 14], 1500.0, 0.1, 5, 15)]:
        occ = np.ones(int(np.prod(shape)), dtype=np.int32)
        e = orc.RandomNumberEngine()
        e.seed(seed)
        r = orc.sgc_run(shape, occ, J, T, mu, True, e, {"max_count": max_count}, 1)
        out["serial"].append({
            "shape": shape, "T": T, "mu": mu, "mt19937_64_seed": seed, "max_count": max_count,
            "occupation_sha256": digest(r["occupation"].astype(np.int32)),
            "n_accept": int(r["n_accept"]),
            "param_composition": hexf(r["samplers"]["param_composition"]),
            "potential_energy": hexf(r["samplers"]["potential_energy"]),
        })
    rng = np.random.default_rng(2024)
    for n in (50, 501, 1500):
        x = np.cumsum(rng.normal(size=n)) * 0.05 + rng.normal(size=n) + 2.0
        w = rng.exponential(size=n) + 1e-3
        mean, prec = orc.basic_statistics(x)
        f, k = orc.autocorrelation_factor(x)
        wm1, wp1 = orc.basic_statistics(x, w, method=1, n_resamples=1000)
        wm2, wp2 = orc.basic_statistics(x, w, method=2, n_resamples=1000)
        out["statistics"].append({
            "x": hexf(x), "w": hexf(w), "mean": float(mean).hex(), "precision": float(prec).hex(), "k_star": int(k),
            "equilibration_abs_0.05": list(orc.default_equilibration_check(x, abs=0.05)),
            "weighted_equilibration_abs_0.05": list(orc.default_equilibration_check(x, w, abs=0.05)),
            "weighted_method1": [float(wm1).hex(), float(wp1).hex()],
            "weighted_method2": [float(wm2).hex(), float(wp2).hex()],
        })
    with open(os.path.join(HERE, "ising_sgc_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "ising_sgc_golden.json"))


if __name__ == "__main__":
    main()
