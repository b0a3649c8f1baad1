"""Run the chains of tests/test_gpu_ensemble_statistics.py through the CPU oracle
(test infrastructure) and print the 3-sigma table.  Both device update orders are
bit-identical to the oracle's restatements, so this predicts the GPU test's
outcome for the committed seeds without a GPU.

usage: python tools/validate_ensemble_on_oracle.py [n_procs] > profiles/ensemble_oracle_r2.json
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _ensemble as E  # noqa: E402


def run_chain(job):
    import _monte_oracle as orc

    mode, g = job
    T, mu = E.CONDITIONS[g // E.M_CHAINS]
    occ = np.full(E.N_SITES, E.initial_fill(g), dtype=np.int32)
    n_pass = E.N_EQUIL + E.N_MEASURE
    if mode == "serial":
        eng = orc.RandomNumberEngine()
        eng.seed(E.MT_SEED0 + g)
        r = orc.sgc_run(list(E.SHAPE), occ, E.J, T, mu, True, eng, {"max_count": n_pass}, 1)
        x = np.asarray(r["samplers"]["param_composition"]).ravel()
        e = np.asarray(r["samplers"]["potential_energy"]).ravel()
    else:
        r = orc.checkerboard_run(list(E.SHAPE), occ, E.J, T, mu, E.PHILOX_SEED, g, 0, n_pass, 1)
        x, e = E.observables_from_sb(r["S"], r["B"], T, mu)
    assert x.size == n_pass
    return mode, g, x[E.N_EQUIL :], e[E.N_EQUIL :]


def main():
    n_procs = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    jobs = [(m, g) for m in ("serial", "checkerboard") for g in range(len(E.CONDITIONS) * E.M_CHAINS)]
    with mp.Pool(n_procs) as pool:
        results = pool.map(run_chain, jobs, chunksize=1)
    data = {m: {} for m in ("serial", "checkerboard")}
    for m, g, x, e in results:
        data[m][g] = (x, e)
    report = []
    ok = True
    for ci, (T, mu) in enumerate(E.CONDITIONS):
        est = {}
        for m in data:
            xs = np.stack([data[m][ci * E.M_CHAINS + c][0] for c in range(E.M_CHAINS)])
            es = np.stack([data[m][ci * E.M_CHAINS + c][1] for c in range(E.M_CHAINS)])
            est[m] = E.jackknife(xs, es, T)
        for name, a, b, sig, good in E.compare(est["serial"], est["checkerboard"]):
            report.append({"T": T, "mu": mu, "quantity": name, "serial_reference": a, "checkerboard": b, "sigma": sig,
                           "n_sigma": abs(a - b) / sig if sig > 0 else 0.0, "within_3_sigma": bool(good)})
            ok &= bool(good)
    print(json.dumps({"shape": E.SHAPE, "chains": E.M_CHAINS, "n_equil": E.N_EQUIL, "n_measure": E.N_MEASURE, "all_within_3_sigma": ok, "rows": report}, indent=1))


if __name__ == "__main__":
    main()
