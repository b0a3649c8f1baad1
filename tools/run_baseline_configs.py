"""Runs the single-GPU BASELINE.json configs through the C ABI and writes one
JSON record per config (throughput + physics checks) to stdout.

  python tools/run_baseline_configs.py config2   # 4096^2 temperature sweep vs Onsager
  python tools/run_baseline_configs.py config3   # 512^3 run to precision, on-device statistics
  python tools/run_baseline_configs.py config4   # 128 of the 1024 (T, mu) chains of 256^2 (one GPU's share)

(config 1 is the CPU reference case: oracle/oracle_bench; config 5 needs several
GPUs: tools/slab_multi_gpu.py under torchrun.)
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import casmcode_monte_b200 as cm

J = 0.1
KB = cm.KB
Z95 = 1.959964


def onsager_energy_per_site(T):
    """Exact internal energy per site of the infinite square lattice, E = -J sum_<ij> s_i s_j."""
    from scipy.special import ellipk

    K = J / (KB * T)
    kappa = 2.0 * math.sinh(2 * K) / math.cosh(2 * K) ** 2
    return -J / math.tanh(2 * K) * (1.0 + (2.0 / math.pi) * (2.0 * math.tanh(2 * K) ** 2 - 1.0) * ellipk(kappa**2))


def onsager_x(T):
    K = J / (KB * T)
    s = math.sinh(2 * K)
    if s <= 1.0:
        return 0.5
    return 0.5 * (1.0 + (1.0 - s**-4) ** 0.125)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def config2():
    temps = [1800.0, 2200.0, 2500.0, 2600.0, 2633.0, 2660.0, 2800.0, 3200.0]
    n0 = n1 = 4096
    n_eq, n_meas = 1000, 4000
    lat = cm.IsingLatticeGPU([n0, n1], n_chains=len(temps), J=J)
    tc = 2 * J / (KB * math.log(1 + math.sqrt(2)))
    for ch, T in enumerate(temps):
        lat.set_conditions(T, 0.0, chain=ch)
        if T < tc:
            lat.fill(1, chain=ch)  # ordered phase: start inside one magnetisation sector
        else:
            lat.randomize(12345 + ch, 0.5, chain=ch)
    lat.seed_philox(0xC0FFEE)
    lat.run_passes(n_eq, cm.MODE_CHECKERBOARD, 0)
    lat.sync()
    dt = timed(lambda: (lat.run_passes(n_meas, cm.MODE_CHECKERBOARD, 1), lat.sync()))
    rows = []
    for ch, T in enumerate(temps):
        st_e = lat.series_stats(cm.Q_FORMATION_ENERGY, ch)
        st_x = lat.series_stats(cm.Q_PARAM_COMPOSITION, ch)
        e = lat.samples(cm.Q_POTENTIAL_ENERGY, ch)
        x = lat.samples(cm.Q_PARAM_COMPOSITION, ch)
        n = n0 * n1
        row = {
            "T": T,
            "e_formation": st_e["mean"], "e_precision": st_e["calculated_precision"], "e_onsager": onsager_energy_per_site(T),
            "x": st_x["mean"], "x_precision": st_x["calculated_precision"], "x_onsager": onsager_x(T),
            "heat_capacity_per_site_kB": n * float(np.var(e)) / (KB * T * T) / KB,
            "susceptibility_per_site": n * float(np.var(x)) / (KB * T),
            "acceptance": lat.counters(ch)[1] / (lat.counters(ch)[1] + lat.counters(ch)[2]),
        }
        # 3-sigma check against the exact infinite-lattice values away from T_c
        # (|T - T_c| > 150 K: finite-size and critical-slowing effects are negligible there)
        if abs(T - tc) > 150:
            row["e_within_3sigma"] = bool(abs(row["e_formation"] - row["e_onsager"]) < 3 * row["e_precision"] / Z95 + 2e-7)
            row["x_within_3sigma"] = bool(abs(row["x"] - row["x_onsager"]) < 3 * row["x_precision"] / Z95 + 2e-6)
        rows.append(row)
    return {
        "config": "2: 2D 4096x4096 SGC temperature sweep through T_c, 8 temperatures as 8 concurrent lattices, 1000 + 4000 passes, sample every pass",
        "kernel": lat.kernel_variant,
        "attempts_per_s": len(temps) * n0 * n1 * n_meas / dt,
        "seconds_measured": dt,
        "T_c": tc,
        "rows": rows,
    }


def config2seq():
    """The sweep as the reference runs one: ONE 4096^2 supercell heated through
    T_c, every temperature starting from the previous final state."""
    temps = [1800.0, 2200.0, 2500.0, 2600.0, 2633.0, 2660.0, 2800.0, 3200.0]
    n0 = n1 = 4096
    n = n0 * n1
    n_eq, n_meas = 1000, 4000
    tc = 2 * J / (KB * math.log(1 + math.sqrt(2)))
    lat = cm.IsingLatticeGPU([n0, n1], J=J)
    lat.seed_philox(0xC0FFEE)
    lat.fill(1)
    rows, total_dt = [], 0.0
    for T in temps:
        lat.set_conditions(T, 0.0)
        lat.run_passes(n_eq, cm.MODE_CHECKERBOARD, 0)
        lat.clear_samples()
        lat.reset_counters()
        lat.sync()
        dt = timed(lambda: (lat.run_passes(n_meas, cm.MODE_CHECKERBOARD, 1), lat.sync()))
        total_dt += dt
        st_e = lat.series_stats(cm.Q_FORMATION_ENERGY, 0)
        st_x = lat.series_stats(cm.Q_PARAM_COMPOSITION, 0)
        e = lat.samples(cm.Q_POTENTIAL_ENERGY, 0)
        x = lat.samples(cm.Q_PARAM_COMPOSITION, 0)
        row = {
            "T": T, "attempts_per_s": n * n_meas / dt,
            "e_formation": st_e["mean"], "e_precision": st_e["calculated_precision"], "e_onsager": onsager_energy_per_site(T),
            "x": st_x["mean"], "x_precision": st_x["calculated_precision"], "x_onsager": onsager_x(T),
            "heat_capacity_per_site_kB": n * float(np.var(e)) / (KB * T * T) / KB,
            "susceptibility_per_site": n * float(np.var(x)) / (KB * T),
            "acceptance": lat.counters(0)[1] / (lat.counters(0)[1] + lat.counters(0)[2]),
        }
        if abs(T - tc) > 150:
            row["e_within_3sigma"] = bool(abs(row["e_formation"] - row["e_onsager"]) < 3 * row["e_precision"] / Z95 + 2e-7)
            row["x_within_3sigma"] = bool(abs(row["x"] - row["x_onsager"]) < 3 * row["x_precision"] / Z95 + 2e-6)
        rows.append(row)
    return {
        "config": "2 (sequential form): ONE 2D 4096x4096 SGC supercell heated through T_c, 8 temperatures one after the other, 1000 + 4000 passes each, sample every pass",
        "kernel": lat.kernel_variant,
        "attempts_per_s": len(temps) * n * n_meas / total_dt,
        "seconds_measured": total_dt,
        "T_c": tc,
        "rows": rows,
    }


def config3():
    """BASELINE configs[2]: all six (T, mu), each run to abs precision 1e-4 on potential_energy and
    param_composition (cap 2e4 passes) with the whole completion-check loop on the device series:
    a check every 100 samples = one cmg_series_check call (equilibration scan of both series, then
    autocorrelation-aware statistics of the common tail)."""
    shape = [512, 512, 512]
    n = 512**3
    out = []
    quantities = (cm.Q_POTENTIAL_ENERGY, cm.Q_PARAM_COMPOSITION)
    for T in (4000.0, 5235.0, 6500.0):
        for mu in (0.0, 0.05):
            lat = cm.IsingLatticeGPU(shape, J=J)
            lat.set_conditions(T, mu)
            lat.seed_philox(0xC0FFEE)
            lat.fill(1)
            target = 1e-4
            check_begin, check_period, max_count = 100, 100, 20000
            lat.sync()
            t0 = time.perf_counter()
            n_pass, n_checks, t_checks = 0, 0, 0.0
            done = False
            res = None
            while not done and n_pass < max_count:
                lat.run_passes(check_begin if n_pass == 0 else check_period, cm.MODE_CHECKERBOARD, 1)
                n_pass = lat.counters()[0]
                tc0 = time.perf_counter()
                res = lat.series_check(quantities, [target, target])
                t_checks += time.perf_counter() - tc0
                n_checks += 1
                done = all(res["is_equilibrated"]) and res["n_stats"] > 0 and all(p < target for p in res["calculated_precision"])
            lat.sync()
            dt = time.perf_counter() - t0
            st = [lat.series_stats(q, first=max(res["n_equil"])) for q in quantities] if all(res["is_equilibrated"]) else None
            out.append({
                "T": T, "mu": mu, "n_pass": n_pass, "converged": bool(done), "hit_cap": bool(not done), "n_checks": n_checks,
                "seconds": dt, "seconds_in_checks": t_checks, "attempts_per_s_including_checks": n * n_pass / dt,
                "kernel": lat.kernel_variant, "last_check": res,
                "k_star": [s["k_star"] for s in st] if st else None,
            })
            lat.close()
    return {"config": "3: 3D simple-cubic 512^3 SGC, T in {4000, 5235 (T_c), 6500} K x mu in {0, 0.05} eV, run to abs precision 1e-4 on potential_energy and param_composition (cap 2e4 passes) with on-device sampling, equilibration and convergence statistics", "runs": out}


def config4(n_shards=8, variant="auto"):
    from casmcode_monte_b200.parallel import shard_chains

    temps = np.linspace(1500.0, 4000.0, 32)
    mus = np.linspace(-0.2, 0.2, 32)
    conds = [(float(T), float(mu)) for T in temps for mu in mus]
    mine = shard_chains(len(conds), n_shards, 0)  # rank 0 of 8: 128 chains; n_shards = 1: the whole grid on this GPU
    shape = [256, 256]
    lat = cm.IsingLatticeGPU(shape, n_chains=len(mine), J=J)
    lat.set_kernel_variant(variant)
    for local, g in enumerate(mine):
        lat.set_conditions(*conds[g], chain=local)
    lat.seed_philox(0xC0FFEE)
    lat.set_chain_offset(mine[0])
    n_eq, n_meas = 2000, 8000
    lat.run_passes(n_eq, cm.MODE_CHECKERBOARD, 0)
    lat.sync()
    dt = timed(lambda: (lat.run_passes(n_meas, cm.MODE_CHECKERBOARD, 1), lat.sync()))
    t_stats = timed(lambda: lat.series_stats_all(cm.Q_POTENTIAL_ENERGY))
    m, p, v, k = lat.series_stats_all(cm.Q_POTENTIAL_ENERGY)
    mx, px, vx, kx = lat.series_stats_all(cm.Q_PARAM_COMPOSITION)
    return {
        "config": "4: 1024-point (T, mu) grid of independent 256x256 SGC chains; %d chains on this GPU (%s), 2000 + 8000 passes, sample every pass, per-chain statistics on the device" % (len(mine), "one GPU's share of eight" if n_shards == 8 else "the whole grid"),
        "kernel": lat.kernel_variant,
        "n_chains": len(mine),
        "attempts_per_s": len(mine) * 256 * 256 * n_meas / dt,
        "seconds_measured": dt,
        "seconds_device_statistics_all_chains": t_stats,
        "sample": [{"T": conds[g][0], "mu": conds[g][1], "e_pot": float(m[i]), "e_pot_precision": float(p[i]), "x": float(mx[i]), "k_star": int(k[i])} for i, g in list(enumerate(mine))[:: max(1, len(mine) // 8)]],
    }


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "config2"
    print(json.dumps({"config2": config2, "config2seq": config2seq, "config3": config3, "config4": config4,
                      "config4full": lambda: config4(1), "config4full_tile": lambda: config4(1, "tile2d")}[which]()), flush=True)
