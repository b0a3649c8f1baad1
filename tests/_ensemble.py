"""Shared definitions of the ensemble-average parity test (north_star: checkerboard
production runs match the reference ordering's ensemble averages -- energy,
composition, heat capacity, susceptibility -- within 3 sigma of the combined
error bars).

Used by tests/test_gpu_ensemble_statistics.py (both update orders on the GPU)
and by tools/validate_ensemble_on_oracle.py (the same chains through the CPU
oracle, to which both device modes are bit-identical; it is how the fixed seeds
below were checked before the test was committed).

Estimators.  For every condition, M independent chains per update order; after
discarding N_EQUIL passes every pass is a sample.  All estimators pool the M
chains (mean over all samples; population variance about the pooled mean, as
include/casm/monte/misc/math.hh:31-39), so their bias is O(tau / (M n)), and the
error bar is the delete-one-chain jackknife over the M independent chains --
no assumption about the shape of the autocorrelation function.
  e   = <potential_energy>           (per unit cell)
  x   = <param_composition>
  C_v = N Var(potential_energy) / (KB T^2)     
  chi = N Var(param_composition) / (KB T)
"""
import numpy as np

KB = 8.6173303e-05
J = 0.1
T_C = 2.0 * J / (KB * np.log(1.0 + np.sqrt(2.0)))  # 2633.05 K
SHAPE = (64, 64)
N_SITES = SHAPE[0] * SHAPE[1]
M_CHAINS = 32
N_EQUIL = 4000
N_MEASURE = 16000
PHILOX_SEED = 0xC0FFEE
MT_SEED0 = 1000  # chain g of the serial runs is std::mt19937_64(MT_SEED0 + g)

# (T [K], mu [eV]): ordered phase; 2 % above T_c; disordered with mu != 0; ordered with mu != 0
CONDITIONS = [
    (2000.0, 0.0),
    (round(1.02 * T_C, 1), 0.0),
    (3200.0, 0.05),
    (2400.0, -0.1),
]
# Initial state of chain c of a condition: all +1, except that above T_c at mu = 0 the
# chains start alternately all +1 / all -1 -- the magnetisation of a 64 x 64 lattice 2 %
# above T_c relaxes over many thousand passes, and a common start would bias every chain
# of an update order the same way (an equilibration artefact the jackknife cannot see)
ALTERNATE_START = [False, True, False, False]


def initial_fill(g):
    """+1 or -1: the uniform initial occupation of global chain g."""
    return -1 if (ALTERNATE_START[g // M_CHAINS] and (g % M_CHAINS) % 2 == 1) else 1


def observables_from_sb(S, B, T, mu):
    """param_composition and potential_energy per unit cell from the integer sums, with the
    reference's expression order (model.hh:266-270, :412-422; basic_semigrand_canonical.hh:165-174)."""
    S = np.asarray(S, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    e_formation = B * (-J)
    Nx = (N_SITES + S) / 2.0
    return Nx / N_SITES, (e_formation - mu * Nx) / N_SITES


def estimators(x, e, T):
    """x, e: arrays [M, n] of post-equilibration samples -> {name: value}."""
    return {
        "potential_energy": float(e.mean()),
        "param_composition": float(x.mean()),
        "heat_capacity": float(N_SITES * e.var() / (KB * T * T)),
        "susceptibility": float(N_SITES * x.var() / (KB * T)),
    }


def jackknife(x, e, T):
    """{name: (value, standard error)} by deleting one chain at a time."""
    M = x.shape[0]
    full = estimators(x, e, T)
    loo = [estimators(np.delete(x, c, axis=0), np.delete(e, c, axis=0), T) for c in range(M)]
    out = {}
    for k, v in full.items():
        t = np.array([d[k] for d in loo])
        out[k] = (v, float(np.sqrt((M - 1) / M * np.sum((t - t.mean()) ** 2))))
    return out


def compare(serial, checker, n_sigma=3.0):
    """[(name, value_serial, value_checkerboard, combined sigma, ok)]"""
    rows = []
    for k in serial:
        (a, sa), (b, sb) = serial[k], checker[k]
        sig = float(np.hypot(sa, sb))
        rows.append((k, a, b, sig, abs(a - b) < n_sigma * sig))
    return rows
