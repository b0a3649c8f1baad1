"""The statistical half of the parity contract (BASELINE.json north_star):

* checkerboard production runs against the serial reference mode (the reference's
  loop on the reference's mt19937_64 stream) within 3 sigma of the combined error
  bars for energy, composition, heat capacity AND susceptibility, at four (T, mu)
  including one within 2 % of T_c and two with mu != 0;
* BASELINE config 2 as a test: 4096 x 4096 lattices through T_c against Onsager's
  and Yang's exact results, 3-sigma flags asserted at every T that is not T_c;
* the BASELINE-size kernels compared DIRECTLY with the CPU oracle (not only with
  each other): every 2-d variant on 4096 x 4096 and the 3-d kernel on a
  512 x 512 x 8 slab.
Seeds are fixed; tools/validate_ensemble_on_oracle.py ran the same chains through
the oracle (profiles/ensemble_oracle_r2.json).
"""
import math

import numpy as np
import pytest

import _ensemble as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import casmcode_monte_b200 as m

    return m


# ------------------------------------------------ checkerboard vs serial reference ----
def _run_all_chains(cm, mode):
    """All len(CONDITIONS) * M_CHAINS chains in one context (one CTA per chain)."""
    n_chains = len(E.CONDITIONS) * E.M_CHAINS
    lat = cm.IsingLatticeGPU(list(E.SHAPE), n_chains=n_chains, J=E.J)
    for g in range(n_chains):
        T, mu = E.CONDITIONS[g // E.M_CHAINS]
        lat.set_conditions(T, mu, chain=g)
        lat.fill(E.initial_fill(g), chain=g)
    if mode == cm.MODE_SERIAL_REFERENCE:
        for g in range(n_chains):
            lat.seed_mt19937_64(E.MT_SEED0 + g, chain=g)
    else:
        lat.seed_philox(E.PHILOX_SEED)
    lat.run_passes(E.N_EQUIL + E.N_MEASURE, mode, 1)
    out = {}
    for ci, (T, mu) in enumerate(E.CONDITIONS):
        xs = np.stack([lat.samples(cm.Q_PARAM_COMPOSITION, ci * E.M_CHAINS + c)[E.N_EQUIL :] for c in range(E.M_CHAINS)])
        es = np.stack([lat.samples(cm.Q_POTENTIAL_ENERGY, ci * E.M_CHAINS + c)[E.N_EQUIL :] for c in range(E.M_CHAINS)])
        assert xs.shape == (E.M_CHAINS, E.N_MEASURE)
        out[(T, mu)] = E.jackknife(xs, es, T)
    lat.close()
    return out


def test_checkerboard_ensemble_averages_match_serial_reference_within_3_sigma(cm):
    serial = _run_all_chains(cm, cm.MODE_SERIAL_REFERENCE)
    checker = _run_all_chains(cm, cm.MODE_CHECKERBOARD)
    assert any(abs(T / E.T_C - 1.0) <= 0.02 for T, _ in E.CONDITIONS) and sum(mu != 0 for _, mu in E.CONDITIONS) >= 1
    failures = []
    for cond in E.CONDITIONS:
        rows = E.compare(serial[cond], checker[cond], n_sigma=3.0)
        assert {r[0] for r in rows} == {"potential_energy", "param_composition", "heat_capacity", "susceptibility"}
        for name, a, b, sig, ok in rows:
            assert sig > 0 and math.isfinite(a) and math.isfinite(b)
            if not ok:
                failures.append((cond, name, a, b, sig))
    assert not failures, failures
    # the error bars are tight enough for the comparison to mean something: energies to
    # < 1e-3 eV and compositions to < 3e-2 (combined 1-sigma; the widest is the composition
    # 2 % above T_c, where the susceptibility is ~700 / eV) at every condition
    for cond in E.CONDITIONS:
        assert math.hypot(serial[cond]["potential_energy"][1], checker[cond]["potential_energy"][1]) < 1e-3
        assert math.hypot(serial[cond]["param_composition"][1], checker[cond]["param_composition"][1]) < 3e-2


# ----------------------------------------------------- config 2 against Onsager ----
def onsager_energy_per_site(T):
    """Exact internal energy per site of the square-lattice Ising model (Onsager 1944),
    in units where the bond energy is -J s s'."""
    from scipy.special import ellipk

    K = E.J / (E.KB * T)
    k = 2.0 * math.sinh(2.0 * K) / math.cosh(2.0 * K) ** 2
    kp = 2.0 * math.tanh(2.0 * K) ** 2 - 1.0
    return -E.J / math.tanh(2.0 * K) * (1.0 + 2.0 / math.pi * kp * ellipk(k * k))


def yang_magnetisation(T):
    K = E.J / (E.KB * T)
    s = math.sinh(2.0 * K)
    return (1.0 - s**-4) ** 0.125 if s > 1.0 else 0.0


@pytest.mark.parametrize(
    "shape,temps,replicas,n_equil,n_measure,period",
    [
        # away from T_c: short correlation times, the full 4096 x 4096 supercell
        ([4096, 4096], [1800.0, 2200.0, 2500.0, 2800.0, 3200.0], 1, 3000, 20000, 10),
        # 1.3 % below / 1.0 % above T_c: xi ~ 50 lattice constants and tau ~ 10^4 passes, so the
        # run must be ~10^6 passes long for 40 independent blocks; xi << L already holds at
        # 1024 x 1024 (finite-size corrections ~ exp(-L/xi) < 1e-8), four replicas per temperature
        ([1024, 1024], [2600.0, 2660.0], 4, 100000, 1000000, 100),
    ],
)
def test_config2_sweep_against_onsager(cm, shape, temps, replicas, n_equil, n_measure, period):
    """BASELINE configs[1]: mu = 0, the temperatures of the sweep through T_c (T_c
    itself only has to run, see below: its error bars are not Gaussian on any
    affordable run length).  Cold start below T_c, random start above.  Error bars
    by blocking: 40 blocks, each much longer than the correlation time."""
    N = shape[0] * shape[1]
    lat = cm.IsingLatticeGPU(shape, n_chains=len(temps) * replicas, J=E.J)
    lat.seed_philox(0xC0FFEE)
    for c in range(len(temps) * replicas):
        T = temps[c // replicas]
        lat.set_conditions(T, 0.0, chain=c)
        if T > E.T_C:
            lat.randomize(4242 + c, 0.5, chain=c)
    lat.run_passes(n_equil, cm.MODE_CHECKERBOARD, 0)
    lat.run_passes(n_measure, cm.MODE_CHECKERBOARD, period)
    for ti, T in enumerate(temps):
        e_blocks, m_blocks = [], []
        for r in range(replicas):
            S, B = lat.samples_sb(ti * replicas + r)
            assert len(S) == n_measure // period
            nb = 40 // replicas
            n = (len(S) // nb) * nb
            e_blocks += list((-E.J * B[:n].astype(np.float64) / N).reshape(nb, -1).mean(axis=1))
            m_blocks += list((S[:n].astype(np.float64) / N).reshape(nb, -1).mean(axis=1))
        e, se = float(np.mean(e_blocks)), float(np.std(e_blocks, ddof=1) / math.sqrt(len(e_blocks)))
        m, sm = float(np.mean(m_blocks)), float(np.std(m_blocks, ddof=1) / math.sqrt(len(m_blocks)))
        e_exact, m_exact = onsager_energy_per_site(T), yang_magnetisation(T)
        assert abs(e - e_exact) < 3.0 * se + 1e-7, ("energy", T, e, e_exact, se)
        assert abs(m - m_exact) < 3.0 * sm + 1e-7, ("magnetisation", T, m, m_exact, sm)
        assert se < 2e-4 and sm < 2e-2, (T, se, sm)
    lat.close()


def test_config2_critical_point_runs(cm):
    lat = cm.IsingLatticeGPU([4096, 4096], J=E.J)
    lat.set_conditions(2633.0, 0.0)
    lat.seed_philox(0xC0FFEE)
    lat.run_passes(2000, cm.MODE_CHECKERBOARD, 10)
    assert lat.kernel_variant == "ring2d"
    S, B = lat.samples_sb()
    e = -E.J * B[-1] / 4096.0**2
    # relaxing from the ground state towards e_c = -sqrt(2) J
    assert -2 * E.J < e < -math.sqrt(2.0) * E.J and len(S) == 200
    lat.close()


# -------------------------------------------- BASELINE-size kernels vs the oracle ----
def test_full_size_4096_kernels_match_the_oracle_directly(cm, oracle):
    shape = [4096, 4096]
    T, mu, seed, n_passes = 2633.0, 0.03, 0xC0FFEE, 3
    occ = np.random.default_rng(99).choice(np.array([-1, 1], dtype=np.int32), size=shape[0] * shape[1])
    ref = oracle.checkerboard_run(shape, occ, E.J, T, mu, seed, 0, 0, n_passes, 1)
    for variant in ("ring2d", "bulk2d", "tile2d", "auto"):
        lat = cm.IsingLatticeGPU(shape, J=E.J)
        lat.set_conditions(T, mu)
        lat.seed_philox(seed)
        lat.set_kernel_variant(variant)
        lat.upload(occ)
        lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 1)
        if variant != "auto":
            assert lat.kernel_variant == variant
        assert np.array_equal(lat.download(), ref["occupation"]), variant
        S, B = lat.samples_sb()
        assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"]), variant
        assert lat.counters()[1] == ref["n_accept"], variant
        assert np.array_equal(lat.samples(cm.Q_POTENTIAL_ENERGY), ref["potential_energy"]), variant
        lat.close()


def test_full_plane_3d_kernel_matches_the_oracle_directly(cm, oracle):
    shape = [512, 512, 8]  # full 512 x 512 layers of BASELINE config 3
    T, mu, seed, n_passes = 5235.0, 0.05, 77, 3
    occ = np.random.default_rng(5).choice(np.array([-1, 1], dtype=np.int32), size=int(np.prod(shape)))
    ref = oracle.checkerboard_run(shape, occ, E.J, T, mu, seed, 0, 0, n_passes, 1)
    for variant in ("bulk3d", "tma3d", "auto"):
        lat = cm.IsingLatticeGPU(shape, J=E.J)
        lat.set_conditions(T, mu)
        lat.seed_philox(seed)
        lat.set_kernel_variant(variant)
        lat.upload(occ)
        lat.run_passes(n_passes, cm.MODE_CHECKERBOARD, 1)
        assert lat.kernel_variant in ("bulk3d", "tma3d") and (variant == "auto" or lat.kernel_variant == variant)
        assert np.array_equal(lat.download(), ref["occupation"])
        S, B = lat.samples_sb()
        assert np.array_equal(S, ref["S"]) and np.array_equal(B, ref["B"])
        assert lat.counters()[1] == ref["n_accept"]
        lat.close()


# ------------------------------------------------ small contract checks of round 2 ----
def test_compact_host_formats_round_trip(cm):
    for shape in ([64, 48], [25, 25], [32, 4, 6]):
        n = int(np.prod(shape))
        occ = np.random.default_rng(n).choice(np.array([-1, 1], dtype=np.int32), size=n)
        lat = cm.IsingLatticeGPU(shape, n_chains=2, J=E.J)
        lat.upload_i8(occ.astype(np.int8), 1)
        assert np.array_equal(lat.download(1), occ)
        assert np.array_equal(lat.download_i8(1), occ.astype(np.int8))
        bits = np.packbits(occ > 0, bitorder="little")
        assert np.array_equal(lat.download_bits(1), bits)
        lat.upload_bits(bits, 0)
        assert np.array_equal(lat.download(0), occ)
        with pytest.raises(cm.CmgError):
            lat.upload_i8(np.zeros(n, dtype=np.int8))  # values must be +-1
        with pytest.raises(cm.CmgError):
            lat.upload_i8(np.ones(n - 1, dtype=np.int8))
        lat.close()


def test_underflowed_probability_never_accepts(cm, oracle):
    """exp(-dE*beta) == 0: the reference's `rand < prob` never accepts (methods/metropolis.hh:33);
    the table entry is then exact (no 2^-32 floor).  Device == oracle on a random state."""
    shape, T, mu, seed = [64, 48], 1.0, 0.0, 5
    tab = oracle.accept_table(2, E.J, T, mu)
    assert tab["never"].any() and (tab["thr_m1"][tab["never"]] == 0).all()
    occ = np.random.default_rng(8).choice(np.array([-1, 1], dtype=np.int32), size=shape[0] * shape[1])
    ref = oracle.checkerboard_run(shape, occ, E.J, T, mu, seed, 0, 0, 4, 1)
    for variant in ("generic", "bulk2d", "tile2d"):
        lat = cm.IsingLatticeGPU(shape, J=E.J)
        lat.set_conditions(T, mu)
        lat.seed_philox(seed)
        lat.set_kernel_variant(variant)
        lat.upload(occ)
        lat.run_passes(4, cm.MODE_CHECKERBOARD, 1)
        assert np.array_equal(lat.download(), ref["occupation"])
        assert lat.counters()[1] == ref["n_accept"]
        # zero-temperature dynamics: the energy never goes up
        assert np.all(np.diff(lat.samples(cm.Q_FORMATION_ENERGY)) <= 0)
        lat.close()


def test_samples_keep_the_conditions_they_were_taken_under(cm, oracle):
    """A context re-used for a mu sweep: samples already taken are converted with the
    (J, mu) in force when they were taken, also when the double series is re-allocated."""
    shape = [64, 64]
    lat = cm.IsingLatticeGPU(shape, J=E.J)
    lat.seed_philox(3)
    lat.set_conditions(3000.0, 0.05)
    lat.run_passes(600, cm.MODE_CHECKERBOARD, 1)
    S, B = lat.samples_sb()
    lat.set_conditions(3000.0, -0.2)  # no read in between: the first 600 must still use mu = 0.05
    lat.run_passes(1000, cm.MODE_CHECKERBOARD, 1)  # 1600 > 1024: the double series grows
    ep = lat.samples(cm.Q_POTENTIAL_ENERGY)
    S2, B2 = lat.samples_sb()
    n = shape[0] * shape[1]
    first = [oracle.observables_from_sums(int(s), int(b), n, E.J, 0.05)[2] for s, b in zip(S[:5], B[:5])]
    last = [oracle.observables_from_sums(int(s), int(b), n, E.J, -0.2)[2] for s, b in zip(S2[-5:], B2[-5:])]
    assert list(ep[:5]) == first and list(ep[-5:]) == last
    lat.close()
