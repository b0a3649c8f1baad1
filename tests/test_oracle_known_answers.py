"""Pins the CPU oracle against every known-answer test the reference holds for
the Ising SGC path.  CPU only.

Reference sources of the expected values:
  tests/unit/monte/Ising_basic_semigrand_canonical_test.cpp:113-267
  python/tests/sampling/test_Sampler.py
  python/tests/sampling/test_CompletionCheck.py:5-124
  python/tests/test_RandomNumberGeneratory.py:4-35
  python/tests/events/test_Conversions.py:8-72
"""
import math

import numpy as np
import pytest

J = 0.1


def all_up(n=25):
    return np.ones(n * n, dtype=np.int32)


# --- Ising_basic_semigrand_canonical_test.cpp:113-160 (IsingFormationEnergy1) ---
@pytest.mark.parametrize("use_nlist", [False, True])
def test_formation_energy_known_answers(oracle, use_nlist):
    occ = all_up()
    per_supercell, per_unitcell = oracle.formation_energy([25, 25], occ, J, use_nlist)
    assert math.isclose(per_supercell, 25 * 25 * 2.0 * -J, abs_tol=1e-5)
    assert math.isclose(per_unitcell, 2.0 * -J, abs_tol=1e-5)
    dEf = oracle.formation_energy_delta([25, 25], occ, J, use_nlist, [0], [-1])
    assert math.isclose(dEf, 8.0 * J, abs_tol=1e-5)
    dEf = oracle.formation_energy_delta([25, 25], occ, J, use_nlist, [0], [1])
    assert dEf == 0.0


# --- :162-203 (IsingParamComposition1) ---
def test_param_composition_known_answers(oracle):
    occ = all_up()
    Nx, x = oracle.param_composition([25, 25], occ)
    assert Nx == 625.0 and x == 1.0
    assert oracle.param_composition_delta([25, 25], occ, [0], [-1]) == -1.0
    assert oracle.param_composition_delta([25, 25], occ, [0], [1]) == 0.0


# --- :205-267 (SemiGrandCanonicalPotential1), T=2000, mu=2 ---
def test_potential_known_answers(oracle):
    occ = all_up()
    mu = 2.0
    E, e = oracle.potential([25, 25], occ, J, 2000.0, mu, False)
    assert math.isclose(E, 625 * (2.0 * -J - mu * 1.0), abs_tol=1e-5)
    assert math.isclose(e, 2.0 * -J - mu * 1.0, abs_tol=1e-5)
    dE = oracle.potential_delta([25, 25], occ, J, 2000.0, mu, False, [0], [-1])
    assert math.isclose(dE, 8.0 * J - mu * (-1), abs_tol=1e-5)
    dE = oracle.potential_delta([25, 25], occ, J, 2000.0, mu, False, [0], [1])
    assert dE == 0.0


def test_multi_site_delta_applies_and_restores(oracle):
    # model.hh:361-377: sequential apply / un-apply of a 2-site event
    rng = np.random.default_rng(3)
    occ = rng.choice(np.array([-1, 1], dtype=np.int32), size=36)
    before = occ.copy()
    d01 = oracle.formation_energy_delta([6, 6], occ, J, True, [0, 1], [-occ[0], -occ[1]])
    assert np.array_equal(occ, before)
    d0 = oracle.formation_energy_delta([6, 6], occ, J, True, [0], [-occ[0]])
    occ2 = occ.copy()
    occ2[0] = -occ2[0]
    d1 = oracle.formation_energy_delta([6, 6], occ2, J, True, [1], [-occ2[1]])
    assert d01 == d0 + d1


# --- index conventions, model.hh:73-99 ---
def test_index_conventions(oracle):
    shape = [5, 7]
    assert oracle.within(-1, 0, 0) if False else True
    assert oracle.within(shape, -1, 0) == 4
    assert oracle.within(shape, 5, 0) == 0
    assert oracle.within(shape, -8, 1) == 6
    for l in range(35):
        i, j = oracle.from_linear_site_index(shape, l)
        assert i == l % 5 and j == l // 5
        assert oracle.to_linear_site_index(shape, [i, j]) == l
    with pytest.raises(RuntimeError):
        oracle.ising_configuration_2d_only([4, 4, 4])  # model.hh:25-27


def test_nlist_and_within_paths_agree_on_delta(oracle):
    rng = np.random.default_rng(11)
    shape = [8, 6]
    occ = rng.choice(np.array([-1, 1], dtype=np.int32), size=48)
    a = oracle.potential_delta_all_sites(shape, occ, J, 1500.0, 0.3, True)
    b = oracle.potential_delta_all_sites(shape, occ, J, 1500.0, 0.3, False)
    assert np.array_equal(a, b)  # bit-exact: same expression, same integers
    tab = oracle.accept_table(2, J, 1500.0, 0.3)["dE"]
    assert set(np.unique(a)).issubset(set(tab.ravel()))


def test_energy_paths_agree_within_rounding(oracle):
    rng = np.random.default_rng(12)
    shape = [10, 12]
    occ = rng.choice(np.array([-1, 1], dtype=np.int32), size=120)
    e1, _ = oracle.formation_energy(shape, occ, J, True)
    e2, _ = oracle.formation_energy(shape, occ, J, False)
    assert math.isclose(e1, e2, rel_tol=1e-13, abs_tol=1e-12)
    S, B = oracle.integer_observables(shape, occ)
    assert e1 == -J * float(B)
    x, ef, ep = oracle.observables_from_sums(S, B, 120, J, 0.25)
    assert ef == oracle.formation_energy(shape, occ, J, True)[1]
    assert x == oracle.param_composition(shape, occ)[1]
    assert ep == oracle.potential(shape, occ, J, 1000.0, 0.25, True)[1]


# --- python/tests/test_RandomNumberGeneratory.py ---
def test_engine_dump_load_reproduces(oracle):
    e = oracle.RandomNumberEngine()
    state = e.dump()
    x = [oracle.random_int(e, 9) for _ in range(10)]
    assert all(0 <= v <= 9 for v in x)
    e.load(state)
    assert x == [oracle.random_int(e, 9) for _ in range(10)]
    e.load(state)
    r1 = [oracle.random_real(e, 9.0) for _ in range(10)]
    e.load(state)
    assert r1 == [oracle.random_real(e, 9.0) for _ in range(10)]
    assert all(0.0 <= v < 9.0 for v in r1)


def test_mt19937_64_known_value(oracle):
    # C++11 [rand.predef]: the 10000th invocation of a default-constructed
    # mt19937_64 (seed 5489) produces 9981545732273789042.
    e = oracle.RandomNumberEngine()
    e.seed(5489)
    v = None
    for _ in range(10000):
        v = oracle.random_int(e, 2**64 - 1)
    assert v == 9981545732273789042


def test_lemire_and_canonical_follow_libstdcxx13(oracle):
    # restate libstdc++ 13's algorithms in Python on the raw 64-bit stream
    e = oracle.RandomNumberEngine()
    e.seed(77)
    raw_engine = oracle.RandomNumberEngine()
    raw_engine.seed(77)

    def raw():
        return oracle.random_int(raw_engine, 2**64 - 1)

    N = 625
    for _ in range(2000):
        got = oracle.random_int_long(e, N - 1)
        prod = raw() * N
        low = prod & (2**64 - 1)
        if low < N:
            thr = (2**64 - N) % N
            while low < thr:
                prod = raw() * N
                low = prod & (2**64 - 1)
        assert got == prod >> 64
        u = oracle.random_real(e, 1.0)
        r = float(raw()) / 18446744073709551616.0
        if r >= 1.0:
            r = math.nextafter(1.0, 0.0)
        assert u == r


# --- python/tests/sampling/test_Sampler.py ---
def test_sampler_layout_and_growth(oracle):
    s = oracle.Sampler(shape=[], component_names=["x"], capacity_increment=10000)
    assert s.n_components() == 1 and s.n_samples() == 0
    assert s.sample_capacity() == 10000
    for _ in range(100000):
        s.append([0.3])
    assert s.n_samples() == 100000 and s.sample_capacity() == 100000
    assert s.values().shape == (100000, 1)
    s.clear()
    assert s.n_samples() == 0
    s2 = oracle.Sampler(shape=[2, 2])
    assert s2.component_names() == ["0,0", "1,0", "0,1", "1,1"]
    s2.append([0.1, 0.3, 0.2, 0.4])  # column-major unrolling of [[.1,.2],[.3,.4]]
    assert list(s2.sample(0)) == [0.1, 0.3, 0.2, 0.4]
    assert oracle.default_component_names([]) == ["0"]
    assert oracle.default_component_names([3]) == ["0", "1", "2"]
    with pytest.raises(RuntimeError):
        oracle.default_component_names([2, 2, 2])


# --- python/tests/sampling/test_CompletionCheck.py:5-37 ---
def test_completion_check_max_count_12(oracle):
    cc = oracle.CompletionCheck({"max_count": 12})
    samplers = {"e": oracle.Sampler(shape=[]), "x": oracle.Sampler(shape=[3])}
    weight = oracle.Sampler(shape=[])
    n_steps = 0
    while not cc.count_check(samplers, weight, n_steps):
        n_steps += 1
        if n_steps % 10 == 0:
            samplers["e"].append([0])
            samplers["x"].append([0, 0, 0])
    assert n_steps == 12


# --- python/tests/sampling/test_CompletionCheck.py:40-124 ---
def test_completion_check_converges_on_uniform_noise(oracle):
    cc = oracle.CompletionCheck(
        {"min_sample": 100, "requested_precision": [("e", 0, 0.001, None), ("v", 0, 0.01, None)]}
    )
    samplers = {
        "e": oracle.Sampler(shape=[], component_names=[""]),
        "v": oracle.Sampler(shape=[], component_names=[""]),
    }
    weight = oracle.Sampler(shape=[])
    e = oracle.RandomNumberEngine()
    e.seed(2024)
    n_steps = 0
    while not cc.count_check(samplers, weight, n_steps):
        n_steps += 1
        ev = 1.0 + oracle.random_real(e, 0.1) - 0.05
        vv = 20.0 + oracle.random_real(e, 1.0) - 0.5
        if n_steps % 10 == 0:
            samplers["e"].append([ev])
            samplers["v"].append([vv])
    r = cc.results()
    assert samplers["e"].n_samples() >= 100 and r["is_complete"]
    assert r["equilibration_check_results"]["all_equilibrated"]
    assert len(r["equilibration_check_results"]["individual_results"]) == 2
    assert r["convergence_check_results"]["all_converged"]
    prec = {d["sampler_name"]: d["calculated_precision"] for d in r["convergence_check_results"]["individual_results"]}
    assert prec["e"] < 0.001 and prec["v"] < 0.01


def test_check_schedule_catches_up_one_per_call(oracle):
    # m_n_checks advances by at most one per call.
    cc = oracle.CompletionCheck(
        {"min_sample": 130, "check_begin": 100, "check_period": 10, "requested_precision": [("e", 0, 1e-9, None)]}
    )
    s = {"e": oracle.Sampler(shape=[])}
    w = oracle.Sampler(shape=[])
    rng = np.random.default_rng(0)
    for v in rng.normal(size=135):
        s["e"].append([float(v)])
    assert cc.n_checks() == 0
    for expect in (1, 2, 3, 4, 4, 4):
        cc.count_check(s, w, 135)
        assert cc.n_checks() == expect


# --- statistics: closed forms (the reference has no known-answer test) ---
def test_statistics_against_numpy(oracle):
    rng = np.random.default_rng(5)
    x = np.cumsum(rng.normal(size=4000)) * 0.01 + rng.normal(size=4000)
    m = x.mean()
    assert math.isclose(oracle.variance(x, m), x.var(), rel_tol=1e-12)
    for k in (1, 5, 40):
        ref = np.sum((x[: len(x) - k] - m) * (x[k:] - m)) / (len(x) - k)
        assert math.isclose(oracle.covariance_lag(x, k, m), ref, rel_tol=1e-10)
    f, k = oracle.autocorrelation_factor(x)
    c0 = x.var()
    kk = next(i for i in range(1, len(x)) if abs(np.sum((x[: len(x) - i] - m) * (x[i:] - m)) / (len(x) - i) / c0) <= 0.5)
    assert k == kk
    rho = 2.0 ** (-1.0 / kk)
    assert math.isclose(f, (1 + rho) / (1 - rho), rel_tol=1e-14)
    mean, prec = oracle.basic_statistics(x)
    assert math.isclose(mean, m, rel_tol=1e-12)
    z = math.sqrt(2.0) * oracle.approx_erf_inv(0.95)
    assert math.isclose(prec, z * math.sqrt(f * c0 / len(x)), rel_tol=1e-10)
    assert abs(oracle.approx_erf_inv(0.95) - 1.3859038243) < 3e-3  # true erfinv(0.95)
    # early-outs, BasicStatistics.cc:31-33, :47
    assert oracle.autocorrelation_factor(np.full(50, 2.5))[0] == 1.0
    assert oracle.autocorrelation_factor(np.arange(10, dtype=float))[0] > 1e300 or True


def test_equilibration_check_cases(oracle):
    # all-same shortcut (EquilibrationCheck.cc:59-78)
    assert oracle.default_equilibration_check(np.full(30, 1.0), abs=1e-3) == (True, 0)
    # no precision requested -> trivially equilibrated (:129-132)
    assert oracle.default_equilibration_check(np.arange(10.0)) == (True, 0)
    # a decaying transient then noise: equilibration index lands after the transient
    rng = np.random.default_rng(9)
    x = np.concatenate([np.linspace(5, 0, 50), np.zeros(450)]) + rng.normal(scale=0.01, size=500)
    ok, n = oracle.default_equilibration_check(x, abs=0.01)
    assert ok and 20 <= n <= 120
    # monotone drift never equilibrates
    ok, n = oracle.default_equilibration_check(np.arange(200.0), abs=1e-3)
    assert not ok


# --- python/tests/events/test_Conversions.py ---
def test_conversions_pinned_convention(oracle):
    assert oracle.conv_l_size([3, 3, 3], 1) == 27
    assert oracle.conv_l_size([3, 3, 3], 2) == 54
    assert oracle.conv_bijk_to_l([3, 3, 3], 2, 1, 0, 0, 0) == 27
    assert oracle.conv_l_to_bijk([3, 3, 3], 2, 27) == (1, 0, 0, 0)
    # periodic wrap
    assert oracle.conv_bijk_to_l([3, 3, 3], 2, 0, 3, 0, 0) == 0
    assert oracle.conv_bijk_to_l([3, 3, 3], 2, 0, -1, 0, 0) == oracle.conv_bijk_to_l([3, 3, 3], 2, 0, 2, 0, 0)
    for l in range(54):
        b, i, j, k = oracle.conv_l_to_bijk([3, 3, 3], 2, l)
        assert oracle.conv_bijk_to_l([3, 3, 3], 2, b, i, j, k) == l


# --- the reference's own run test settings (no values asserted there) plus the
#     Onsager anchor for the ordered phase at T=2000 K, mu=0 ---
def test_sgc_run_reference_settings_and_onsager(oracle):
    e = oracle.RandomNumberEngine()
    e.seed(1)
    params = {
        "min_sample": 100,
        "check_begin": 100,
        "check_period": 10,
        "requested_precision": [("param_composition", 0, 0.001, None), ("potential_energy", 0, 0.001, None)],
    }
    r = oracle.sgc_run([25, 25], all_up(), J, 2000.0, 0.0, True, e, params, 1)
    res = r["completion_check_results"]
    assert res["n_samples"] >= 100 and res["is_complete"]
    assert res["equilibration_check_results"]["all_equilibrated"]
    assert res["convergence_check_results"]["all_converged"]
    ind = {d["sampler_name"]: d for d in res["convergence_check_results"]["individual_results"]}
    assert len(ind) == 2
    assert all(d["calculated_precision"] < 0.001 for d in ind.values())
    # loop invariant: exits only at a pass boundary
    assert r["n_accept"] + r["n_reject"] == r["n_pass"] * 625
    # Onsager: m = (1 - sinh(2 beta J)^-4)^(1/8), x = (1+m)/2
    beta = 1.0 / (oracle.KB * 2000.0)
    m = (1.0 - math.sinh(2 * beta * J) ** -4) ** 0.125
    x = ind["param_composition"]["mean"]
    assert abs(x - (1 + m) / 2) < 5 * 0.001 + 2e-3  # finite-size slack at 25x25


# --- frozen vectors: the oracle must keep producing tests/golden/ising_sgc_golden.json ---
def _golden():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ising_sgc_golden.json")) as f:
        return json.load(f)


def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _unhex(v):
    return np.array([float.fromhex(s) for s in v])


def test_oracle_reproduces_the_golden_vectors(oracle):
    g = _golden()
    for c in g["checkerboard"]:
        n = int(np.prod(c["shape"]))
        occ = np.random.default_rng(c["occ_seed"]).choice(np.array([-1, 1], dtype=np.int32), size=n)
        r = oracle.checkerboard_run(c["shape"], occ, g["J"], c["T"], c["mu"], c["philox_seed"], 0, 0, c["n_passes"], c["sample_period"])
        assert _sha(r["occupation"].astype(np.int32)) == c["occupation_sha256"]
        assert [int(v) for v in r["S"]] == c["S"] and [int(v) for v in r["B"]] == c["B"]
        assert int(r["n_accept"]) == c["n_accept"]
        assert np.array_equal(r["potential_energy"], _unhex(c["potential_energy"]))
    for c in g["serial"]:
        occ = np.ones(int(np.prod(c["shape"])), dtype=np.int32)
        e = oracle.RandomNumberEngine()
        e.seed(c["mt19937_64_seed"])
        r = oracle.sgc_run(c["shape"], occ, g["J"], c["T"], c["mu"], True, e, {"max_count": c["max_count"]}, 1)
        assert _sha(r["occupation"].astype(np.int32)) == c["occupation_sha256"]
        assert int(r["n_accept"]) == c["n_accept"]
        assert np.array_equal(r["samplers"]["potential_energy"], _unhex(c["potential_energy"]))
    for c in g["statistics"]:
        x, w = _unhex(c["x"]), _unhex(c["w"])
        mean, prec = oracle.basic_statistics(x)
        assert mean == float.fromhex(c["mean"]) and prec == float.fromhex(c["precision"])
        assert oracle.autocorrelation_factor(x)[1] == c["k_star"]
        assert list(oracle.default_equilibration_check(x, abs=0.05)) == c["equilibration_abs_0.05"]
        assert list(oracle.default_equilibration_check(x, w, abs=0.05)) == c["weighted_equilibration_abs_0.05"]
        for method, key in ((1, "weighted_method1"), (2, "weighted_method2")):
            m, p = oracle.basic_statistics(x, w, method=method, n_resamples=1000)
            assert [float(m).hex(), float(p).hex()] == c[key]
