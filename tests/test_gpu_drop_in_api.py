"""The drop-in Python API on the GPU.  The first tests are the reference's own
tests for this path with only the import root changed
(python/tests/ising_cpp/test_ising_semigrand_canonical_cpp.py,
 python/tests/ising_cpp/test_ising_cpp_custom_functions.py,
 python/tests/sampling/test_CompletionCheck.py:40-124,
 tests/unit/monte/Ising_basic_semigrand_canonical_test.cpp:113-267);
the rest pin the run() driver against the CPU oracle."""
import json
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1


@pytest.fixture(scope="module")
def api():
    import casmcode_monte_b200.monte as monte
    import casmcode_monte_b200.monte.events as events
    import casmcode_monte_b200.monte.ising_cpp as ising
    import casmcode_monte_b200.monte.ising_cpp.semigrand_canonical as sgc
    import casmcode_monte_b200.monte.methods as methods
    import casmcode_monte_b200.monte.sampling as sampling

    class A:
        pass

    a = A()
    a.monte, a.ising, a.sgc, a.sampling, a.methods, a.events = monte, ising, sgc, sampling, methods, events
    return a


def make_calculator(api, use_nlist=True):
    return api.sgc.SemiGrandCanonicalCalculator(
        system=api.ising.IsingSystem(
            formation_energy_calculator=api.ising.IsingFormationEnergy(J=J, lattice_type=1, use_nlist=use_nlist),
            param_composition_calculator=api.ising.IsingParamComposition(),
        )
    )


def make_state(api, shape, T, mu, occ=None):
    st = api.ising.IsingState(
        configuration=api.ising.IsingConfiguration(shape=shape),
        conditions=api.monte.ValueMap.from_dict({"temperature": T, "exchange_potential": [mu]}),
    )
    if occ is not None:
        st.configuration.set_occupation(occ)
    return st


def reference_test_params(api, fns):
    p = api.sampling.CompletionCheckParams()
    p.cutoff_params.min_sample = 100
    p.log_spacing = False
    p.check_begin = 100
    p.check_period = 10
    api.sampling.converge(fns, p).set_precision("potential_energy", abs=0.001).set_precision("param_composition", abs=0.001)
    return p


def check_reference_run_assertions(api, mc, params):
    samplers = mc.data.samplers
    results = mc.data.completion_check.results()
    json.dumps(results.to_dict(), indent=2)
    assert api.sampling.get_n_samples(samplers) >= 100
    assert results.is_complete
    assert results.equilibration_check_results.all_equilibrated
    assert len(results.equilibration_check_results.individual_results) == 2
    assert results.convergence_check_results.all_converged
    assert len(results.convergence_check_results.individual_results) == 2
    converge_results = results.convergence_check_results.individual_results
    for key, req in params.requested_precision.items():
        assert converge_results[key].stats.calculated_precision < req.abs_precision
    return results


# ---- python/tests/ising_cpp/test_ising_semigrand_canonical_cpp.py ----
@pytest.mark.parametrize("shape", [(25, 25), (24, 24), (64, 64)])  # odd -> serial reference mode, even -> checkerboard
def test_ising_basic_semigrand_canonical_cpp(api, tmp_path, shape):
    mc = make_calculator(api)
    fns = mc.default_sampling_functions()
    jfns = mc.default_json_sampling_functions()
    state = make_state(api, shape, 2000.0, 0.0)
    for l in range(state.configuration.n_sites):
        state.configuration.set_occ(l, 1)
    params = reference_test_params(api, fns)
    log = api.monte.MethodLog(logfile_path=str(tmp_path / "status.json"), log_frequency=0.2)
    mc.run(
        state=state,
        sampling_functions=fns,
        json_sampling_functions=jfns,
        completion_check_params=params,
        event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
        sample_period=1,
        method_log=log,
        random_engine=None,
    )
    res = check_reference_run_assertions(api, mc, params)
    # Onsager anchor (ordered phase at 2000 K): x = (1 + m)/2, m = (1 - sinh(2 beta J)^-4)^(1/8)
    beta = 1.0 / (api.monte.KB * 2000.0)
    m = (1.0 - math.sinh(2 * beta * J) ** -4) ** 0.125
    ind = {k.sampler_name: v for k, v in res.convergence_check_results.individual_results.items()}
    assert abs(ind["param_composition"].stats.mean - (1 + m) / 2) < 7e-3
    # run() leaves the final occupation in the caller's state and counts every attempt
    d = mc.data
    n = state.configuration.n_sites
    assert d.n_accept + d.n_reject == d.n_pass * n
    assert d.to_dict()["n_steps_per_pass"] == n
    x_last = (n + int(np.sum(state.configuration.occupation()))) / 2.0 / n
    assert x_last == mc.data.samplers["param_composition"].component(0)[-1]
    assert len(mc.data.json_samplers["configuration"].to_list()) == api.sampling.get_n_samples(mc.data.samplers)
    assert mc.data.json_samplers["configuration"].to_list()[-1]["occupation"] == list(state.configuration.occupation())
    assert mc.last_kernel == ("serial_reference" if shape[0] % 2 else ("tile2d" if shape[0] % 64 == 0 else "generic"))


# ---- python/tests/ising_cpp/test_ising_cpp_custom_functions.py ----
def test_ising_cpp_custom_python_functions(api, tmp_path):
    mc = make_calculator(api)
    calls = {"status": 0}

    def make_f(name, shape, getter):
        return api.sampling.StateSamplingFunction(name=name, description=name, shape=shape, function=getter, component_names=["0"])

    fns = api.sampling.StateSamplingFunctionMap()
    fns["param_composition"] = make_f("param_composition", [1], lambda: mc.param_composition_calculator.per_unitcell())
    fns["formation_energy"] = make_f("formation_energy", [], lambda: [mc.formation_energy_calculator.per_unitcell()])
    fns["potential_energy"] = make_f("potential_energy", [], lambda: [mc.potential.per_unitcell()])
    jfns = api.sampling.jsonStateSamplingFunctionMap()
    jfns["configuration"] = api.sampling.jsonStateSamplingFunction(
        name="configuration", description="cfg", function=lambda: mc.state.configuration.to_dict()
    )

    def write_status(mc_calculator, method_log):
        calls["status"] += 1
        assert mc_calculator.data.n_pass >= 0
        method_log.begin_lap()

    state = make_state(api, (25, 25), 2000.0, 0.0)
    params = reference_test_params(api, fns)
    log = api.monte.MethodLog(logfile_path=str(tmp_path / "status.json"), log_frequency=0.2)
    e = api.monte.RandomNumberEngine()
    e.seed(99)
    mc.run(
        state=state,
        sampling_functions=fns,
        json_sampling_functions=jfns,
        completion_check_params=params,
        event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
        sample_period=1,
        method_log=log,
        random_engine=e,
        write_status_f=write_status,
    )
    check_reference_run_assertions(api, mc, params)
    assert calls["status"] >= 1
    for key, value in mc.data.json_samplers.items():
        assert isinstance(value.to_list(), list)
    # the Python callbacks saw the same numbers the device path samples: rerun with the defaults
    mc2 = make_calculator(api)
    state2 = make_state(api, (25, 25), 2000.0, 0.0)
    e2 = api.monte.RandomNumberEngine()
    e2.seed(99)
    fns2 = mc2.default_sampling_functions()
    mc2.run(
        state=state2,
        sampling_functions=fns2,
        json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
        completion_check_params=reference_test_params(api, fns2),
        event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
        random_engine=e2,
    )
    for name in ("param_composition", "formation_energy", "potential_energy"):
        assert np.array_equal(mc.data.samplers[name].component(0), mc2.data.samplers[name].component(0))
    assert np.array_equal(state.configuration.occupation(), state2.configuration.occupation())
    assert e.dump() == e2.dump()


# ---- tests/unit/monte/Ising_basic_semigrand_canonical_test.cpp:113-267 through the classes ----
@pytest.mark.parametrize("use_nlist", [False, True])
def test_calculator_known_answers(api, use_nlist):
    state = make_state(api, (25, 25), 2000.0, 2.0)
    f = api.ising.IsingFormationEnergy(J=J, lattice_type=1, use_nlist=use_nlist)
    f.set_state(state)
    assert math.isclose(f.per_supercell(), 25 * 25 * 2.0 * -J, abs_tol=1e-5)
    assert math.isclose(f.per_unitcell(), 2.0 * -J, abs_tol=1e-5)
    assert math.isclose(f.occ_delta_per_supercell([0], [-1]), 8.0 * J, abs_tol=1e-5)
    assert f.occ_delta_per_supercell([0], [1]) == 0.0
    c = api.ising.IsingParamComposition()
    c.set_state(state)
    assert c.n_independent_compositions() == 1
    assert c.per_supercell()[0] == 625.0 and c.per_unitcell()[0] == 1.0
    assert c.occ_delta_per_supercell([0], [-1])[0] == -1.0 and c.occ_delta_per_supercell([0], [1])[0] == 0.0
    pot = api.sgc.SemiGrandCanonicalPotential(system=api.ising.IsingSystem(f, c))
    pot.set_state(state, api.sgc.SemiGrandCanonicalConditions.from_values(state.conditions))
    assert math.isclose(pot.per_supercell(), 625 * (2.0 * -J - 2.0), abs_tol=1e-5)
    assert math.isclose(pot.per_unitcell(), 2.0 * -J - 2.0, abs_tol=1e-5)
    assert math.isclose(pot.occ_delta_per_supercell([0], [-1]), 8.0 * J + 2.0, abs_tol=1e-5)
    assert pot.occ_delta_per_supercell([0], [1]) == 0.0


def test_calculators_match_oracle_on_random_state(api, oracle):
    rng = np.random.default_rng(1)
    shape = [10, 12]
    occ = rng.choice(np.array([-1, 1], dtype=np.int32), size=120)
    state = make_state(api, shape, 1500.0, 0.3, occ)
    for use_nlist in (True, False):
        f = api.ising.IsingFormationEnergy(J=J, lattice_type=1, use_nlist=use_nlist)
        f.set_state(state)
        assert (f.per_supercell(), f.per_unitcell()) == oracle.formation_energy(shape, occ, J, use_nlist)
        for l in (0, 7, 119):
            assert f.occ_delta_per_supercell([l], [-int(occ[l])]) == oracle.formation_energy_delta(shape, occ, J, use_nlist, [l], [-int(occ[l])])
        assert f.occ_delta_per_supercell([0, 1, 10], [-int(occ[0]), -int(occ[1]), -int(occ[10])]) == oracle.formation_energy_delta(
            shape, occ, J, True, [0, 1, 10], [-int(occ[0]), -int(occ[1]), -int(occ[10])]
        )
    assert np.array_equal(state.configuration.occupation(), occ)  # multi-site delta un-applies (model.hh:374-376)
    c = api.ising.IsingParamComposition()
    c.set_state(state)
    assert (c.per_supercell()[0], c.per_unitcell()[0]) == oracle.param_composition(shape, occ)


# ---- the event generator + a host-driven loop (ising_py/test_ising_semigrand_canonical_mixed.py in small) ----
def test_host_driven_loop_matches_oracle(api, oracle):
    shape = [6, 4]
    T, mu, seed, n_passes = 1200.0, 0.1, 5, 30
    occ = np.ones(24, dtype=np.int32)
    state = make_state(api, shape, T, mu, occ)
    mc = make_calculator(api)
    pot = mc.potential
    pot.set_state(state, api.sgc.SemiGrandCanonicalConditions.from_values(state.conditions))
    gen = api.sgc.SemiGrandCanonicalEventGenerator()
    gen.set_state(state)
    e = api.monte.RandomNumberEngine()
    e.seed(seed)
    rng = api.monte.RandomNumberGenerator(e)
    beta = 1.0 / (api.monte.KB * T)
    n_accept = 0
    for _ in range(n_passes * 24):
        ev = gen.propose(rng)
        assert 0 <= ev.linear_site_index[0] < 24 and ev.new_occ[0] in (-1, 1)
        dE = pot.occ_event_delta_per_supercell(ev)
        if api.methods.metropolis_acceptance(dE, beta, rng):
            gen.apply(ev)
            n_accept += 1
    oe = oracle.RandomNumberEngine()
    oe.seed(seed)
    ref = oracle.sgc_run(shape, occ, J, T, mu, True, oe, {"max_count": n_passes}, 1)
    assert np.array_equal(state.configuration.occupation(), ref["occupation"])
    assert n_accept == ref["n_accept"] and e.dump() == oe.dump()


# ---- run() against the oracle ----
@pytest.mark.parametrize("shape,T,mu,use_nlist", [((25, 25), 2000.0, 0.0, True), ((10, 8), 1500.0, 0.2, True), ((12, 10), 2633.0, -0.1, False)])
def test_run_serial_reference_mode_is_trajectory_exact(api, oracle, shape, T, mu, use_nlist):
    n = shape[0] * shape[1]
    occ = np.random.default_rng(2).choice(np.array([-1, 1], dtype=np.int32), size=n)
    oe = oracle.RandomNumberEngine()
    oe.seed(31)
    params_o = {
        "min_sample": 50,
        "check_begin": 50,
        "check_period": 7,
        "max_count": 400,
        "requested_precision": [("param_composition", 0, 0.002, None), ("potential_energy", 0, 0.002, None)],
    }
    ref = oracle.sgc_run(list(shape), occ, J, T, mu, use_nlist, oe, params_o, 2)

    mc = make_calculator(api, use_nlist)
    fns = mc.default_sampling_functions()
    p = api.sampling.CompletionCheckParams()
    p.cutoff_params.min_sample = 50
    p.cutoff_params.max_count = 400
    p.check_begin = 50
    p.check_period = 7
    api.sampling.converge(fns, p).set_precision("potential_energy", abs=0.002).set_precision("param_composition", abs=0.002)
    state = make_state(api, shape, T, mu, occ)
    e = api.monte.RandomNumberEngine()
    e.seed(31)
    mc.run(
        state=state,
        sampling_functions=fns,
        json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
        completion_check_params=p,
        event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
        sample_period=2,
        random_engine=e,
        update_mode="serial_reference",
    )
    d = mc.data
    assert (d.n_pass, d.n_accept, d.n_reject) == (ref["n_pass"], ref["n_accept"], ref["n_reject"])
    assert np.array_equal(state.configuration.occupation(), ref["occupation"])
    for name in ("param_composition", "formation_energy", "potential_energy"):
        a, b = d.samplers[name].component(0), ref["samplers"][name]
        if use_nlist or name == "param_composition":
            assert np.array_equal(a, b)
        else:  # row/column energy form: same expression order, same bits
            assert np.array_equal(a, b)
    assert e.dump() == oe.dump()
    r = d.completion_check.results().to_dict()
    ro = ref["completion_check_results"]
    assert r["is_complete"] == ro["is_complete"] and r["n_samples"] == ro["n_samples"]
    assert d.completion_check.n_checks() == ref["n_checks"]
    if ro["n_samples_at_convergence_check"] is not None:
        assert r["n_samples_at_convergence_check"] == ro["n_samples_at_convergence_check"]
        eq, eqo = r["equilibration_check_results"], ro["equilibration_check_results"]
        assert eq["all_equilibrated"] == eqo["all_equilibrated"]
        if eqo["all_equilibrated"]:
            assert eq["N_samples_for_all_to_equilibrate"] == eqo["N_samples_for_all_to_equilibrate"]
            cv, cvo = r["convergence_check_results"], ro["convergence_check_results"]
            assert cv["all_converged"] == cvo["all_converged"]
            for a, b in zip(cv["individual_results"], cvo["individual_results"]):
                assert a["sampler_name"] == b["sampler_name"]
                assert math.isclose(a["stats"]["mean"], b["mean"], rel_tol=1e-12)
                assert math.isclose(a["stats"]["calculated_precision"], b["calculated_precision"], rel_tol=1e-9)


def test_run_checkerboard_with_the_row_column_energy_form(api, oracle):
    # use_nlist=False: the calculators' row/column sums, sampled on the device
    shape = (64, 48)
    n = shape[0] * shape[1]
    occ = np.random.default_rng(4).choice(np.array([-1, 1], dtype=np.int32), size=n)
    T, mu = 2633.0, 0.05
    mc = make_calculator(api, use_nlist=False)
    fns = mc.default_sampling_functions()
    p = api.sampling.CompletionCheckParams()
    p.cutoff_params.max_count = 10
    state = make_state(api, shape, T, mu, occ)
    e = api.monte.RandomNumberEngine()
    e.seed(78)
    mc.run(state=state, sampling_functions=fns, json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
           completion_check_params=p, event_generator=api.sgc.SemiGrandCanonicalEventGenerator(), sample_period=5, random_engine=e)
    oe = oracle.RandomNumberEngine()
    oe.seed(78)
    philox_seed = oracle.random_int(oe, 2**64 - 1)
    cur = occ
    d = mc.data
    for k in range(2):
        cur = oracle.checkerboard_run(list(shape), cur, J, T, mu, philox_seed, 0, 5 * k, 5, 5)["occupation"]
        assert d.samplers["formation_energy"].component(0)[k] == oracle.formation_energy(list(shape), cur, J, False)[1]
        assert d.samplers["potential_energy"].component(0)[k] == oracle.potential(list(shape), cur, J, T, mu, False)[1]
    assert np.array_equal(state.configuration.occupation(), cur)
    # the calculators agree with what was sampled last
    assert mc.formation_energy_calculator.per_unitcell() == d.samplers["formation_energy"].component(0)[-1]


# (the second runs resident in shared memory, k_ring2d; the third in tiles with halos, k_tile2d; the fourth streams)
@pytest.mark.parametrize("shape", [(64, 48), (1024, 512), (512, 96), (96, 34)])
def test_run_checkerboard_mode_matches_oracle_checkerboard(api, oracle, shape):
    n = shape[0] * shape[1]
    occ = np.random.default_rng(3).choice(np.array([-1, 1], dtype=np.int32), size=n)
    T, mu = 2633.0, 0.05
    mc = make_calculator(api)
    fns = mc.default_sampling_functions()
    p = api.sampling.CompletionCheckParams()
    p.cutoff_params.max_count = 37
    state = make_state(api, shape, T, mu, occ)
    e = api.monte.RandomNumberEngine()
    e.seed(77)
    mc.run(
        state=state,
        sampling_functions=fns,
        json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
        completion_check_params=p,
        event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
        sample_period=3,
        random_engine=e,
    )
    oe = oracle.RandomNumberEngine()
    oe.seed(77)
    philox_seed = oracle.random_int(oe, 2**64 - 1)  # run() seeds Philox with one engine draw
    ref = oracle.checkerboard_run(list(shape), occ, J, T, mu, philox_seed, 0, 0, 37, 3)
    d = mc.data
    assert d.n_pass == 37 and d.n_accept == ref["n_accept"] and d.n_reject == ref["n_reject"]
    assert np.array_equal(state.configuration.occupation(), ref["occupation"])
    assert np.array_equal(d.samplers["potential_energy"].component(0), ref["potential_energy"])
    assert np.array_equal(d.samplers["param_composition"].component(0), ref["param_composition"])
    assert api.sampling.get_n_samples(d.samplers) == 12
    assert d.completion_check.results().has_any_maximum_met


def test_checkerboard_ensemble_matches_serial_reference_within_3_sigma(api):
    """BASELINE north_star: production checkerboard runs match the reference
    ordering's ensemble averages within 3 sigma of the combined error bars."""
    shape = (32, 32)
    out = {}
    for mode in ("serial_reference", "checkerboard"):
        mc = make_calculator(api)
        fns = mc.default_sampling_functions()
        p = api.sampling.CompletionCheckParams()
        p.cutoff_params.min_sample = 2000
        p.cutoff_params.max_sample = 6000
        p.check_begin = 2000
        p.check_period = 500
        api.sampling.converge(fns, p).set_precision("potential_energy", abs=4e-4).set_precision("param_composition", abs=4e-4).set_precision(
            "formation_energy", abs=4e-4
        )
        state = make_state(api, shape, 3000.0, 0.05)
        e = api.monte.RandomNumberEngine()
        e.seed(2024)
        mc.run(
            state=state,
            sampling_functions=fns,
            json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
            completion_check_params=p,
            event_generator=api.sgc.SemiGrandCanonicalEventGenerator(),
            random_engine=e,
            update_mode=mode,
        )
        res = mc.data.completion_check.results()
        out[mode] = {k.sampler_name: (v.stats.mean, v.stats.calculated_precision) for k, v in res.convergence_check_results.individual_results.items()}
        assert len(out[mode]) == 3
    z95 = 1.96
    for name in out["checkerboard"]:
        (m1, p1), (m2, p2) = out["serial_reference"][name], out["checkerboard"][name]
        sigma = math.hypot(p1, p2) / z95  # calculated_precision is a 95 % half-width
        assert abs(m1 - m2) < 3 * sigma, (name, m1, m2, sigma)


# ---- python/tests/sampling/test_CompletionCheck.py:40-124 (statistics on the device) ----
def test_completion_check_converges_on_uniform_noise(api, tmp_path):
    sampling, monte = api.sampling, api.monte
    params = sampling.CompletionCheckParams()
    params.cutoff_params.min_sample = 100
    e_key = sampling.SamplerComponent(sampler_name="e", component_name="", component_index=0)
    params.requested_precision[e_key] = sampling.RequestedPrecision(abs=0.001)
    v_key = sampling.SamplerComponent(sampler_name="v", component_name="", component_index=0)
    params.requested_precision[v_key] = sampling.RequestedPrecision(abs=0.01)
    cc = sampling.CompletionCheck(params)
    samplers = sampling.SamplerMap()
    samplers["e"] = sampling.Sampler(shape=[], component_names=[""])
    samplers["v"] = sampling.Sampler(shape=[], component_names=[""])
    weight = sampling.Sampler(shape=[])
    log = monte.MethodLog(str(tmp_path / "log.txt"))
    rng = monte.RandomNumberGenerator()
    n_steps = 0
    while not cc.count_check(samplers=samplers, sample_weight=weight, count=n_steps, method_log=log):
        n_steps += 1
        e = 1.0 + rng.random_real(0.1) - 0.05
        v = 20.0 + rng.random_real(1.0) - 0.5
        if n_steps % 10 == 0:
            samplers["e"].append([e])
            samplers["v"].append([v])
    results = cc.results()
    assert sampling.get_n_samples(samplers) >= 100 and results.is_complete
    assert results.equilibration_check_results.all_equilibrated
    assert len(results.equilibration_check_results.individual_results) == 2
    assert results.convergence_check_results.all_converged
    cr = results.convergence_check_results.individual_results
    assert cr[e_key].stats.calculated_precision < 0.001 and cr[v_key].stats.calculated_precision < 0.01


def test_statistics_classes_match_oracle(api, oracle):
    rng = np.random.default_rng(4)
    x = np.cumsum(rng.normal(size=2000)) * 0.02 + rng.normal(size=2000) + 3
    calc = api.sampling.BasicStatisticsCalculator(confidence=0.9)
    s = calc(x)
    mean, prec = oracle.basic_statistics(x, confidence=0.9)
    assert math.isclose(s.mean, mean, rel_tol=1e-12) and math.isclose(s.calculated_precision, prec, rel_tol=1e-10)
    assert math.isclose(s.relative_precision(), abs(prec / mean), rel_tol=1e-10)
    r = api.sampling.default_equilibration_check(x, None, api.sampling.RequestedPrecision(abs=0.05))
    assert (r.is_equilibrated, r.N_samples_for_equilibration) == oracle.default_equilibration_check(x, abs=0.05)
    r = api.sampling.default_equilibration_check(x, None, api.sampling.RequestedPrecision(rel=0.02))
    assert (r.is_equilibrated, r.N_samples_for_equilibration) == oracle.default_equilibration_check(x, rel=0.02)
    with pytest.raises(RuntimeError):
        calc([])  # BasicStatistics.cc:116-119


# ---- python/tests/events/test_Conversions.py (index arithmetic subset) ----
def test_conversions_class(api):
    f = api.events.Conversions([3, 3, 3], n_basis=1)
    assert f.l_size() == 27
    f = api.events.Conversions([3, 3, 3], n_basis=2)
    assert f.l_size() == 54
    assert f.bijk_to_l([1, 0, 0, 0]) == 27
    assert list(f.l_to_bijk(27)) == [1, 0, 0, 0]
    assert f.l_to_b(30) == 1 and list(f.l_to_ijk(30)) == [0, 1, 0]
    assert f.bijk_to_l([0, 3, 0, 0]) == 0 and f.bijk_to_l([0, -1, 0, 0]) == f.bijk_to_l([0, 2, 0, 0])
    b = f.l_to_bijk_batch(list(range(54)))
    assert list(f.bijk_to_l_batch(b)) == list(range(54))


@pytest.mark.parametrize("shape", [(64, 64), (512, 96), (1024, 128)])  # one CTA; tiles with halos (two copies of the planes); resident
def test_run_with_overlapped_checks_gives_the_same_results(api, shape):
    """overlap_checks = True (next block enqueued before the pending check, check on a second
    stream, rollback on completion) must change nothing: samplers, counters, completion results
    and the final occupation are those of the waiting loop."""
    out = {}
    for overlap in (False, True):
        mc = make_calculator(api)
        mc.overlap_checks = overlap
        fns = mc.default_sampling_functions()
        p = api.sampling.CompletionCheckParams()
        p.cutoff_params.min_sample = 50
        p.cutoff_params.max_sample = 900
        p.log_spacing = False  # a check every 25 samples from 50 on
        p.check_begin = 50
        p.check_period = 25
        # (oracle, same conditions: the composition reaches 6e-3 after ~250 samples, 2.7e-3 after 900)
        api.sampling.converge(fns, p).set_precision("potential_energy", abs=2e-3).set_precision("param_composition", abs=6e-3)
        state = make_state(api, shape, 3200.0, 0.03)
        e = api.monte.RandomNumberEngine()
        e.seed(77)
        mc.run(state=state, sampling_functions=fns, json_sampling_functions=api.sampling.jsonStateSamplingFunctionMap(),
               completion_check_params=p, event_generator=api.sgc.SemiGrandCanonicalEventGenerator(), random_engine=e,
               update_mode="checkerboard")
        d = mc.data
        res = d.completion_check.results()
        out[overlap] = (
            d.n_pass, d.n_accept, d.n_reject, np.array(state.configuration.occupation()).copy(),
            np.array(d.samplers["potential_energy"].component(0)).copy(), res.is_complete, res.n_samples,
            json.dumps(res.to_dict(), sort_keys=True, default=str).replace(str(res.clocktime), ""),
        )
    a, b = out[False], out[True]
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[5] and b[5] and a[6] == b[6]
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    assert 50 <= a[6] < 900  # converged before the cutoff: the speculative block was rolled back
    # (larger lattices converge sooner; every shape stops at a check, with a block in flight)
