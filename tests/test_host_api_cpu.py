"""Host-side logic of the drop-in Python API that needs no GPU: mirrors the
reference's own tests python/tests/test_ValueMap.py, test_RandomNumberGeneratory.py,
sampling/test_Sampler.py and sampling/test_CompletionCheck.py:5-37, with only
the import root changed."""
import math

import numpy as np
import pytest

import casmcode_monte_b200.monte as monte
import casmcode_monte_b200.monte.ising_cpp as ising
import casmcode_monte_b200.monte.ising_cpp.semigrand_canonical as sgc
import casmcode_monte_b200.monte.sampling as sampling


def test_value_map_round_trip():
    values = monte.ValueMap()
    values.scalar_values["check"] = 1.0
    assert math.isclose(values.scalar_values["check"], 1.0)
    v = np.array([1.0, 1.0])
    values.vector_values["check"] = v
    assert np.allclose(values.vector_values["check"], v)
    values.vector_values["check"][0] = 2.0  # in-place, as with the reference's Eigen views
    assert np.allclose(values.vector_values["check"], np.array([2.0, 1.0]))
    values = monte.ValueMap()
    values.scalar_values["is_scalar"] = 1.0
    values.vector_values["is_vector"] = [2.0, 1.0]
    values.boolean_values["flag"] = True
    data = values.to_dict()
    values_2 = monte.ValueMap.from_dict(data)
    assert len(values_2.scalar_values) == 1 and "is_scalar" in values_2.scalar_values
    assert len(values_2.vector_values) == 1 and np.allclose(values_2.vector_values["is_vector"], [2.0, 1.0])
    assert values_2.boolean_values["flag"] is True
    # from_dict typing rule: bool -> boolean, number -> scalar, array -> vector
    d = monte.ValueMap.from_dict({"temperature": 2000.0, "exchange_potential": [0.0], "m": [[1, 2], [3, 4]]})
    assert "temperature" in d.scalar_values and "exchange_potential" in d.vector_values
    assert d.to_dict()["m"] == [[1.0, 2.0], [3.0, 4.0]]
    inc = monte.ValueMap.from_dict({"temperature": 10.0, "exchange_potential": [0.5]})
    out = d.make_incremented_values(inc, 3)
    assert out.scalar_values["temperature"] == 2030.0 and np.allclose(out.vector_values["exchange_potential"], [1.5])
    assert not d.is_mismatched(inc) and inc.is_mismatched(d)


def test_rng_range_and_reproducibility():
    rng = monte.RandomNumberGenerator()
    for _ in range(10000):
        assert 0 <= rng.random_int(9) <= 9
    e = monte.RandomNumberEngine()
    state = e.dump()
    rng = monte.RandomNumberGenerator(e)
    x = [rng.random_int(9) for _ in range(10)]
    e.load(state)
    assert x == [rng.random_int(9) for _ in range(10)]
    e.load(state)
    r = [rng.random_real(9) for _ in range(10)]
    e.load(state)
    assert r == [rng.random_real(9) for _ in range(10)]


def test_rng_matches_oracle_stream(oracle):
    e = monte.RandomNumberEngine()
    e.seed(4242)
    g = monte.RandomNumberGenerator(e)
    o = oracle.RandomNumberEngine()
    o.seed(4242)
    for _ in range(50):
        assert g.random_int(624) == oracle.random_int(o, 624)
        assert g.random_real(1.0) == oracle.random_real(o, 1.0)
    assert e.dump() == o.dump()


def test_sampler_scalar_vector_matrix():
    sampler = sampling.Sampler(shape=[], component_names=["x"], capacity_increment=10000)
    assert sampler.n_components() == 1 and sampler.n_samples() == 0
    assert sampler.component_names() == ["x"]
    for _ in range(100):
        sampler.append([0.3])
    assert sampler.n_samples() == 100
    sampler.clear()
    for _ in range(100):
        sampler.append(sampling.scalar_as_vector(0.3))
    assert sampler.values().shape == (100, 1)
    sampler.clear()
    assert sampler.sample_capacity() == 10000
    n = 100000
    for _ in range(n):
        sampler.append([0.3])
    assert sampler.n_samples() == n and sampler.values().shape == (n, 1)
    assert sampler.sample_capacity() == n
    s2 = sampling.Sampler(shape=[2], component_names=["x1", "x2"])
    s2.append([0.3, 0.5])
    s2.append(sampling.vector_as_vector([0.3, 0.5]))
    assert s2.values().shape == (2, 2) and np.allclose(s2.component(1), [0.5, 0.5])
    s3 = sampling.Sampler(shape=[2, 2])
    assert s3.component_names() == ["0,0", "1,0", "0,1", "1,1"]
    s3.append(sampling.matrix_as_vector(np.array([[0.1, 0.2], [0.3, 0.4]])))
    assert np.allclose(s3.sample(0), [0.1, 0.3, 0.2, 0.4])  # column-major unrolling
    assert sampling.default_component_names([]) == ["0"]
    assert sampling.default_component_names([3]) == ["0", "1", "2"]
    with pytest.raises(RuntimeError):
        sampling.default_component_names([2, 2, 2])
    with pytest.raises(RuntimeError):
        s3.append([1.0])


def test_completion_check_max_count(tmp_path):
    params = sampling.CompletionCheckParams()
    params.cutoff_params.max_count = 12
    cc = sampling.CompletionCheck(params)
    samplers = sampling.SamplerMap()
    samplers["e"] = sampling.Sampler(shape=[])
    samplers["x"] = sampling.Sampler(shape=[3])
    weight = sampling.Sampler(shape=[])
    log = monte.MethodLog(str(tmp_path / "log.txt"))
    n_steps = 0
    while not cc.count_check(samplers=samplers, sample_weight=weight, count=n_steps, method_log=log):
        n_steps += 1
        if n_steps % 10 == 0:
            samplers["e"].append([0])
            samplers["x"].append([0, 0, 0])
    assert n_steps == 12
    r = cc.results().to_dict()
    assert r["is_complete"] and r["has_any_maximum_met"] and r["count"] == 12 and r["n_samples"] == 1


def test_requested_precision_and_converge_helper():
    rp = sampling.RequestedPrecision(abs=0.001)
    assert rp.abs_convergence_is_required and not rp.rel_convergence_is_required
    assert rp.to_dict() == {"abs_precision": 0.001}
    assert sampling.RequestedPrecision.from_dict({"precision": 0.5, "rel_precision": 0.1}).abs_precision == 0.5
    mc = sgc.SemiGrandCanonicalCalculator(
        system=ising.IsingSystem(
            formation_energy_calculator=ising.IsingFormationEnergy(J=0.1, lattice_type=1),
            param_composition_calculator=ising.IsingParamComposition(),
        )
    )
    fns = mc.default_sampling_functions()
    assert sorted(fns.keys()) == ["formation_energy", "param_composition", "potential_energy"]
    params = sampling.CompletionCheckParams()
    sampling.converge(fns, params).set_precision("potential_energy", abs=0.001).set_precision("param_composition", rel=0.01)
    keys = sorted((k.sampler_name, k.component_index, k.component_name) for k, _ in params.requested_precision.items())
    assert keys == [("param_composition", 0, "0"), ("potential_energy", 0, "0")]
    with pytest.raises(Exception):
        sampling.converge(fns, params).set_precision("nope", abs=1.0)
    with pytest.raises(Exception):
        sampling.converge(fns, params).set_precision("potential_energy")


def test_configuration_host_side():
    c = ising.IsingConfiguration(shape=(25, 25))
    assert c.n_sites == 625 and c.n_variable_sites == 625 and c.n_unitcells == 625
    assert list(c.shape) == [25, 25]
    assert c.occ(7) == 1
    c.set_occ(7, -1)
    assert c.occ(7) == -1 and c.occupation()[7] == -1
    assert c.within(-1, 0) == 24 and c.within(25, 1) == 0
    assert list(c.from_linear_site_index(27)) == [2, 1]
    assert c.to_linear_site_index([2, 1]) == 27
    d = c.to_dict()
    assert d["shape"] == [25, 25] and len(d["occupation"]) == 625
    c2 = ising.IsingConfiguration.from_dict(d)
    assert np.array_equal(c2.occupation(), c.occupation())
    with pytest.raises(RuntimeError):
        ising.IsingConfiguration(shape=[4])  # model.hh:25-27
    with pytest.raises(RuntimeError):
        c.set_occupation(np.ones(3, dtype=int))  # model.hh:56-58
    with pytest.raises(RuntimeError):
        ising.IsingFormationEnergy(J=0.1, lattice_type=2)  # model.hh:175-177
    import copy

    c3 = copy.deepcopy(c)
    c3.set_occ(0, -1)
    assert c.occ(0) == 1
    cond = sgc.SemiGrandCanonicalConditions(temperature=2000.0, exchange_potential=[0.5])
    v = cond.to_values()
    assert v.scalar_values["temperature"] == 2000.0
    assert sgc.SemiGrandCanonicalConditions.from_values(v).exchange_potential[0] == 0.5
    with pytest.raises(RuntimeError):
        sgc.SemiGrandCanonicalConditions.from_values(monte.ValueMap.from_dict({"temperature": 1.0}))


def test_generic_loop_with_python_callbacks_matches_oracle(oracle, tmp_path):
    """methods.basic_occupation_metropolis with Python callbacks
    (python/src/monte_methods.cpp:196-263): a tiny host-side Ising model drives it
    and the trajectory equals the oracle's reference loop on the same engine."""
    import casmcode_monte_b200.monte.events as events
    import casmcode_monte_b200.monte.methods as methods

    n0, n1, Jc, T, mu, seed, n_passes = 6, 4, 0.1, 1200.0, 0.1, 5, 25
    N = n0 * n1
    occ = np.ones(N, dtype=np.int32)
    ev = events.OccEvent()
    ev.linear_site_index.append(0)
    ev.new_occ.append(1)

    def propose(rng):
        ev.linear_site_index[0] = rng.random_int(N - 1)
        ev.new_occ[0] = -int(occ[ev.linear_site_index[0]])
        return ev

    def dpot(e):
        l = e.linear_site_index[0]
        i, j = l % n0, l // n0
        nb = occ[(i + 1) % n0 + n0 * j] + occ[i + n0 * ((j + 1) % n1)] + occ[(i - 1) % n0 + n0 * j] + occ[i + n0 * ((j - 1) % n1)]
        ds = e.new_occ[0] - int(occ[l])
        return (-Jc * ds) * int(nb) - mu * (ds / 2.0)

    def apply(e):
        occ[e.linear_site_index[0]] = e.new_occ[0]

    fns = sampling.StateSamplingFunctionMap()
    fns["param_composition"] = sampling.StateSamplingFunction(
        name="param_composition", description="x", shape=[1], function=lambda: [(N + int(occ.sum())) / 2.0 / N], component_names=["0"]
    )
    params = sampling.CompletionCheckParams()
    params.cutoff_params.max_count = n_passes
    data = methods.BasicOccupationMetropolisData(fns, sampling.jsonStateSamplingFunctionMap(), N, params)
    e = monte.RandomNumberEngine()
    e.seed(seed)
    calls = []
    methods.basic_occupation_metropolis(
        data, T, dpot, propose, apply, sample_period=1, method_log=monte.MethodLog(str(tmp_path / "status.json"), 1e9), random_engine=e,
        write_status_f=lambda d, log: calls.append(d.n_pass),
    )
    oe = oracle.RandomNumberEngine()
    oe.seed(seed)
    ref = oracle.sgc_run([n0, n1], np.ones(N, dtype=np.int32), Jc, T, mu, True, oe, {"max_count": n_passes}, 1)
    assert np.array_equal(occ, ref["occupation"])
    assert (data.n_pass, data.n_accept, data.n_reject) == (ref["n_pass"], ref["n_accept"], ref["n_reject"])
    assert np.array_equal(data.samplers["param_composition"].component(0), ref["samplers"]["param_composition"])
    assert e.dump() == oe.dump() and calls == [n_passes]
    d = data.to_dict()
    assert d["n_steps_per_pass"] == N and abs(d["acceptance_rate"] + d["rejection_rate"] - 1.0) < 1e-15


# ---- constructors and parsers of the completion-check parameters ----
# python/src/monte_sampling.cpp:222-283, :2704-2925; include/casm/monte/checks/io/json/
# CompletionCheck_json_io.hh:36-395; src/casm/monte/checks/io/json/CutoffCheck_json_io.cc:11-92
def _functions(sampling):
    fns = sampling.StateSamplingFunctionMap()
    for name, shape in (("potential_energy", []), ("param_composition", [2])):
        f = sampling.StateSamplingFunction(name=name, description="", shape=shape, function=lambda: [0.0] * (2 if shape else 1))
        fns[f.name] = f
    return fns


def test_completion_check_params_keyword_constructor():
    import casmcode_monte_b200.monte.sampling as sampling

    p = sampling.CompletionCheckParams()
    assert (p.log_spacing, p.check_begin, p.check_period) == (False, 100, 100)
    p = sampling.CompletionCheckParams(check_period=25)
    assert (p.check_begin, p.check_period) == (25, 25)  # begin defaults to the period for linear spacing
    p = sampling.CompletionCheckParams(log_spacing=True)
    assert (p.log_spacing, p.check_begin, p.check_base, p.check_shift, p.check_period_max) == (True, 0, 10.0, 2.0, 10000)
    rp = sampling.RequestedPrecisionMap()
    key = sampling.SamplerComponent(sampler_name="e", component_index=0, component_name="0")
    rp[key] = sampling.RequestedPrecision(abs=1e-3)
    cut = sampling.CutoffCheckParams(min_sample=10, max_count=500)
    p = sampling.CompletionCheckParams(requested_precision=rp, cutoff_params=cut, check_begin=7, check_period=3, check_shift=1.5)
    assert p.cutoff_params.min_sample == 10 and p.cutoff_params.max_count == 500 and p.cutoff_params.min_count is None
    assert (p.check_begin, p.check_period, p.check_shift) == (7, 3, 1.5)
    assert p.requested_precision[key].abs_precision == 1e-3
    calls = []

    def stats(obs, w):
        calls.append(len(obs))
        s = sampling.BasicStatistics()
        s.mean, s.calculated_precision = 1.0, 0.0
        return s

    p = sampling.CompletionCheckParams(requested_precision=rp, calc_statistics_f=stats, check_begin=2, check_period=2)
    cc = sampling.CompletionCheck(p)
    samplers = sampling.SamplerMap()
    samplers["e"] = sampling.Sampler(shape=[], component_names=["0"])
    for v in (1.0, 2.0, 1.0, 2.0):
        samplers["e"].append([v])
    import casmcode_monte_b200.monte as monte

    # the default equilibration check runs on the device: only reached with a GPU
    assert p.to_dict()["convergence"][0]["quantity"] == "e"
    assert monte is not None and cc.params().check_begin == 2 and not calls


def test_completion_check_params_from_dict_round_trip():
    import casmcode_monte_b200.monte.sampling as sampling

    fns = _functions(sampling)
    data = {
        "cutoff": {"count": {"min": 5, "max": 1000}, "sample": {"min": 20}, "clocktime": {"max": 3600.0}},
        "convergence": [
            {"quantity": "potential_energy", "abs_precision": 1e-4},
            {"quantity": "param_composition", "precision": 1e-3, "rel_precision": 0.01, "component_index": [1]},
        ],
        "spacing": "log",
        "begin": 10,
        "base": 4.0,
        "confidence": 0.9,
    }
    p = sampling.CompletionCheckParams.from_dict(data, fns)
    assert p.log_spacing and p.check_begin == 10 and p.check_base == 4.0 and p.check_shift == 2.0 and p.check_period_max == 10000
    c = p.cutoff_params
    assert (c.min_count, c.max_count, c.min_sample, c.max_sample, c.max_clocktime, c.min_time) == (5, 1000, 20, None, 3600.0, None)
    keys = {(k.sampler_name, k.component_index, k.component_name): v for k, v in p.requested_precision.items()}
    assert set(keys) == {("potential_energy", 0, "0"), ("param_composition", 1, "1")}
    r = keys[("param_composition", 1, "1")]
    assert r.abs_convergence_is_required and r.abs_precision == 1e-3 and r.rel_convergence_is_required and r.rel_precision == 0.01
    d = p.to_dict()
    assert d["spacing"] == "log" and d["begin"] == 10 and d["base"] == 4.0 and d["cutoff"]["count"] == {"min": 5, "max": 1000}
    again = sampling.CompletionCheckParams.from_dict(d, fns)
    assert again.to_dict() == d
    # defaults: linear spacing, all components of a quantity
    p = sampling.CompletionCheckParams.from_dict({"convergence": [{"quantity": "param_composition", "precision": 0.1}]}, fns)
    assert (p.log_spacing, p.check_begin, p.check_period) == (False, 100, 100) and len(p.requested_precision) == 2
    p = sampling.CompletionCheckParams.from_dict({"convergence": [{"quantity": "param_composition", "precision": 0.1, "component_name": ["1"]}]}, fns)
    assert [k.component_index for k in p.requested_precision] == [1]
    # errors are collected and raised (RuntimeError, as the reference's report_and_throw_if_invalid)
    for bad in (
        {"convergence": [{"quantity": "nope", "precision": 0.1}]},
        {"convergence": [{"quantity": "param_composition", "precision": 0.1, "component_index": [2]}]},
        {"convergence": [{"quantity": "param_composition", "precision": 0.1, "component_name": ["x"]}]},
        {"convergence": {"quantity": "potential_energy"}},
        {"spacing": "cubic"},
        {"period": 1},
        {"spacing": "log", "base": 1.0},
    ):
        with pytest.raises(RuntimeError):
            sampling.CompletionCheckParams.from_dict(bad, fns)


def test_cutoff_and_statistics_calculator_dict_io():
    import casmcode_monte_b200.monte.sampling as sampling

    c = sampling.CutoffCheckParams(min_count=1, max_sample=9, min_clocktime=0.5)
    assert c.to_dict() == {"count": {"min": 1}, "sample": {"max": 9}, "clocktime": {"min": 0.5}}
    assert sampling.CutoffCheckParams.from_dict(c.to_dict()).to_dict() == c.to_dict()
    assert sampling.CutoffCheckParams.from_dict({}).to_dict() == {}
    calc = sampling.BasicStatisticsCalculator.from_dict({"confidence": 0.8, "n_resamples": 50})
    assert calc.to_dict() == {"confidence": 0.8, "weighted_observations_method": 1, "n_resamples": 50}
    assert sampling.BasicStatisticsCalculator.from_dict(sampling.BasicStatisticsCalculator(0.99, 2, 7).to_dict()).to_dict() == {
        "confidence": 0.99, "weighted_observations_method": 2, "n_resamples": 7}
    r = sampling.RequestedPrecision.from_dict({"precision": 0.5})
    assert r.abs_convergence_is_required and r.abs_precision == 0.5 and not r.rel_convergence_is_required
