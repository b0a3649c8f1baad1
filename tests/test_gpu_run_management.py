"""The run-management driver on the GPU:
occupation_metropolis(mc_calculator, state, run_manager) with several sampling
fixtures.  In serial_reference mode it must reproduce the CPU restatement of the
reference loop (oracle/run_management_oracle.hh) sample for sample; in
checkerboard mode it is checked against the low-level C-ABI lattice driven at the
same counts."""
import json
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

J = 0.1


@pytest.fixture(scope="module")
def api():
    import casmcode_monte_b200.monte as monte
    import casmcode_monte_b200.monte.ising_cpp as ising
    import casmcode_monte_b200.monte.ising_cpp.semigrand_canonical as sgc
    import casmcode_monte_b200.monte.run_management as rm
    import casmcode_monte_b200.monte.sampling as sampling

    class A:
        pass

    a = A()
    a.monte, a.ising, a.sgc, a.sampling, a.rm = monte, ising, sgc, sampling, rm
    return a


def make_calculator(api, use_nlist=True):
    return api.sgc.SemiGrandCanonicalCalculator(
        system=api.ising.IsingSystem(
            formation_energy_calculator=api.ising.IsingFormationEnergy(J=J, lattice_type=1, use_nlist=use_nlist),
            param_composition_calculator=api.ising.IsingParamComposition(),
        )
    )


def make_state(api, shape, T, mu, occ):
    st = api.ising.IsingState(
        configuration=api.ising.IsingConfiguration(shape=shape),
        conditions=api.monte.ValueMap.from_dict({"temperature": T, "exchange_potential": [mu]}),
    )
    st.configuration.set_occupation(occ)
    return st


NAMES = ["param_composition", "formation_energy", "potential_energy"]


def two_fixtures(api, mc, stochastic=False, results_io=None):
    """'thermo': every pass, run to precision with a count cutoff; 'log': log-spaced
    samples with the trajectory.  Returns (params list, the same as oracle dicts)."""
    fns = mc.default_sampling_functions()
    afs = api.rm.ResultsAnalysisFunctionMap()
    for f in (api.rm.make_heat_capacity_f(mc), api.rm.make_susceptibility_f(mc)):
        afs[f.name] = f
    sp1 = api.sampling.SamplingParams(sampler_names=NAMES, stochastic_sample_period=stochastic, period=2.0 if stochastic else 1.0)
    cc1 = api.sampling.CompletionCheckParams()
    cc1.cutoff_params.min_sample = 40
    cc1.cutoff_params.max_count = 300
    cc1.check_begin = 40
    cc1.check_period = 9
    api.sampling.converge(fns, cc1).set_precision("potential_energy", abs=0.002).set_precision("param_composition", abs=0.002)
    sp2 = api.sampling.SamplingParams(
        sampler_names=NAMES[:2], sample_method=api.sampling.SAMPLE_METHOD.LOG, begin=0.0, base=1.6, shift=1.0,
        do_sample_trajectory=True,
    )
    cc2 = api.sampling.CompletionCheckParams()
    cc2.cutoff_params.max_count = 1000
    jfs = api.sampling.jsonStateSamplingFunctionMap()
    params = [
        api.rm.SamplingFixtureParams("thermo", fns, jfs, afs, sp1, cc1, analysis_names=["heat_capacity", "susceptibility"],
                                     results_io=results_io),
        api.rm.SamplingFixtureParams("log", fns, jfs, afs, sp2, cc2),
    ]
    oracle_fixtures = [
        {
            "label": "thermo",
            "sampling_params": {"sampler_names": NAMES, "period": sp1.period, "begin": sp1.begin,
                                "stochastic_sample_period": stochastic},
            "completion_check_params": {
                "min_sample": 40, "max_count": 300, "check_begin": 40, "check_period": 9,
                "requested_precision": [("param_composition", 0, 0.002, None), ("potential_energy", 0, 0.002, None)],
            },
            "analysis_names": ["heat_capacity", "susceptibility"],
        },
        {
            "label": "log",
            "sampling_params": {"sampler_names": NAMES[:2], "sample_method": "log", "begin": 0.0, "base": 1.6,
                                "shift": 1.0, "do_sample_trajectory": True},
            "completion_check_params": {"max_count": 1000},
        },
    ]
    return params, oracle_fixtures


@pytest.mark.parametrize("stochastic", [False, True])
@pytest.mark.parametrize("shape,T,mu,use_nlist", [((25, 25), 2000.0, 0.0, True), ((16, 12), 2633.0, 0.03, False)])
def test_serial_reference_mode_reproduces_the_restated_loop(api, oracle, shape, T, mu, use_nlist, stochastic):
    n = shape[0] * shape[1]
    occ = np.random.default_rng(4).choice(np.array([-1, 1], dtype=np.int32), size=n)
    mc = make_calculator(api, use_nlist)
    params, oracle_fixtures = two_fixtures(api, mc, stochastic)
    oe = oracle.RandomNumberEngine()
    oe.seed(77)
    ref = oracle.run_management_sgc_run(list(shape), occ, J, T, mu, use_nlist, oe, oracle_fixtures, True)

    e = api.monte.RandomNumberEngine()
    e.seed(77)
    run_manager = api.rm.RunManager(e, params, global_cutoff=True)
    state = make_state(api, shape, T, mu, occ)
    api.rm.occupation_metropolis(mc, state, run_manager, update_mode="serial_reference")

    assert np.array_equal(state.configuration.occupation(), ref["occupation"])
    assert e.dump() == oe.dump()  # the engine is left exactly where the reference loop leaves it
    assert math.isclose(state.properties.scalar_values["potential_energy"], ref["potential_energy_property"], rel_tol=1e-9)
    for fixture, want in zip(run_manager.sampling_fixtures, ref["fixtures"]):
        r = fixture.results()
        assert fixture.label() == want["label"]
        assert list(r.sample_count) == want["sample_count"]
        assert (r.n_accept, r.n_reject) == (want["n_accept"], want["n_reject"])
        c = fixture.counter()
        assert (c.count, c.pass_, c.step) == (want["count"], want["pass"], want["step"])
        for name, values in want["samplers"].items():
            assert np.array_equal(r.samplers[name].component(0), values), name
        assert len(r.sample_trajectory) == len(want["sample_trajectory"])
        for a, b in zip(r.sample_trajectory, want["sample_trajectory"]):
            assert np.array_equal(a, b)
        got_cc, want_cc = r.completion_check_results.to_dict(), want["completion_check_results"]
        assert got_cc["is_complete"] == want_cc["is_complete"] and got_cc["n_samples"] == want_cc["n_samples"]
        assert got_cc["count"] == want_cc["count"]
        for name, value in want["analysis"].items():
            assert math.isclose(r.analysis[name][0], value[0], rel_tol=1e-10), name
    thermo = run_manager.sampling_fixtures[0].results()
    assert set(thermo.analysis) == {"heat_capacity", "susceptibility"} and thermo.analysis["heat_capacity"][0] > 0
    assert thermo.completion_check_results.is_complete  # global cutoff: the first complete fixture ends the run


@pytest.mark.parametrize("shape", [(64, 48), (512, 96)])  # one CTA per lattice; tiles with halos (two copies of the planes)
def test_checkerboard_mode_samples_at_the_scheduled_counts(api, tmp_path, shape):
    """Checkerboard update order: the fixtures' samples equal what the C-ABI lattice
    gives when driven to the same pass counts with the same Philox key."""
    import casmcode_monte_b200 as cm

    T, mu = 2500.0, 0.01
    n = shape[0] * shape[1]
    occ = np.random.default_rng(8).choice(np.array([-1, 1], dtype=np.int32), size=n)
    mc = make_calculator(api)
    io = api.rm.jsonResultsIO(tmp_path / "results", write_trajectory=False, write_observations=True)
    params, _ = two_fixtures(api, mc, results_io=io)
    e = api.monte.RandomNumberEngine()
    e.seed(5)
    e2 = api.monte.RandomNumberEngine()
    e2.seed(5)
    philox_seed = api.monte.RandomNumberGenerator(e2).random_int(2**64 - 1)  # the driver's one draw
    run_manager = api.rm.RunManager(e, params, global_cutoff=False)
    state = make_state(api, shape, T, mu, occ)
    api.rm.occupation_metropolis(mc, state, run_manager)  # auto -> checkerboard (even extents)
    thermo, log = (f.results() for f in run_manager.sampling_fixtures)
    assert mc.last_kernel in ("generic", "bulk2d", "tile2d")
    assert list(log.sample_count) == [2, 3, 4, 7, 10, 17, 27, 43, 69, 110, 176, 281, 450, 721]
    assert run_manager.sampling_fixtures[1].counter().count == 1000  # both fixtures had to complete
    assert list(thermo.sample_count) == list(range(1, 1001))
    assert thermo.n_accept + thermo.n_reject == 1000 * n

    lat = cm.IsingLatticeGPU(list(shape), J=J)
    lat.set_conditions(T, mu)
    lat.seed_philox(philox_seed)
    lat.upload(occ)
    done = 0
    for k, count in enumerate(log.sample_count):
        lat.run_passes(count - done, cm.MODE_CHECKERBOARD, 0)
        done = count
        S, B = lat.sample_now()
        assert log.samplers["param_composition"].component(0)[k] == (n + S) / 2.0 / n
        assert log.samplers["formation_energy"].component(0)[k] == (-J * B) / n
        assert np.array_equal(log.sample_trajectory[k], lat.download())
        assert thermo.samplers["param_composition"].component(0)[count - 1] == (n + S) / 2.0 / n
    lat.run_passes(1000 - done, cm.MODE_CHECKERBOARD, 0)
    assert np.array_equal(state.configuration.occupation(), lat.download())
    assert lat.counters()[1] == thermo.n_accept

    # results files, reference layout (jsonResultsIO_impl.hh)
    s = json.load(open(tmp_path / "results" / "summary.json"))
    assert s["conditions"]["temperature"]["value"] == [T]
    assert s["conditions"]["exchange_potential"]["0"] == [mu]
    for name in NAMES:
        v = s["statistics"][name]["value" if name != "param_composition" else "0"]
        assert len(v["mean"]) == 1 and len(v["calculated_precision"]) == 1
    assert s["statistics"]["potential_energy"]["value"]["is_converged"] in ([True], [False])
    c = s["completion_check_results"]
    assert c["N_samples"] == [1000] and c["count"] == [1000] and len(c["all_equilibrated"]) == 1
    assert set(s["analysis"]) == {"heat_capacity", "susceptibility"}
    obs = json.load(open(tmp_path / "results" / "run.0" / "observations.json"))
    assert obs["count"] == list(range(1, 1001)) and obs["param_composition"]["component_names"] == ["0"]


def test_by_step_schedules_need_pass_boundaries(api):
    shape, T, mu = (8, 8), 2000.0, 0.0
    occ = np.ones(64, dtype=np.int32)
    mc = make_calculator(api)
    fns = mc.default_sampling_functions()

    def run(period, max_count):
        sp = api.sampling.SamplingParams(sampler_names=NAMES[:1], sample_mode=api.sampling.SAMPLE_MODE.BY_STEP, period=period)
        cc = api.sampling.CompletionCheckParams()
        cc.cutoff_params.max_count = max_count
        p = api.rm.SamplingFixtureParams("s", fns, api.sampling.jsonStateSamplingFunctionMap(),
                                         api.rm.ResultsAnalysisFunctionMap(), sp, cc)
        m = api.rm.RunManager(api.monte.RandomNumberEngine(), [p])
        api.rm.occupation_metropolis(mc, make_state(api, shape, T, mu, occ), m)
        return m.sampling_fixtures[0]

    f = run(128.0, 640)  # every second pass, counted in steps
    assert list(f.results().sample_count) == [128, 256, 384, 512, 640] and f.counter().pass_ == 10
    with pytest.raises(RuntimeError, match="pass boundary"):
        run(100.0, 640)
    sp = api.sampling.SamplingParams(sampler_names=NAMES[:1], sample_mode=api.sampling.SAMPLE_MODE.BY_TIME)
    p = api.rm.SamplingFixtureParams("t", fns, api.sampling.jsonStateSamplingFunctionMap(),
                                     api.rm.ResultsAnalysisFunctionMap(), sp, api.sampling.CompletionCheckParams())
    with pytest.raises(RuntimeError, match="BY_TIME"):
        api.rm.occupation_metropolis(mc, make_state(api, shape, T, mu, occ),
                                     api.rm.RunManager(api.monte.RandomNumberEngine(), [p]))


import casmcode_monte_b200.monte as monte
import casmcode_monte_b200.monte.ising_cpp as ising
import casmcode_monte_b200.monte.run_management as rm
import casmcode_monte_b200.monte.sampling as sampling


def _fixture_params(label, sp, cc, results_io=None, analysis_functions=None, analysis_names=(), fns=None):
    sp.sampler_names = ["x"]
    return rm.SamplingFixtureParams(
        label, fns, sampling.jsonStateSamplingFunctionMap(),
        analysis_functions if analysis_functions is not None else rm.ResultsAnalysisFunctionMap(), sp, cc,
        analysis_names=list(analysis_names), results_io=results_io,
    )


def _state():
    config = ising.IsingConfiguration([2, 2])
    return ising.IsingState(config, monte.ValueMap.from_dict({"temperature": 1000.0, "exchange_potential": [0.0]}))


def test_json_results_io_layout(tmp_path):
    """summary.json / observations.json with the reference's keys
    (jsonResultsIO_impl.hh:33-365): one array element appended per run."""
    values = iter(np.linspace(0.0, 1.0, 1000))
    fns = sampling.StateSamplingFunctionMap()
    fns["x"] = sampling.StateSamplingFunction("x", "ramp", [], lambda: np.array([next(values)]))
    afs = rm.ResultsAnalysisFunctionMap()
    afs["twice_mean"] = rm.ResultsAnalysisFunction(
        "twice_mean", "2 <x>", [], lambda results: np.array([2.0 * np.mean(results.samplers["x"].component(0))]))
    afs["broken"] = rm.ResultsAnalysisFunction("broken", "raises", [], lambda results: 1 / 0)
    io = rm.jsonResultsIO(tmp_path / "out", write_trajectory=True, write_observations=True)
    assert io.to_dict()["kwargs"]["write_observations"] is True
    cc = sampling.CompletionCheckParams()
    cc.cutoff_params.max_count = 8
    sp = sampling.SamplingParams(period=2.0, do_sample_trajectory=True)
    params = _fixture_params("thermo", sp, cc, results_io=io, analysis_functions=afs,
                             analysis_names=["twice_mean", "broken", "unknown"], fns=fns)
    m = rm.RunManager(monte.RandomNumberEngine(), [params])
    state = _state()
    for run_index in range(2):
        m.run_index = run_index
        m.initialize(4)
        m.sample_data_by_count_if_due(state)
        while not m.is_complete():
            m.increment_step()
            m.increment_n_accept()
            m.sample_data_by_count_if_due(state)
        m.finalize(state)
    r = m.sampling_fixtures[0].results()
    assert list(r.sample_count) == [2, 4, 6, 8] and r.acceptance_rate() == 1.0
    assert set(r.analysis) == {"twice_mean", "broken"} and math.isnan(r.analysis["broken"][0])
    s = json.load(open(tmp_path / "out" / "summary.json"))
    assert set(s) == {"conditions", "statistics", "completion_check_results", "analysis"}
    assert s["conditions"]["temperature"] == {"shape": [], "value": [1000.0, 1000.0]}
    assert s["conditions"]["exchange_potential"] == {"shape": [1], "component_names": ["0"], "0": [0.0, 0.0]}
    st = s["statistics"]["x"]
    assert st["shape"] == [] and len(st["value"]["mean"]) == 2 and len(st["value"]["calculated_precision"]) == 2
    assert "is_converged" not in st["value"]  # nothing was requested to converge
    c = s["completion_check_results"]
    assert c["N_samples"] == [4, 4] and c["N_samples_for_statistics"] == [4, 4] and c["count"] == [8, 8]
    assert c["acceptance_rate"] == [1.0, 1.0] and len(c["elapsed_clocktime"]) == 2 and "all_equilibrated" not in c
    assert s["analysis"]["twice_mean"]["shape"] == [] and len(s["analysis"]["twice_mean"]["value"]) == 2
    obs = json.load(open(tmp_path / "out" / "run.1" / "observations.json"))
    assert obs["count"] == [2, 4, 6, 8] and obs["x"]["shape"] == [] and len(obs["x"]["value"]) == 4
    assert len(obs["clocktime"]) == 4
    traj = json.load(open(tmp_path / "out" / "run.0" / "trajectory.json"))
    assert traj == [[1, 1, 1, 1]] * 4
