"""Run management without a GPU: sampling schedules, counters,
fixtures, the run manager's completion logic and the JSON results layout, each
against the CPU restatement in oracle/run_management_oracle.hh (parity
unpinned: the reference holds no golden vector for this path) and against the
values the reference's formulas give by hand."""
import json
import math

import numpy as np
import pytest

import casmcode_monte_b200.monte as monte
import casmcode_monte_b200.monte.ising_cpp as ising
import casmcode_monte_b200.monte.run_management as rm
import casmcode_monte_b200.monte.sampling as sampling

MODE = {sampling.SAMPLE_MODE.BY_PASS: "pass", sampling.SAMPLE_MODE.BY_STEP: "step"}
METHOD = {sampling.SAMPLE_METHOD.LINEAR: "linear", sampling.SAMPLE_METHOD.LOG: "log"}


def as_oracle(p):
    return {
        "sample_mode": MODE[p.sample_mode], "sample_method": METHOD[p.sample_method], "period": p.period,
        "begin": p.begin, "base": p.base, "shift": p.shift, "stochastic_sample_period": p.stochastic_sample_period,
    }


def test_sampling_params_defaults_follow_the_reference_binding():
    # python/src/monte_sampling.cpp:49-88: begin defaults to period (LINEAR) or 0 (LOG)
    p = sampling.SamplingParams()
    assert (p.sample_mode, p.sample_method) == (sampling.SAMPLE_MODE.BY_PASS, sampling.SAMPLE_METHOD.LINEAR)
    assert (p.begin, p.period, p.shift) == (1.0, 1.0, 0.0) and p.base == 10.0 ** (1.0 / 10.0)
    assert not (p.stochastic_sample_period or p.do_sample_trajectory or p.do_sample_time)
    assert sampling.SamplingParams(period=5.0).begin == 5.0
    assert sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.LOG).begin == 0.0
    with pytest.raises(RuntimeError, match="CUSTOM"):
        sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.CUSTOM)
    p = sampling.SamplingParams(sampler_names=["a"])
    p.append_to_sampler_names("b")
    p.extend_sampler_names(["c", "d"])
    p.remove_from_sampler_names("a")
    p.remove_from_sampler_names("not there")
    assert list(p.sampler_names) == ["b", "c", "d"]
    p.append_to_json_sampler_names("j")
    p.extend_json_sampler_names(["k"])
    p.remove_from_json_sampler_names("j")
    assert list(p.json_sampler_names) == ["k"]


def test_sample_at_matches_oracle_and_formulas(oracle):
    # SamplingParams.hh:229-245
    cases = [
        sampling.SamplingParams(),
        sampling.SamplingParams(period=10.0),
        sampling.SamplingParams(period=2.5, begin=0.0),
        sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.LOG),
        sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.LOG, begin=3.0, base=2.0, shift=1.0),
        sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.LOG, base=1.7, shift=10.0),
    ]
    for p in cases:
        for i in list(range(40)) + [1000, 123456]:
            got = rm.sample_at(i, p)
            assert got == oracle.rm_sample_at(i, as_oracle(p)) == p.sample_at(i)
            if p.sample_method == sampling.SAMPLE_METHOD.LINEAR:
                assert got == p.begin + p.period * float(i)
    assert [rm.sample_at(i, sampling.SamplingParams()) for i in range(3)] == [1.0, 2.0, 3.0]
    log10 = sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.LOG)
    assert math.isclose(rm.sample_at(10, log10), 10.0) and math.isclose(rm.sample_at(20, log10), 100.0)
    custom = sampling.SamplingParams(sample_method=sampling.SAMPLE_METHOD.CUSTOM, custom_sample_at=lambda n: 3.0 * n * n)
    assert [rm.sample_at(i, custom) for i in range(4)] == [0.0, 3.0, 12.0, 27.0]


@pytest.mark.parametrize("mode", [sampling.SAMPLE_MODE.BY_PASS, sampling.SAMPLE_MODE.BY_STEP])
def test_monte_counter_matches_oracle(oracle, mode):
    # SamplingFixture.hh:79-119
    c = rm.MonteCounter()
    c.reset(mode, 7)
    trace = []
    for _ in range(50):
        c.increment_step()
        trace.append([c.step, c.pass_, c.count])
    assert trace == [list(t) for t in oracle.rm_monte_counter_trace(MODE[mode], 7, 50)]
    # a block of whole passes == that many single steps
    a, b = rm.MonteCounter(), rm.MonteCounter()
    a.reset(mode, 7)
    b.reset(mode, 7)
    for _ in range(3 * 7):
        a.increment_step()
    b.advance_passes(3, 5, 16)
    assert (a.step, a.pass_, a.count) == (b.step, b.pass_, b.count) == (0, 3, 21 if mode == sampling.SAMPLE_MODE.BY_STEP else 3)
    assert (b.n_accept, b.n_reject) == (5, 16)
    a.increment_step()
    with pytest.raises(RuntimeError, match="pass boundary"):
        a.advance_passes(1, 0, 0)


def test_stochastic_count_step_matches_oracle(oracle):
    # SamplingParams.hh:247-259: geometric waiting time, one random_real per trial
    e = monte.RandomNumberEngine()
    e.seed(1234)
    rng = monte.RandomNumberGenerator(e)
    got = [rm.stochastic_count_step(0.2, rng) for _ in range(200)]
    oe = oracle.RandomNumberEngine()
    oe.seed(1234)
    assert got == oracle.rm_stochastic_count_steps(oe, 0.2, 200)
    assert abs(np.mean(got) - 5.0) < 1.0 and min(got) >= 1


def _constant_functions(value=1.0):
    fns = sampling.StateSamplingFunctionMap()
    fns["x"] = sampling.StateSamplingFunction("x", "constant", [], lambda: np.array([value]))
    return fns


def _fixture_params(label, sp, cc, results_io=None, analysis_functions=None, analysis_names=(), fns=None):
    sp.sampler_names = ["x"]
    return rm.SamplingFixtureParams(
        label, fns if fns is not None else _constant_functions(), sampling.jsonStateSamplingFunctionMap(),
        analysis_functions if analysis_functions is not None else rm.ResultsAnalysisFunctionMap(), sp, cc,
        analysis_names=list(analysis_names), results_io=results_io,
    )


def _state():
    config = ising.IsingConfiguration([2, 2])
    return ising.IsingState(config, monte.ValueMap.from_dict({"temperature": 1000.0, "exchange_potential": [0.0]}))


SCHEDULES = [
    (dict(), dict(max_count=30), 3),
    (dict(period=4.0), dict(max_count=41), 5),
    (dict(period=4.0, begin=0.0), dict(max_count=20), 2),
    (dict(sample_method=sampling.SAMPLE_METHOD.LOG, base=2.0), dict(max_count=70), 3),
    (dict(sample_method=sampling.SAMPLE_METHOD.LOG, shift=10.0), dict(max_sample=25), 2),
    (dict(sample_mode=sampling.SAMPLE_MODE.BY_STEP, period=7.0), dict(max_count=100), 4),
    (dict(stochastic_sample_period=True, period=3.0), dict(max_count=200), 3),
    (dict(stochastic_sample_period=True, sample_method=sampling.SAMPLE_METHOD.LOG, base=1.5, begin=1.0), dict(max_sample=12), 3),
]


@pytest.mark.parametrize("sp_kwargs,cutoffs,steps_per_pass", SCHEDULES)
def test_fixture_schedule_matches_oracle(oracle, sp_kwargs, cutoffs, steps_per_pass):
    """A fixture driven step by step (SamplingFixture.hh:141-189, :335-548) samples at
    the same counts and completes at the same step as the restated reference."""
    sp = sampling.SamplingParams(**sp_kwargs)
    cc = sampling.CompletionCheckParams()
    for k, v in cutoffs.items():
        setattr(cc.cutoff_params, k, v)
    e = monte.RandomNumberEngine()
    e.seed(99)
    f = rm.SamplingFixture(_fixture_params("schedule", sp, cc), e)
    state = _state()
    f.initialize(steps_per_pass)
    f.sample_data_by_count_if_due(state)
    n = 0
    while not f.is_complete() and n < 100000:
        assert f.steps_to_next_event() >= 1
        f.increment_step()
        f.sample_data_by_count_if_due(state)
        n += 1
    oe = oracle.RandomNumberEngine()
    oe.seed(99)
    want = oracle.rm_fixture_schedule(as_oracle(sp), dict(cutoffs), steps_per_pass, 100000, oe)
    assert list(f.results().sample_count) == want["sample_count"]
    assert (n, f.counter().count, f.is_complete()) == (want["steps"], want["count"], want["is_complete"])
    assert f.results().samplers["x"].n_samples() == len(want["sample_count"])


def test_steps_to_next_event_never_skips_a_sample_or_cutoff():
    """Advancing by steps_to_next_event (the device driver's block size) visits every
    count at which a per-step loop would have sampled or stopped."""
    for sp_kwargs, cutoffs, spp in SCHEDULES[:5]:
        sp = sampling.SamplingParams(**sp_kwargs)
        cc = sampling.CompletionCheckParams()
        for k, v in cutoffs.items():
            setattr(cc.cutoff_params, k, v)

        def run(block):
            e = monte.RandomNumberEngine()
            e.seed(3)
            f = rm.SamplingFixture(_fixture_params("s", sp, cc), e)
            state = _state()
            f.initialize(spp)
            f.sample_data_by_count_if_due(state)
            while not f.is_complete():
                if block:
                    steps = f.steps_to_next_event()
                    assert steps % spp == 0
                    f.advance_passes(steps // spp, 0, 0)
                else:
                    f.increment_step()
                f.sample_data_by_count_if_due(state)
            return list(f.results().sample_count), f.counter().count

        assert run(True) == run(False)


def test_run_manager_completion_logic():
    # RunManager.hh:94-112: with global_cutoff any complete fixture ends the run, otherwise all must be
    def manager(global_cutoff):
        e = monte.RandomNumberEngine()
        ps = []
        for label, max_count in (("short", 5), ("long", 12)):
            cc = sampling.CompletionCheckParams()
            cc.cutoff_params.max_count = max_count
            ps.append(_fixture_params(label, sampling.SamplingParams(), cc))
        return rm.RunManager(e, ps, global_cutoff)

    for global_cutoff, expect in ((True, 5), (False, 12)):
        m = manager(global_cutoff)
        state = _state()
        m.initialize(4)
        m.sample_data_by_count_if_due(state)
        while not m.is_complete():
            m.increment_step()
            m.increment_n_reject()
            m.sample_data_by_count_if_due(state)
        assert [f.counter().count for f in m.sampling_fixtures] == [expect, expect]
        assert [f.counter().n_reject for f in m.sampling_fixtures] == [4 * expect, 4 * expect]
        # fixtures keep sampling after they are individually complete (no check in sample_data_by_count_if_due)
        assert [len(f.results().sample_count) for f in m.sampling_fixtures] == [expect, expect]
        m.finalize(state)
        assert all(f.results().n_reject == 4 * expect for f in m.sampling_fixtures)
        assert m.sampling_fixtures[0].results().completion_check_results.is_complete


def test_constructor_errors():
    sp = sampling.SamplingParams(sampler_names=["missing"])
    with pytest.raises(RuntimeError, match="No sampling function for 'missing'"):
        rm.SamplingFixtureParams("l", _constant_functions(), sampling.jsonStateSamplingFunctionMap(),
                                 rm.ResultsAnalysisFunctionMap(), sp, sampling.CompletionCheckParams())
    sp = sampling.SamplingParams(json_sampler_names=["missing"])
    with pytest.raises(RuntimeError, match="No sampling function for 'missing'"):
        rm.SamplingFixtureParams("l", _constant_functions(), sampling.jsonStateSamplingFunctionMap(),
                                 rm.ResultsAnalysisFunctionMap(), sp, sampling.CompletionCheckParams())
    # a schedule that does not advance (SamplingFixture.hh:509-516)
    sp = sampling.SamplingParams(period=0.2)
    cc = sampling.CompletionCheckParams()
    cc.cutoff_params.max_count = 10
    f = rm.SamplingFixture(_fixture_params("stuck", sp, cc), monte.RandomNumberEngine())
    f.initialize(2)
    with pytest.raises(RuntimeError, match="next_sample_count <= current count"):
        f.sample_data(_state())
