"""Throughput of the drop-in Python API (SemiGrandCanonicalCalculator.run) on the
headline lattice: the call a libcasm.monte user makes, completion checks and all.
usage (GPU box): python tools/api_run_bench.py [n0 n1 max_count check_period]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import casmcode_monte_b200.monte as monte
import casmcode_monte_b200.monte.ising_cpp as ising
import casmcode_monte_b200.monte.ising_cpp.semigrand_canonical as sgc
import casmcode_monte_b200.monte.sampling as sampling

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_count = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
check_period = int(sys.argv[4]) if len(sys.argv) > 4 else 100
J = 0.1
for T in (2800.0, 2200.0, 2800.0):
    mc = sgc.SemiGrandCanonicalCalculator(
        system=ising.IsingSystem(
            formation_energy_calculator=ising.IsingFormationEnergy(J=J, lattice_type=1, use_nlist=(os.environ.get("USE_NLIST", "1") == "1")),
            param_composition_calculator=ising.IsingParamComposition(),
        )
    )
    mc.overlap_checks = os.environ.get("OVERLAP", "0") == "1"
    state = ising.IsingState(
        configuration=ising.IsingConfiguration(shape=(n0, n1)),
        conditions=monte.ValueMap.from_dict({"temperature": T, "exchange_potential": [0.0]}),
    )
    fns = mc.default_sampling_functions()
    p = sampling.CompletionCheckParams()
    p.cutoff_params.min_sample = 100
    p.cutoff_params.max_count = max_count
    p.log_spacing = False
    p.check_begin = 100
    p.check_period = check_period
    sampling.converge(fns, p).set_precision("potential_energy", abs=1e-7).set_precision("param_composition", abs=1e-7)
    e = monte.RandomNumberEngine()
    e.seed(5)
    t0 = time.perf_counter()
    mc.run(
        state=state,
        sampling_functions=fns,
        json_sampling_functions=sampling.jsonStateSamplingFunctionMap(),
        completion_check_params=p,
        event_generator=sgc.SemiGrandCanonicalEventGenerator(),
        sample_period=1,
        random_engine=e,
    )
    dt = time.perf_counter() - t0
    d = mc.data
    r = d.completion_check.results()
    print(json.dumps({
        "lattice": [n0, n1], "T": T, "n_pass": d.n_pass, "n_samples": sampling.get_n_samples(d.samplers),
        "seconds": dt, "attempts_per_s": n0 * n1 * d.n_pass / dt, "is_complete": r.is_complete,
        "n_checks": r.n_samples_at_convergence_check if hasattr(r, "n_samples_at_convergence_check") else None,
        "potential_energy_mean": float(d.samplers["potential_energy"].component(0).mean()),
    }), flush=True)
