// oracle_pybind.cc -- Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY).
// Builds oracle/_monte_oracle*.so ; imported only by tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "kstate_oracle.hh"
#include "monte_oracle.hh"
#include "run_management_oracle.hh"

namespace py = pybind11;
using namespace monte_oracle;

namespace {

typedef py::array_t<int32_t, py::array::c_style | py::array::forcecast> i32arr;
typedef py::array_t<double, py::array::c_style | py::array::forcecast> f64arr;

std::vector<int> to_vec(i32arr const &a) {
  auto r = a.unchecked<1>();
  std::vector<int> v(r.shape(0));
  for (py::ssize_t i = 0; i < r.shape(0); ++i) v[i] = r(i);
  return v;
}
std::vector<double> to_dvec(f64arr const &a) {
  auto r = a.unchecked<1>();
  std::vector<double> v(r.shape(0));
  for (py::ssize_t i = 0; i < r.shape(0); ++i) v[i] = r(i);
  return v;
}
i32arr from_vec(std::vector<int> const &v) {
  i32arr a(v.size());
  auto w = a.mutable_unchecked<1>();
  for (size_t i = 0; i < v.size(); ++i) w(i) = v[i];
  return a;
}
f64arr from_dvec(std::vector<double> const &v) {
  f64arr a(v.size());
  auto w = a.mutable_unchecked<1>();
  for (size_t i = 0; i < v.size(); ++i) w(i) = v[i];
  return a;
}

IsingState make_state(std::vector<int> const &shape, std::vector<int> const &occ,
                      double T, double mu) {
  IsingConfiguration config(shape, 1, /*allow_3d=*/true);
  config.set_occupation(occ);
  ValueMap cond;
  cond.scalar_values["temperature"] = T;
  cond.vector_values["exchange_potential"] = std::vector<double>{mu};
  return IsingState(config, cond);
}

struct Engine {
  std::shared_ptr<std::mt19937_64> e = std::make_shared<std::mt19937_64>();
};

py::dict cc_results_to_dict(CompletionCheckResults const &r) {
  py::dict d;
  d["count"] = r.count.has_value() ? py::cast(r.count.value()) : py::none();
  d["n_samples"] = r.n_samples;
  d["has_all_minimums_met"] = r.has_all_minimums_met;
  d["has_any_maximum_met"] = r.has_any_maximum_met;
  d["n_samples_at_convergence_check"] =
      r.n_samples_at_convergence_check.has_value()
          ? py::cast(r.n_samples_at_convergence_check.value())
          : py::none();
  d["is_complete"] = r.is_complete;
  py::dict eq;
  eq["all_equilibrated"] = r.equilibration_check_results.all_equilibrated;
  eq["N_samples_for_all_to_equilibrate"] =
      r.equilibration_check_results.N_samples_for_all_to_equilibrate;
  py::list eql;
  for (auto const &kv : r.equilibration_check_results.individual_results) {
    py::dict e;
    e["sampler_name"] = kv.first.sampler_name;
    e["component_index"] = kv.first.component_index;
    e["is_equilibrated"] = kv.second.is_equilibrated;
    e["N_samples_for_equilibration"] = kv.second.N_samples_for_equilibration;
    eql.append(e);
  }
  eq["individual_results"] = eql;
  d["equilibration_check_results"] = eq;
  py::dict cv;
  cv["all_converged"] = r.convergence_check_results.all_converged;
  cv["N_samples_for_statistics"] =
      r.convergence_check_results.N_samples_for_statistics;
  py::list cvl;
  for (auto const &kv : r.convergence_check_results.individual_results) {
    py::dict e;
    e["sampler_name"] = kv.first.sampler_name;
    e["component_index"] = kv.first.component_index;
    e["is_converged"] = kv.second.is_converged;
    e["mean"] = kv.second.stats.mean;
    e["calculated_precision"] = kv.second.stats.calculated_precision;
    cvl.append(e);
  }
  cv["individual_results"] = cvl;
  d["convergence_check_results"] = cv;
  return d;
}

// Build CompletionCheckParams from a python dict with the reference's field
// names (include/casm/monte/checks/CompletionCheck.hh:20-61).
CompletionCheckParams params_from_dict(py::dict const &d) {
  CompletionCheckParams p;
  auto opt_count = [&](const char *k, std::optional<CountType> &dst) {
    if (d.contains(k) && !d[k].is_none()) dst = d[k].cast<CountType>();
  };
  auto opt_time = [&](const char *k, std::optional<TimeType> &dst) {
    if (d.contains(k) && !d[k].is_none()) dst = d[k].cast<double>();
  };
  opt_count("min_count", p.cutoff_params.min_count);
  opt_count("max_count", p.cutoff_params.max_count);
  opt_count("min_sample", p.cutoff_params.min_sample);
  opt_count("max_sample", p.cutoff_params.max_sample);
  opt_time("min_time", p.cutoff_params.min_time);
  opt_time("max_time", p.cutoff_params.max_time);
  opt_time("min_clocktime", p.cutoff_params.min_clocktime);
  opt_time("max_clocktime", p.cutoff_params.max_clocktime);
  if (d.contains("log_spacing")) p.log_spacing = d["log_spacing"].cast<bool>();
  if (d.contains("check_begin")) p.check_begin = d["check_begin"].cast<long>();
  if (d.contains("check_period")) p.check_period = d["check_period"].cast<long>();
  if (d.contains("check_base")) p.check_base = d["check_base"].cast<double>();
  if (d.contains("check_shift")) p.check_shift = d["check_shift"].cast<double>();
  if (d.contains("check_period_max"))
    p.check_period_max = d["check_period_max"].cast<long>();
  if (d.contains("confidence"))
    p.calc_statistics_f =
        BasicStatisticsCalculator(d["confidence"].cast<double>());
  if (d.contains("requested_precision")) {
    // list of (sampler_name, component_index, abs or None, rel or None)
    for (auto item : d["requested_precision"].cast<py::list>()) {
      auto t = item.cast<py::tuple>();
      SamplerComponent key(t[0].cast<std::string>(), t[1].cast<long>(),
                           std::to_string(t[1].cast<long>()));
      RequestedPrecision rp;
      if (!t[2].is_none()) {
        rp.abs_convergence_is_required = true;
        rp.abs_precision = t[2].cast<double>();
      }
      if (t.size() > 3 && !t[3].is_none()) {
        rp.rel_convergence_is_required = true;
        rp.rel_precision = t[3].cast<double>();
      }
      p.requested_precision.emplace(key, rp);
    }
  }
  return p;
}

py::dict data_to_dict(BasicOccupationMetropolisData const &data,
                      IsingState const &state) {
  py::dict out;
  out["occupation"] = from_vec(state.configuration.occupation());
  out["n_pass"] = data.n_pass;
  out["n_accept"] = data.n_accept;
  out["n_reject"] = data.n_reject;
  py::dict s;
  for (auto const &kv : data.samplers) s[kv.first.c_str()] =
      from_dvec(kv.second->component(0));
  out["samplers"] = s;
  out["completion_check_results"] =
      cc_results_to_dict(data.completion_check.results());
  out["n_checks"] = data.completion_check.n_checks();
  return out;
}

// Full reference-order SGC run (serial Metropolis, mt19937_64).
py::dict sgc_run(std::vector<int> shape, i32arr occ_in, double J, double T,
                 double mu, bool use_nlist, Engine &engine, py::dict cc_params,
                 int sample_period) {
  IsingState state = make_state(shape, to_vec(occ_in), T, mu);
  auto system = std::make_shared<IsingSystem>(
      IsingFormationEnergy(J, 1, use_nlist), IsingParamComposition());
  auto mc = std::make_shared<SemiGrandCanonicalCalculator>(system);
  StateSamplingFunctionMap fns;
  for (auto const &f : {make_parametric_composition_f(mc),
                        make_formation_energy_f(mc), make_potential_energy_f(mc)})
    fns.emplace(f.name, f);
  CompletionCheckParams p = params_from_dict(cc_params);
  SemiGrandCanonicalCalculator::event_generator_type gen;
  std::optional<MethodLog> log = MethodLog();
  auto no_status = [](BasicOccupationMetropolisData const &, MethodLog &) {};
  {
    py::gil_scoped_release release;
    mc->run(state, fns, p, gen, sample_period, log, engine.e, no_status);
  }
  return data_to_dict(*mc->data, state);
}

py::dict checkerboard_run(std::vector<int> shape, i32arr occ_in, double J,
                          double T, double mu, uint64_t seed, uint32_t chain,
                          uint64_t pass0, long n_passes, long sample_period, int philox_rounds) {
  std::vector<int> occ = to_vec(occ_in);
  const int dim = static_cast<int>(shape.size());
  for (int s : shape)
    if (s % 2) throw std::runtime_error("checkerboard needs even extents");
  AcceptTable tab = make_accept_table(dim, J, T, mu);
  CheckerboardResult res;
  std::vector<long long> Ss, Bs;
  long long N = 1;
  for (int s : shape) N *= s;
  {
    py::gil_scoped_release release;
    for (long t = 0; t < n_passes; ++t) {
      checkerboard_pass(occ, shape, tab, seed, chain, pass0 + t, res, philox_rounds);
      if (sample_period > 0 && ((t + 1) % sample_period) == 0) {
        long long S, B;
        integer_observables(occ, shape, S, B);
        Ss.push_back(S);
        Bs.push_back(B);
      }
    }
  }
  py::dict out;
  out["occupation"] = from_vec(occ);
  out["n_accept"] = res.n_accept;
  out["n_reject"] = n_passes * N - res.n_accept;
  py::array_t<long long> aS(Ss.size()), aB(Bs.size());
  f64arr x(Ss.size()), ef(Ss.size()), ep(Ss.size());
  for (size_t i = 0; i < Ss.size(); ++i) {
    aS.mutable_at(i) = Ss[i];
    aB.mutable_at(i) = Bs[i];
    IntensiveObservables o = observables_from_sums(Ss[i], Bs[i], N, J, mu);
    x.mutable_at(i) = o.param_composition;
    ef.mutable_at(i) = o.formation_energy;
    ep.mutable_at(i) = o.potential_energy;
  }
  out["S"] = aS;
  out["B"] = aB;
  out["param_composition"] = x;
  out["formation_energy"] = ef;
  out["potential_energy"] = ep;
  return out;
}


// ---- run management restatement (run_management_oracle.hh) ----
SamplingParams sampling_params_from_dict(py::dict const &d) {
  SamplingParams s;
  if (d.contains("sampler_names")) s.sampler_names = d["sampler_names"].cast<std::vector<std::string>>();
  if (d.contains("sample_mode")) {
    std::string m = d["sample_mode"].cast<std::string>();
    s.sample_mode = m == "step" ? SAMPLE_MODE::BY_STEP : m == "time" ? SAMPLE_MODE::BY_TIME : SAMPLE_MODE::BY_PASS;
  }
  if (d.contains("sample_method")) {
    std::string m = d["sample_method"].cast<std::string>();
    s.sample_method = m == "log" ? SAMPLE_METHOD::LOG : m == "custom" ? SAMPLE_METHOD::CUSTOM : SAMPLE_METHOD::LINEAR;
  }
  if (d.contains("period")) s.period = d["period"].cast<double>();
  if (d.contains("begin")) s.begin = d["begin"].cast<double>();
  if (d.contains("base")) s.base = d["base"].cast<double>();
  if (d.contains("shift")) s.shift = d["shift"].cast<double>();
  if (d.contains("stochastic_sample_period")) s.stochastic_sample_period = d["stochastic_sample_period"].cast<bool>();
  if (d.contains("do_sample_trajectory")) s.do_sample_trajectory = d["do_sample_trajectory"].cast<bool>();
  return s;
}

py::dict run_management_sgc_run(std::vector<int> shape, i32arr occ_in, double J, double T, double mu,
                                bool use_nlist, Engine &engine, py::list fixtures, bool global_cutoff) {
  IsingState state = make_state(shape, to_vec(occ_in), T, mu);
  auto system = std::make_shared<IsingSystem>(IsingFormationEnergy(J, 1, use_nlist), IsingParamComposition());
  auto mc = std::make_shared<SemiGrandCanonicalCalculator>(system);
  mc->state = &state;
  mc->conditions = std::make_shared<SemiGrandCanonicalConditions>(
      SemiGrandCanonicalConditions::from_values(state.conditions));
  mc->potential.set_state(&state, mc->conditions);
  StateSamplingFunctionMap fns;
  for (auto const &f : {make_parametric_composition_f(mc), make_formation_energy_f(mc), make_potential_energy_f(mc)})
    fns.emplace(f.name, f);
  const double N = static_cast<double>(state.configuration.n_unitcells);
  ResultsAnalysisFunctionMap afs;
  afs.emplace("heat_capacity", ResultsAnalysisFunction{"heat_capacity", "", {}, {"0"}, [=](RunResults const &r) {
                                 return std::vector<double>{N * tail_variance(r, "potential_energy") / (KB * T * T)};
                               }});
  afs.emplace("susceptibility", ResultsAnalysisFunction{"susceptibility", "", {}, {"0"}, [=](RunResults const &r) {
                                  return std::vector<double>{N * tail_variance(r, "param_composition") / (KB * T)};
                                }});
  std::vector<SamplingFixtureParams> params;
  for (auto item : fixtures) {
    py::dict d = item.cast<py::dict>();
    SamplingFixtureParams p;
    p.label = d["label"].cast<std::string>();
    p.sampling_functions = fns;
    p.analysis_functions = afs;
    p.sampling_params = sampling_params_from_dict(d["sampling_params"].cast<py::dict>());
    p.completion_check_params = params_from_dict(d["completion_check_params"].cast<py::dict>());
    if (d.contains("analysis_names")) p.analysis_names = d["analysis_names"].cast<std::vector<std::string>>();
    params.push_back(p);
  }
  RunManager<> run_manager(engine.e, params, global_cutoff);
  SemiGrandCanonicalEventGenerator<> gen;
  gen.set_state(&state);
  RandomNumberGenerator<> rng(engine.e);
  {
    py::gil_scoped_release release;
    ising_occupation_metropolis(state, mc->potential, gen, rng, run_manager);
  }
  py::dict out;
  out["occupation"] = from_vec(state.configuration.occupation());
  out["potential_energy_property"] = state.properties.scalar_values.at("potential_energy");
  py::list fl;
  for (auto const &fp : run_manager.sampling_fixtures) {
    RunResults const &r = fp->results();
    py::dict fd;
    fd["label"] = fp->params().label;
    fd["sample_count"] = r.sample_count;
    py::dict sd;
    for (auto const &kv : r.samplers) {
      auto v = kv.second->component(0);
      f64arr a(v.size());
      std::copy(v.begin(), v.end(), a.mutable_data());
      sd[py::str(kv.first)] = a;
    }
    fd["samplers"] = sd;
    py::list traj;
    for (auto const &occ : r.sample_trajectory) traj.append(from_vec(occ));
    fd["sample_trajectory"] = traj;
    py::dict ad;
    for (auto const &kv : r.analysis) ad[py::str(kv.first)] = kv.second;
    fd["analysis"] = ad;
    fd["n_accept"] = r.n_accept;
    fd["n_reject"] = r.n_reject;
    fd["count"] = fp->counter().count;
    fd["pass"] = fp->counter().pass;
    fd["step"] = fp->counter().step;
    fd["completion_check_results"] = cc_results_to_dict(r.completion_check_results);
    fl.append(fd);
  }
  out["fixtures"] = fl;
  return out;
}

}  // namespace

PYBIND11_MODULE(_monte_oracle, m) {
  m.doc() = "CPU oracle for the Ising SGC Metropolis path (test infrastructure)";
  m.attr("KB") = KB;

  py::class_<Engine>(m, "RandomNumberEngine")
      .def(py::init<>())
      .def("seed", [](Engine &e, uint64_t s) { e.e->seed(s); })
      .def("dump",
           [](Engine const &e) {
             std::stringstream ss;
             ss << *e.e;
             return ss.str();
           })
      .def("load", [](Engine &e, std::string const &s) {
        std::stringstream ss(s);
        ss >> *e.e;
      });

  m.def("random_int", [](Engine &e, uint64_t maximum_value) {
    RandomNumberGenerator<> g(e.e);
    return g.random_int<uint64_t>(maximum_value);
  });
  m.def("random_int_long", [](Engine &e, long maximum_value) {
    RandomNumberGenerator<> g(e.e);
    return g.random_int<long>(maximum_value);
  });
  m.def("random_real", [](Engine &e, double maximum_value) {
    RandomNumberGenerator<> g(e.e);
    return g.random_real<double>(maximum_value);
  });

  m.def("within", [](std::vector<int> shape, long index, int dim) {
    IsingConfiguration c(shape, 1, true);
    return c.within(index, dim);
  });
  m.def("from_linear_site_index", [](std::vector<int> shape, long l) {
    IsingConfiguration c(shape, 1, true);
    return c.from_linear_site_index(l);
  });
  m.def("to_linear_site_index", [](std::vector<int> shape, std::vector<int> mi) {
    IsingConfiguration c(shape, 1, true);
    return c.to_linear_site_index(mi);
  });
  m.def("ising_configuration_2d_only", [](std::vector<int> shape) {
    IsingConfiguration c(shape, 1, false);  // throws unless 2-d
    return c.n_sites;
  });

  m.def("formation_energy",
        [](std::vector<int> shape, i32arr occ, double J, bool use_nlist) {
          IsingState st = make_state(shape, to_vec(occ), 1.0, 0.0);
          IsingFormationEnergy f(J, 1, use_nlist);
          f.set_state(&st);
          return py::make_tuple(f.per_supercell(), f.per_unitcell());
        });
  m.def("formation_energy_delta",
        [](std::vector<int> shape, i32arr occ, double J, bool use_nlist,
           std::vector<long> l, std::vector<int> new_occ) {
          IsingState st = make_state(shape, to_vec(occ), 1.0, 0.0);
          IsingFormationEnergy f(J, 1, use_nlist);
          f.set_state(&st);
          return f.occ_delta_per_supercell(l, new_occ);
        });
  m.def("param_composition", [](std::vector<int> shape, i32arr occ) {
    IsingState st = make_state(shape, to_vec(occ), 1.0, 0.0);
    IsingParamComposition c;
    c.set_state(&st);
    return py::make_tuple(c.per_supercell()[0], c.per_unitcell()[0]);
  });
  m.def("param_composition_delta",
        [](std::vector<int> shape, i32arr occ, std::vector<long> l,
           std::vector<int> new_occ) {
          IsingState st = make_state(shape, to_vec(occ), 1.0, 0.0);
          IsingParamComposition c;
          c.set_state(&st);
          return c.occ_delta_per_supercell(l, new_occ)[0];
        });
  m.def("potential", [](std::vector<int> shape, i32arr occ, double J, double T,
                        double mu, bool use_nlist) {
    IsingState st = make_state(shape, to_vec(occ), T, mu);
    auto sys = std::make_shared<IsingSystem>(
        IsingFormationEnergy(J, 1, use_nlist), IsingParamComposition());
    SemiGrandCanonicalPotential pot(sys);
    pot.set_state(&st, std::make_shared<SemiGrandCanonicalConditions>(
                           SemiGrandCanonicalConditions::from_values(
                               st.conditions)));
    return py::make_tuple(pot.per_supercell(), pot.per_unitcell());
  });
  m.def("potential_delta",
        [](std::vector<int> shape, i32arr occ, double J, double T, double mu,
           bool use_nlist, std::vector<long> l, std::vector<int> new_occ) {
          IsingState st = make_state(shape, to_vec(occ), T, mu);
          auto sys = std::make_shared<IsingSystem>(
              IsingFormationEnergy(J, 1, use_nlist), IsingParamComposition());
          SemiGrandCanonicalPotential pot(sys);
          pot.set_state(&st, std::make_shared<SemiGrandCanonicalConditions>(
                                 SemiGrandCanonicalConditions::from_values(
                                     st.conditions)));
          return pot.occ_delta_per_supercell(l, new_occ);
        });
  // dE of the single-site flip at every site, through the reference call chain
  m.def("potential_delta_all_sites",
        [](std::vector<int> shape, i32arr occ, double J, double T, double mu,
           bool use_nlist) {
          IsingState st = make_state(shape, to_vec(occ), T, mu);
          auto sys = std::make_shared<IsingSystem>(
              IsingFormationEnergy(J, 1, use_nlist), IsingParamComposition());
          SemiGrandCanonicalPotential pot(sys);
          pot.set_state(&st, std::make_shared<SemiGrandCanonicalConditions>(
                                 SemiGrandCanonicalConditions::from_values(
                                     st.conditions)));
          long N = st.configuration.n_sites;
          f64arr out(N);
          auto w = out.mutable_unchecked<1>();
          std::vector<long> l(1);
          std::vector<int> no(1);
          for (long i = 0; i < N; ++i) {
            l[0] = i;
            no[0] = -st.configuration.occ(i);
            w(i) = pot.occ_delta_per_supercell(l, no);
          }
          return out;
        });
  // acceptance decision at every site given one uniform per site
  // (include/casm/monte/methods/metropolis.hh:26-35 with the draw supplied)
  m.def("accept_all_sites", [](f64arr dE, f64arr u, double T) {
    auto d = dE.unchecked<1>();
    auto uu = u.unchecked<1>();
    double beta = 1.0 / (KB * T);
    py::array_t<uint8_t> out(d.shape(0));
    auto w = out.mutable_unchecked<1>();
    for (py::ssize_t i = 0; i < d.shape(0); ++i) {
      bool acc = d(i) < 0.0;
      if (!acc) acc = uu(i) < std::exp(-d(i) * beta);
      w(i) = acc ? 1 : 0;
    }
    return out;
  });

  m.def("accept_table", [](int dim, double J, double T, double mu) {
    AcceptTable t = make_accept_table(dim, J, T, mu);
    const int z = 2 * dim;
    f64arr dE({2, z + 1}), prob({2, z + 1});
    py::array_t<uint32_t> thr({2, z + 1});
    py::array_t<bool> never({2, z + 1});
    for (int s = 0; s < 2; ++s)
      for (int n = 0; n <= z; ++n) {
        dE.mutable_at(s, n) = t.dE[s][n];
        prob.mutable_at(s, n) = t.prob[s][n];
        thr.mutable_at(s, n) = t.thr_m1[s][n];
        never.mutable_at(s, n) = t.never[s][n];
      }
    py::dict d;
    d["dE"] = dE;
    d["prob"] = prob;
    d["thr_m1"] = thr;
    d["never"] = never;
    d["beta"] = t.beta;
    return d;
  });

  m.def("integer_observables", [](std::vector<int> shape, i32arr occ) {
    long long S, B;
    integer_observables(to_vec(occ), shape, S, B);
    return py::make_tuple(S, B);
  });
  m.def("observables_from_sums",
        [](long long S, long long B, long long N, double J, double mu) {
          IntensiveObservables o = observables_from_sums(S, B, N, J, mu);
          return py::make_tuple(o.param_composition, o.formation_energy,
                                o.potential_energy);
        });
  m.def("heat_capacity", [](f64arr e, long long N, double T) {
    return heat_capacity(to_dvec(e), N, T);
  });
  m.def("susceptibility", [](f64arr x, long long N, double T) {
    return susceptibility(to_dvec(x), N, T);
  });

  m.def("run_management_sgc_run", &run_management_sgc_run, py::arg("shape"), py::arg("occupation"), py::arg("J"),
        py::arg("temperature"), py::arg("mu"), py::arg("use_nlist"), py::arg("engine"), py::arg("fixtures"),
        py::arg("global_cutoff") = true);
  m.def("rm_sample_at", [](long i, py::dict d) { return sample_at(i, sampling_params_from_dict(d)); });
  m.def("rm_stochastic_count_steps", [](Engine &engine, double rate, int n) {
    RandomNumberGenerator<> rng(engine.e);
    std::vector<long> out;
    for (int i = 0; i < n; ++i) out.push_back(stochastic_count_step(rate, rng));
    return out;
  });
  m.def("rm_monte_counter_trace", [](std::string mode, long steps_per_pass, long n_steps) {
    MonteCounter c;
    c.reset(mode == "step" ? SAMPLE_MODE::BY_STEP : SAMPLE_MODE::BY_PASS, steps_per_pass);
    std::vector<std::array<long, 3>> out;
    for (long i = 0; i < n_steps; ++i) {
      c.increment_step();
      out.push_back({c.step, c.pass, c.count});
    }
    return out;
  });
  m.def("rm_fixture_schedule", [](py::dict sampling_params, py::dict cc_params, long steps_per_pass, long max_steps,
                                  Engine &engine) {
    // drive one fixture step by step with constant sampling functions; returns the counts at which it sampled
    SamplingFixtureParams p;
    p.label = "schedule";
    p.sampling_functions.emplace("x", StateSamplingFunction("x", "", {}, []() { return std::vector<double>{1.0}; }));
    p.sampling_params = sampling_params_from_dict(sampling_params);
    p.sampling_params.sampler_names = {"x"};
    p.completion_check_params = params_from_dict(cc_params);
    SamplingFixture<> f(p, engine.e);
    IsingState state = make_state({2, 2}, std::vector<int>(4, 1), 1000.0, 0.0);
    f.initialize(steps_per_pass);
    f.sample_data_by_count_if_due(state);
    long n = 0;
    while (!f.is_complete() && n < max_steps) {
      f.increment_step();
      f.sample_data_by_count_if_due(state);
      ++n;
    }
    py::dict out;
    out["sample_count"] = f.results().sample_count;
    out["steps"] = n;
    out["count"] = f.counter().count;
    out["is_complete"] = f.is_complete();
    return out;
  });

  m.def("sgc_run", &sgc_run, py::arg("shape"), py::arg("occupation"),
        py::arg("J"), py::arg("temperature"), py::arg("mu"),
        py::arg("use_nlist"), py::arg("engine"),
        py::arg("completion_check_params"), py::arg("sample_period") = 1);
  m.def("checkerboard_run", &checkerboard_run, py::arg("shape"),
        py::arg("occupation"), py::arg("J"), py::arg("temperature"),
        py::arg("mu"), py::arg("seed"), py::arg("chain") = 0,
        py::arg("pass0") = 0, py::arg("n_passes") = 1,
        py::arg("sample_period") = 1, py::arg("philox_rounds") = 10);

  m.def("checkerboard_half_sweep_slab",
        [](i32arr occ, i32arr halo_lo, i32arr halo_hi, long n0, long col_begin, long n_cols, double J,
           double T, double mu, uint64_t seed, uint32_t chain, uint64_t pass_index, int colour) {
          std::vector<int> o = to_vec(occ);
          AcceptTable tab = make_accept_table(2, J, T, mu);
          long long acc = checkerboard_half_sweep_slab(o, to_vec(halo_lo), to_vec(halo_hi), n0, col_begin,
                                                       n_cols, tab, seed, chain, pass_index, colour);
          return py::make_tuple(from_vec(o), acc);
        });
  m.def("philox4x32_10", [](std::array<uint32_t, 4> c, std::array<uint32_t, 2> k) {
    return Philox4x32::generate(c, k, 10);
  });

  // statistics
  m.def("variance", [](f64arr x, double mean) {
    auto v = to_dvec(x);
    return variance(v.data(), v.size(), mean);
  });
  m.def("covariance_lag", [](f64arr x, long k, double mean) {
    auto v = to_dvec(x);
    long n = static_cast<long>(v.size()) - k;
    return covariance(v.data(), v.data() + k, n, mean);
  });
  m.def("approx_erf_inv", &approx_erf_inv);
  m.def("autocorrelation_factor", [](f64arr x, double increment) {
    auto v = to_dvec(x);
    Index k = 0;
    double f = autocorrelation_factor(v.data(), v.size(), increment, &k);
    return py::make_tuple(f, k);
  }, py::arg("observations"), py::arg("increment") = 1.0);
  m.def("resample", [](f64arr x, f64arr w, double weight_sum, long n) {
    return resample(to_dvec(x), to_dvec(w), weight_sum, n);
  }, py::arg("observations"), py::arg("sample_weight"), py::arg("sample_weight_sum"),
     py::arg("n_equally_spaced"));
  m.def("basic_statistics",
        [](f64arr x, f64arr w, double confidence, long method, long n_resamples) {
          BasicStatisticsCalculator c(confidence, method, n_resamples);
          BasicStatistics s = c(to_dvec(x), to_dvec(w));
          return py::make_tuple(s.mean, s.calculated_precision);
        },
        py::arg("observations"), py::arg("sample_weight") = f64arr(0),
        py::arg("confidence") = 0.95, py::arg("method") = 1,
        py::arg("n_resamples") = 10000);
  m.def("default_equilibration_check",
        [](f64arr x, f64arr w, py::object abs, py::object rel) {
          RequestedPrecision rp;
          if (!abs.is_none()) {
            rp.abs_convergence_is_required = true;
            rp.abs_precision = abs.cast<double>();
          }
          if (!rel.is_none()) {
            rp.rel_convergence_is_required = true;
            rp.rel_precision = rel.cast<double>();
          }
          auto r = default_equilibration_check(to_dvec(x), to_dvec(w), rp);
          return py::make_tuple(r.is_equilibrated,
                                r.N_samples_for_equilibration);
        },
        py::arg("observations"), py::arg("sample_weight") = f64arr(0),
        py::arg("abs") = py::none(), py::arg("rel") = py::none());

  // Sampler + CompletionCheck objects, for tests that mirror
  // python/tests/sampling/test_Sampler.py and test_CompletionCheck.py
  py::class_<Sampler, std::shared_ptr<Sampler>>(m, "Sampler")
      .def(py::init([](std::vector<Index> shape,
                       std::optional<std::vector<std::string>> names,
                       CountType inc) {
             if (names.has_value())
               return std::make_shared<Sampler>(shape, names.value(), inc);
             return std::make_shared<Sampler>(shape, inc);
           }),
           py::arg("shape"), py::arg("component_names") = py::none(),
           py::arg("capacity_increment") = 1000)
      .def("append",
           [](Sampler &s, std::vector<double> v) { s.push_back(v); })
      .def("clear", &Sampler::clear)
      .def("set_sample_capacity", &Sampler::set_sample_capacity)
      .def("set_capacity_increment", &Sampler::set_capacity_increment)
      .def("component_names", &Sampler::component_names)
      .def("shape", &Sampler::shape)
      .def("n_components", &Sampler::n_components)
      .def("n_samples", &Sampler::n_samples)
      .def("sample_capacity", &Sampler::sample_capacity)
      .def("component",
           [](Sampler const &s, Index c) { return from_dvec(s.component(c)); })
      .def("sample",
           [](Sampler const &s, CountType r) { return from_dvec(s.sample(r)); })
      .def("values", [](Sampler const &s) {
        f64arr a({static_cast<py::ssize_t>(s.n_samples()),
                  static_cast<py::ssize_t>(s.n_components())});
        for (Index c = 0; c < s.n_components(); ++c)
          for (CountType r = 0; r < s.n_samples(); ++r)
            a.mutable_at(r, c) = s.component_data(c)[r];
        return a;
      });
  m.def("default_component_names", &default_component_names);
  m.def("colmajor_component_names", &colmajor_component_names);

  struct CC {
    CompletionCheck cc;
    Clock clock;
    explicit CC(py::dict d) : cc(params_from_dict(d)) {}
  };
  py::class_<CC>(m, "CompletionCheck")
      .def(py::init<py::dict>())
      .def("reset", [](CC &c) { c.cc.reset(); })
      .def("count_check",
           [](CC &c, std::map<std::string, std::shared_ptr<Sampler>> samplers,
              std::shared_ptr<Sampler> weight, CountType count) {
             return c.cc.is_complete(samplers, *weight, count, c.clock);
           },
           py::arg("samplers"), py::arg("sample_weight"), py::arg("count"))
      .def("check",
           [](CC &c, std::map<std::string, std::shared_ptr<Sampler>> samplers,
              std::shared_ptr<Sampler> weight) {
             return c.cc.is_complete(samplers, *weight, c.clock);
           })
      .def("n_checks", [](CC const &c) { return c.cc.n_checks(); })
      .def("results", [](CC const &c) { return cc_results_to_dict(c.cc.results()); });

  // Conversions (diagonal transformation matrix)
  m.def("conv_l_size", [](std::array<long, 3> n, long nb) {
    DiagonalConversions c{{n[0], n[1], n[2]}, nb};
    return c.l_size();
  });
  m.def("conv_l_to_bijk", [](std::array<long, 3> n, long nb, long l) {
    DiagonalConversions c{{n[0], n[1], n[2]}, nb};
    long ijk[3];
    c.l_to_ijk(l, ijk);
    return py::make_tuple(c.l_to_b(l), ijk[0], ijk[1], ijk[2]);
  });
  m.def("conv_bijk_to_l",
        [](std::array<long, 3> n, long nb, long b, long i, long j, long k) {
          DiagonalConversions c{{n[0], n[1], n[2]}, nb};
          return c.bijk_to_l(b, i, j, k);
        });

  // ---- k-state model driven by the general proposal machinery (kstate_oracle.hh) ----
  auto kmodel = [](int dim, int K, f64arr V) {
    if (K < 2 || K > kstate::kMaxSpecies) throw std::runtime_error("K must be 2..4");
    kstate::KStateModel m;
    m.dim = dim;
    m.K = K;
    auto v = V.unchecked<2>();
    for (int a = 0; a < K; ++a)
      for (int b = 0; b < K; ++b) m.V[a][b] = v(a, b);
    return m;
  };
  auto kresult = [](kstate::KStateRunResult const &r, int K) {
    py::dict d;
    d["occupation"] = from_vec(r.occupation);
    d["n_accept"] = r.n_accept;
    d["n_reject"] = r.n_reject;
    const py::ssize_t n = static_cast<py::ssize_t>(r.samples.size());
    py::array_t<int64_t> counts({n, static_cast<py::ssize_t>(K)});
    py::array_t<int64_t> bonds({n, static_cast<py::ssize_t>(K), static_cast<py::ssize_t>(K)});
    for (py::ssize_t i = 0; i < n; ++i)
      for (int a = 0; a < K; ++a) {
        counts.mutable_at(i, a) = r.samples[i].count[a];
        for (int b = 0; b < K; ++b) bonds.mutable_at(i, a, b) = r.samples[i].bonds[a][b];
      }
    d["counts"] = counts;
    d["bonds"] = bonds;
    return d;
  };
  m.def("kstate_serial_run",
        [kmodel, kresult](std::vector<int> shape, i32arr occ, int K, f64arr V, double T, f64arr mu,
                          Engine &engine, long n_passes, long sample_period) {
          kstate::KStateModel model = kmodel(static_cast<int>(shape.size()), K, V);
          std::vector<double> muv = to_dvec(mu);
          muv.resize(kstate::kMaxSpecies, 0.0);
          return kresult(kstate::kstate_serial_run(shape, to_vec(occ), model, T, muv.data(), engine.e, n_passes, sample_period), K);
        },
        py::arg("shape"), py::arg("occupation"), py::arg("K"), py::arg("V"), py::arg("temperature"), py::arg("mu"),
        py::arg("engine"), py::arg("n_passes"), py::arg("sample_period") = 1);
  m.def("kstate_checkerboard_run",
        [kmodel, kresult](std::vector<int> shape, i32arr occ, int K, f64arr V, double T, f64arr mu, uint64_t seed,
                          uint32_t chain, uint64_t pass0, long n_passes, long sample_period) {
          kstate::KStateModel model = kmodel(static_cast<int>(shape.size()), K, V);
          std::vector<double> muv = to_dvec(mu);
          muv.resize(kstate::kMaxSpecies, 0.0);
          return kresult(kstate::kstate_checkerboard_run(shape, to_vec(occ), model, T, muv.data(), seed, chain, pass0, n_passes, sample_period), K);
        },
        py::arg("shape"), py::arg("occupation"), py::arg("K"), py::arg("V"), py::arg("temperature"), py::arg("mu"),
        py::arg("seed"), py::arg("chain") = 0, py::arg("pass0") = 0, py::arg("n_passes") = 1, py::arg("sample_period") = 1);
  m.def("kstate_table", [kmodel](int dim, int K, f64arr V, double T, f64arr mu) {
    kstate::KStateModel model = kmodel(dim, K, V);
    std::vector<double> muv = to_dvec(mu);
    muv.resize(kstate::kMaxSpecies, 0.0);
    kstate::KStateTable t = kstate::make_kstate_table(model, T, muv.data());
    py::dict d;
    const py::ssize_t n = static_cast<py::ssize_t>(t.dPhi.size());
    f64arr dPhi(n), prob(n);
    py::array_t<uint32_t> thr(n);
    py::array_t<uint8_t> never(n);
    for (py::ssize_t i = 0; i < n; ++i) {
      dPhi.mutable_at(i) = t.dPhi[i];
      prob.mutable_at(i) = t.prob[i];
      thr.mutable_at(i) = t.thr_m1[i];
      never.mutable_at(i) = t.never[i];
    }
    d["n_cfg"] = t.n_cfg;
    d["dPhi"] = dPhi;
    d["prob"] = prob;
    d["thr_m1"] = thr;
    d["never"] = never;
    return d;
  });
  m.def("kstate_potential", [kmodel](int dim, int K, f64arr V, f64arr mu, py::array_t<int64_t> counts, py::array_t<int64_t> bonds) {
    kstate::KStateModel model = kmodel(dim, K, V);
    std::vector<double> muv = to_dvec(mu);
    muv.resize(kstate::kMaxSpecies, 0.0);
    kstate::KStateSample s;
    for (int a = 0; a < K; ++a) {
      s.count[a] = counts.at(a);
      for (int b = 0; b < K; ++b) s.bonds[a][b] = bonds.at(a, b);
    }
    return kstate::kstate_potential(model, muv.data(), s);
  });
  // propose + apply n semi-grand canonical events (every event accepted) with the restated
  // OccLocation: [(linear_site_index, new_occ)] and the final occupation
  m.def("kstate_propose_sequence", [](int K, i32arr occ_in, Engine &engine, long n) {
    std::vector<int> occ = to_vec(occ_in);
    kstate::SimpleConversions convert{static_cast<Index>(occ.size()), K};
    kstate::OccCandidateList list(convert);
    std::vector<kstate::OccSwap> swaps = kstate::make_semigrand_canonical_swaps(convert, list);
    kstate::OccLocation loc(convert, list);
    loc.initialize(occ);
    RandomNumberGenerator<std::mt19937_64> rng(engine.e);
    kstate::KOccEvent e;
    std::vector<std::pair<long, int>> out;
    for (long i = 0; i < n; ++i) {
      kstate::propose_semigrand_canonical_event(e, loc, swaps, rng);
      out.emplace_back(e.linear_site_index[0], e.new_occ[0]);
      loc.apply(e, occ);
    }
    return py::make_tuple(out, from_vec(occ));
  });
  // N-fold way driver (nfold.hh:80-147) with the BKL class selector
  m.def("nfold_run", [](std::vector<int> shape, i32arr occ, double J, double T, double mu, Engine &engine,
                        long long n_steps, long long sample_period, bool keep_events) {
    nfold::NfoldResult r = nfold::nfold_run(shape, to_vec(occ), J, T, mu, engine.e, n_steps, sample_period, keep_events);
    py::dict d;
    d["occupation"] = from_vec(r.occupation);
    d["S"] = py::array_t<long long>(r.S.size(), r.S.data());
    d["B"] = py::array_t<long long>(r.B.size(), r.B.data());
    d["weight"] = from_dvec(r.weight);
    d["expected_acceptance_rate"] = from_dvec(r.expected_acceptance_rate);
    d["event_site"] = py::array_t<long>(r.event_site.size(), r.event_site.data());
    d["time"] = r.time;
    return d;
  }, py::arg("shape"), py::arg("occupation"), py::arg("J"), py::arg("temperature"), py::arg("mu"), py::arg("engine"),
     py::arg("n_steps"), py::arg("sample_period") = 1, py::arg("keep_events") = false);
  // the restated proposal machinery on its own (for the host-mirror tests)
  m.def("kstate_swaps", [](int K) {
    kstate::SimpleConversions convert{1, K};
    kstate::OccCandidateList list(convert);
    std::vector<std::array<long, 4>> out;
    for (auto const &sw : kstate::make_semigrand_canonical_swaps(convert, list))
      out.push_back({sw.cand_a.asym, sw.cand_a.species_index, sw.cand_b.asym, sw.cand_b.species_index});
    return out;
  });
}
