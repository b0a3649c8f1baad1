// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's run-management
// path for the Ising semi-grand canonical model.  Only
// tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
//
// Restates, in plain scalar C++ on top of monte_oracle.hh:
//   sampling/SamplingParams.hh:163-273      SamplingParams, sample_at,
//                                           stochastic_count_step / _time_step / _sample_at
//   run_management/SamplingFixture.hh:79-119   MonteCounter
//   run_management/SamplingFixture.hh:121-644  SamplingFixture
//   run_management/RunManager.hh:20-236        RunManager (count-based parts)
//   run_management/Results.hh:14-120           Results
//   run_management/ResultsAnalysisFunction.hh  ResultsAnalysisFunction, make_analysis
//   methods/occupation_metropolis.hh:90-154    the main loop
//
// PARITY: unpinned.  The reference holds no golden vector or known-answer test
// for this path (no test constructs a RunManager; occupation_metropolis needs
// OccLocation / OccCandidate machinery from the multi-species path).  The loop
// is restated with the Ising event generator (basic_semigrand_canonical.hh:268-321)
// in place of propose_event_f(event, occ_location, possible_swaps, rng) and
// occ_location.apply: same call order -- propose, dpotential, acceptance, apply,
// increment_step, sample_data_by_count_if_due.
#ifndef MONTE_ORACLE_RUN_MANAGEMENT_HH
#define MONTE_ORACLE_RUN_MANAGEMENT_HH

#include <cmath>
#include <limits>

#include "monte_oracle.hh"

namespace monte_oracle {

enum class SAMPLE_MODE { BY_STEP, BY_PASS, BY_TIME };    // definitions.hh:20
enum class SAMPLE_METHOD { LINEAR, LOG, CUSTOM };       // definitions.hh:23

// sampling/SamplingParams.hh:163-214 (defaults :216-227)
struct SamplingParams {
  std::vector<std::string> sampler_names;
  std::vector<std::string> json_sampler_names;
  SAMPLE_MODE sample_mode = SAMPLE_MODE::BY_PASS;
  SAMPLE_METHOD sample_method = SAMPLE_METHOD::LINEAR;
  double period = 1.0;
  double begin = 1.0;
  double base = std::pow(10.0, 1.0 / 10.0);
  double shift = 10.0;
  std::function<double(CountType)> custom_sample_at;
  bool stochastic_sample_period = false;
  bool do_sample_trajectory = false;
  bool do_sample_time = false;
};

// SamplingParams.hh:229-245
inline double sample_at(CountType sample_index, SamplingParams const &s) {
  double n = static_cast<double>(sample_index);
  if (s.sample_method == SAMPLE_METHOD::LINEAR) return s.begin + s.period * n;
  if (s.sample_method == SAMPLE_METHOD::LOG) return s.begin + std::pow(s.base, (n + s.shift));
  if (!s.custom_sample_at)
    throw std::runtime_error(
        "Error in sample_at: sample_method==SAMPLE_METHOD::CUSTOM and !custom_sample_at");
  return s.custom_sample_at(sample_index);
}

// SamplingParams.hh:247-259: geometric waiting time, one draw per trial
template <typename EngineType>
CountType stochastic_count_step(double sample_rate, RandomNumberGenerator<EngineType> &rng) {
  CountType dn = 1;
  while (true) {
    if (rng.random_real(1.0) < sample_rate) return dn;
    ++dn;
  }
}
// SamplingParams.hh:261-267
template <typename EngineType>
TimeType stochastic_time_step(TimeType sample_rate, RandomNumberGenerator<EngineType> &rng) {
  return -std::log(rng.random_real(1.0)) / sample_rate;
}
// SamplingParams.hh:270-305.  (The reference's CUSTOM branch is unreachable: it
// repeats the LOG test, so `rate` is uninitialised for CUSTOM; restated as an error.)
template <typename EngineType>
double stochastic_sample_at(CountType sample_index, SamplingParams const &s,
                            RandomNumberGenerator<EngineType> &rng,
                            std::vector<CountType> const &sample_count,
                            std::vector<TimeType> const &sample_time) {
  if (sample_index == 0) return s.begin;
  double n = static_cast<double>(sample_index);
  double rate;
  if (s.sample_method == SAMPLE_METHOD::LINEAR) {
    rate = 1.0 / s.period;
  } else if (s.sample_method == SAMPLE_METHOD::LOG) {
    rate = 1.0 / (std::log(s.base) * std::pow(s.base, (n + s.shift)));
  } else {
    throw std::runtime_error("stochastic_sample_at: CUSTOM is not defined by the reference");
  }
  if (s.sample_mode == SAMPLE_MODE::BY_TIME)
    return sample_time.back() + stochastic_time_step(rate, rng);
  return static_cast<double>(sample_count.back() + stochastic_count_step(rate, rng));
}

// SamplingFixture.hh:79-119
struct MonteCounter {
  MonteCounter() { reset(SAMPLE_MODE::BY_PASS, 1); }
  SAMPLE_MODE sample_mode;
  CountType steps_per_pass, step, pass, count;
  TimeType time;
  long long n_accept, n_reject;
  void reset(SAMPLE_MODE m, CountType spp) {
    sample_mode = m;
    steps_per_pass = spp;
    step = pass = count = 0;
    time = 0.0;
    n_accept = n_reject = 0;
  }
  void increment_step() {
    ++step;
    if (sample_mode == SAMPLE_MODE::BY_STEP) ++count;
    if (step == steps_per_pass) {
      ++pass;
      if (sample_mode != SAMPLE_MODE::BY_STEP) ++count;
      step = 0;
    }
  }
};

struct RunResults;
// ResultsAnalysisFunction.hh:21-63
struct ResultsAnalysisFunction {
  std::string name, description;
  std::vector<Index> shape;
  std::vector<std::string> component_names;
  std::function<std::vector<double>(RunResults const &)> function;
};
typedef std::map<std::string, ResultsAnalysisFunction> ResultsAnalysisFunctionMap;

// Results.hh:14-120 (scalar samplers only; JSON samplers are not restated)
struct RunResults {
  std::vector<std::string> sampler_names;
  StateSamplingFunctionMap sampling_functions;
  ResultsAnalysisFunctionMap analysis_functions;
  std::optional<TimeType> elapsed_clocktime;
  SamplerMap samplers;
  std::map<std::string, std::vector<double>> analysis;
  std::vector<CountType> sample_count;
  std::vector<TimeType> sample_time;
  Sampler sample_weight = Sampler(std::vector<Index>{});
  std::vector<TimeType> sample_clocktime;
  std::vector<std::vector<int>> sample_trajectory;
  CompletionCheckResults completion_check_results;
  long long n_accept = 0, n_reject = 0;
  void reset() {
    elapsed_clocktime.reset();
    samplers.clear();
    analysis.clear();
    sample_count.clear();
    sample_time.clear();
    sample_weight.clear();
    sample_clocktime.clear();
    sample_trajectory.clear();
    completion_check_results.full_reset();
    n_accept = n_reject = 0;
    for (auto const &name : sampler_names) {
      auto it = sampling_functions.find(name);
      if (it == sampling_functions.end())
        throw std::runtime_error("Results::reset error. Failed to find sampling function '" + name + "'.");
      auto const &f = it->second;
      samplers.emplace(f.name, std::make_shared<Sampler>(f.shape, f.component_names));
    }
  }
};

// ResultsAnalysisFunction.hh:107-136: unknown names are skipped, a throwing
// function yields NaNs
inline std::map<std::string, std::vector<double>> make_analysis(
    RunResults const &results, ResultsAnalysisFunctionMap const &fs, std::vector<std::string> const &names) {
  std::map<std::string, std::vector<double>> analysis;
  for (auto const &name : names) {
    auto it = fs.find(name);
    if (it == fs.end()) continue;
    auto const &f = it->second;
    try {
      analysis.emplace(f.name, f.function(results));
    } catch (std::exception &) {
      analysis.emplace(f.name, std::vector<double>(f.component_names.size(),
                                                   std::numeric_limits<double>::quiet_NaN()));
    }
  }
  return analysis;
}

// SamplingFixture.hh:24-77
struct SamplingFixtureParams {
  std::string label;
  StateSamplingFunctionMap sampling_functions;
  ResultsAnalysisFunctionMap analysis_functions;
  SamplingParams sampling_params;
  CompletionCheckParams completion_check_params;
  std::vector<std::string> analysis_names;
};

// SamplingFixture.hh:121-644 (count-based sampling; no status files)
template <typename EngineType = default_engine_type>
class SamplingFixture {
 public:
  SamplingFixture(SamplingFixtureParams const &p, std::shared_ptr<EngineType> engine)
      : m_params(p), m_rng(engine), m_completion_check(p.completion_check_params) {
    for (auto const &name : p.sampling_params.sampler_names)
      if (!p.sampling_functions.count(name))
        throw std::runtime_error("SamplingFixtureParams constructor error: No sampling function for '" + name + "'");
    m_results.sampler_names = p.sampling_params.sampler_names;
    m_results.sampling_functions = p.sampling_functions;
    m_results.analysis_functions = p.analysis_functions;
  }
  SamplingFixtureParams const &params() const { return m_params; }
  MonteCounter const &counter() const { return m_counter; }
  RunResults const &results() const { return m_results; }
  CompletionCheck const &completion_check() const { return m_completion_check; }
  CountType next_sample_count() const { return m_next_sample_count; }

  void initialize(Index steps_per_pass) {
    m_is_complete = false;
    m_counter.reset(m_params.sampling_params.sample_mode, steps_per_pass);
    m_completion_check.reset();
    m_results.reset();
    if (m_params.sampling_params.sample_mode == SAMPLE_MODE::BY_TIME)
      throw std::runtime_error("oracle: BY_TIME sampling is not restated (no simulated time in Metropolis)");
    m_next_sample_count = static_cast<CountType>(std::round(this->sample_at(m_results.sample_count.size())));
    if (m_next_sample_count < 0)
      throw std::runtime_error("Error: sampling period parameter error, next_sample_count < 0");
    m_clock.restart_clock();
  }
  bool is_complete() {
    if (m_is_complete) return true;
    m_is_complete = m_completion_check.is_complete(m_results.samplers, m_results.sample_weight,
                                                   m_counter.count, m_clock);
    return m_is_complete;
  }
  void increment_n_accept() { ++m_counter.n_accept; }
  void increment_n_reject() { ++m_counter.n_reject; }
  void increment_step() { m_counter.increment_step(); }

  void sample_data(IsingState const &state) {
    m_results.sample_count.push_back(m_counter.count);
    m_results.sample_clocktime.push_back(m_clock.time_s());
    if (m_params.sampling_params.do_sample_trajectory)
      m_results.sample_trajectory.push_back(state.configuration.occupation());
    for (auto const &name : m_params.sampling_params.sampler_names)
      m_results.samplers.at(name)->push_back(m_params.sampling_functions.at(name)());
    m_next_sample_count = static_cast<CountType>(std::round(this->sample_at(m_results.sample_count.size())));
    if (m_next_sample_count <= m_counter.count)
      throw std::runtime_error(
          "Error: state sampling period parameter error, next_sample_count <= current count");
  }
  void sample_data_by_count_if_due(IsingState const &state) {
    if (m_counter.count == m_next_sample_count) sample_data(state);
  }
  double sample_at(CountType sample_index) {
    if (m_params.sampling_params.stochastic_sample_period)
      return stochastic_sample_at(sample_index, m_params.sampling_params, m_rng, m_results.sample_count,
                                  m_results.sample_time);
    return monte_oracle::sample_at(sample_index, m_params.sampling_params);
  }
  void finalize(IsingState const &) {
    m_results.elapsed_clocktime = m_clock.time_s();
    m_results.completion_check_results = m_completion_check.results();
    m_results.analysis = make_analysis(m_results, m_params.analysis_functions, m_params.analysis_names);
    m_results.n_accept = m_counter.n_accept;
    m_results.n_reject = m_counter.n_reject;
  }

 private:
  SamplingFixtureParams m_params;
  RandomNumberGenerator<EngineType> m_rng;
  bool m_is_complete = false;
  MonteCounter m_counter;
  CountType m_next_sample_count = 0;
  CompletionCheck m_completion_check;
  RunResults m_results;
  Clock m_clock;
};

// RunManager.hh:20-236
template <typename EngineType = default_engine_type>
struct RunManager {
  typedef SamplingFixture<EngineType> fixture_type;
  std::shared_ptr<EngineType> engine;
  std::vector<std::shared_ptr<fixture_type>> sampling_fixtures;
  bool global_cutoff;
  RunManager(std::shared_ptr<EngineType> _engine, std::vector<SamplingFixtureParams> const &params,
             bool _global_cutoff = true)
      : engine(_engine), global_cutoff(_global_cutoff) {
    if (!engine) throw std::runtime_error("Error constructing RunManager: engine==nullptr");
    for (auto const &p : params) sampling_fixtures.push_back(std::make_shared<fixture_type>(p, engine));
  }
  void initialize(Index steps_per_pass) {
    for (auto &f : sampling_fixtures) f->initialize(steps_per_pass);
  }
  bool is_complete() {  // every fixture is consulted (no early exit), :94-112
    bool all_complete = true, any_complete = false;
    for (auto &f : sampling_fixtures) {
      if (f->is_complete())
        any_complete = true;
      else
        all_complete = false;
    }
    if (global_cutoff && any_complete) return true;
    return all_complete;
  }
  void increment_n_accept() {
    for (auto &f : sampling_fixtures) f->increment_n_accept();
  }
  void increment_n_reject() {
    for (auto &f : sampling_fixtures) f->increment_n_reject();
  }
  void increment_step() {
    for (auto &f : sampling_fixtures) f->increment_step();
  }
  void sample_data_by_count_if_due(IsingState const &state) {
    for (auto &f : sampling_fixtures) f->sample_data_by_count_if_due(state);
  }
  void finalize(IsingState const &state) {
    for (auto &f : sampling_fixtures) f->finalize(state);
  }
};

// methods/occupation_metropolis.hh:90-154 with the Ising event generator
template <typename EngineType>
void ising_occupation_metropolis(IsingState &state, SemiGrandCanonicalPotential &potential,
                                 SemiGrandCanonicalEventGenerator<EngineType> &event_generator,
                                 RandomNumberGenerator<EngineType> &rng, RunManager<EngineType> &run_manager) {
  double n_unitcells = static_cast<double>(state.configuration.n_unitcells);
  state.properties.scalar_values["potential_energy"] = 0.;
  double &potential_energy_per_unitcell = state.properties.scalar_values["potential_energy"];
  potential_energy_per_unitcell = potential.per_supercell() / n_unitcells;
  double beta = 1.0 / (KB * state.conditions.scalar_values.at("temperature"));

  run_manager.initialize(state.configuration.n_variable_sites);
  run_manager.sample_data_by_count_if_due(state);
  while (!run_manager.is_complete()) {
    OccEvent const &event = event_generator.propose(rng);
    double delta_potential_energy = potential.occ_delta_per_supercell(event);
    bool accept = metropolis_acceptance(delta_potential_energy, beta, rng);
    if (accept) {
      run_manager.increment_n_accept();
      event_generator.apply(event);
      potential_energy_per_unitcell += (delta_potential_energy / n_unitcells);
    } else {
      run_manager.increment_n_reject();
    }
    run_manager.increment_step();
    run_manager.sample_data_by_count_if_due(state);
  }
  run_manager.finalize(state);
}

// Analysis functions (not in this reference repository; libcasm-clexmonte defines
// them for its own calculators).  Restated from SURVEY Appendix B.10 on the
// samples used for statistics (after equilibration): population covariance.
inline CountType n_samples_for_statistics(RunResults const &r) {
  if (r.completion_check_results.params.requested_precision.size() != 0)
    return r.completion_check_results.convergence_check_results.N_samples_for_statistics;
  return get_n_samples(r.samplers);
}
inline double tail_variance(RunResults const &r, std::string const &sampler_name) {
  auto const &s = *r.samplers.at(sampler_name);
  std::vector<double> x = s.component(0);
  CountType n = n_samples_for_statistics(r);
  if (n <= 0 || n > static_cast<CountType>(x.size())) n = static_cast<CountType>(x.size());
  std::vector<double> t(x.end() - n, x.end());
  double m = mean_of(t.data(), static_cast<Index>(t.size()));
  return variance(t.data(), static_cast<Index>(t.size()), m);
}

}  // namespace monte_oracle
#endif
