// Run management for the device-backed Ising SGC path:
// sampling schedules, sampling fixtures, the run manager and the
// occupation-Metropolis driver that libcasm-clexmonte-style callers use instead
// of basic_occupation_metropolis.  Same names, members, argument meaning and
// error messages as the reference's
//   include/casm/monte/sampling/SamplingParams.hh
//   include/casm/monte/run_management/{SamplingFixture,RunManager,Results,
//                                      ResultsAnalysisFunction}.hh
//   include/casm/monte/methods/occupation_metropolis.hh
// but the main loop does not step site by site on the host: every input of the
// reference's per-step calls (sample_data_by_count_if_due, is_complete,
// write_status_if_due) changes only at counts this header can compute in
// advance (the next sample of a fixture, a count cutoff), so whole blocks of
// passes run on the GPU (cmg_run_passes) between those counts.
//
// Differences from the reference, all forced by the device loop:
//  * sampling BY_TIME is rejected (Metropolis has no simulated time), BY_STEP
//    is accepted when every scheduled count falls on a pass boundary;
//  * state.properties["potential_energy"] is re-evaluated from the configuration
//    when a sample is taken and at the end, instead of being accumulated
//    step by step (occupation_metropolis.hh:105-109, :143): same value up to
//    the rounding of the running sum;
//  * JSON values are carried as JSON text (std::string), and results are
//    written through a callable (results_io_f) -- the Python package supplies
//    the reference's jsonResultsIO layout.
#ifndef CASM_MONTE_B200_RUN_MANAGEMENT_HH
#define CASM_MONTE_B200_RUN_MANAGEMENT_HH

#include <cmath>
#include <limits>

#include "casm_monte_b200/monte.hh"

namespace casm_monte_b200 {

enum class SAMPLE_MODE { BY_STEP, BY_PASS, BY_TIME };   // definitions.hh:20
enum class SAMPLE_METHOD { LINEAR, LOG, CUSTOM };      // definitions.hh:23

/// sampling/SamplingParams.hh:163-227
struct SamplingParams {
  SamplingParams()
      : sample_mode(SAMPLE_MODE::BY_PASS), sample_method(SAMPLE_METHOD::LINEAR), period(1.0),
        begin(1.0), base(std::pow(10.0, 1.0 / 10.0)), shift(10.0), stochastic_sample_period(false),
        do_sample_trajectory(false), do_sample_time(false) {}
  std::vector<std::string> sampler_names;
  std::vector<std::string> json_sampler_names;
  SAMPLE_MODE sample_mode;
  SAMPLE_METHOD sample_method;
  double period;
  double begin;
  double base;
  double shift;
  std::function<double(CountType)> custom_sample_at;
  bool stochastic_sample_period;
  bool do_sample_trajectory;
  bool do_sample_time;
};

/// SamplingParams.hh:229-245: the count (or time) at which sample `sample_index` is due
inline double sample_at(CountType sample_index, SamplingParams const &s) {
  const double n = static_cast<double>(sample_index);
  switch (s.sample_method) {
    case SAMPLE_METHOD::LINEAR:
      return s.begin + s.period * n;
    case SAMPLE_METHOD::LOG:
      return s.begin + std::pow(s.base, (n + s.shift));
    default:
      if (!s.custom_sample_at)
        throw std::runtime_error(
            "Error in sample_at: sample_method==SAMPLE_METHOD::CUSTOM and !custom_sample_at");
      return s.custom_sample_at(sample_index);
  }
}

/// SamplingParams.hh:247-259
template <typename EngineType>
CountType stochastic_count_step(double sample_rate, RandomNumberGenerator<EngineType> &rng) {
  for (CountType dn = 1;; ++dn)
    if (rng.random_real(1.0) < sample_rate) return dn;
}
/// SamplingParams.hh:261-267
template <typename EngineType>
TimeType stochastic_time_step(TimeType sample_rate, RandomNumberGenerator<EngineType> &rng) {
  return -std::log(rng.random_real(1.0)) / sample_rate;
}
/// SamplingParams.hh:270-305 (the reference's CUSTOM branch is unreachable and
/// leaves the rate undefined; here it is an error)
template <typename EngineType>
double stochastic_sample_at(CountType sample_index, SamplingParams const &s,
                            RandomNumberGenerator<EngineType> &rng,
                            std::vector<CountType> const &sample_count,
                            std::vector<TimeType> const &sample_time) {
  if (sample_index == 0) return s.begin;
  const double n = static_cast<double>(sample_index);
  double rate;
  if (s.sample_method == SAMPLE_METHOD::LINEAR)
    rate = 1.0 / s.period;
  else if (s.sample_method == SAMPLE_METHOD::LOG)
    rate = 1.0 / (std::log(s.base) * std::pow(s.base, (n + s.shift)));
  else
    throw std::runtime_error(
        "Error in stochastic_sample_at: sample_method==SAMPLE_METHOD::CUSTOM is not supported");
  if (s.sample_mode == SAMPLE_MODE::BY_TIME)
    return sample_time.back() + stochastic_time_step(rate, rng);
  return static_cast<double>(sample_count.back() + stochastic_count_step(rate, rng));
}

/// run_management/SamplingFixture.hh:79-119, plus advance_passes for drivers
/// that move a whole number of passes at once
struct MonteCounter {
  MonteCounter() { reset(SAMPLE_MODE::BY_PASS, 1); }
  SAMPLE_MODE sample_mode;
  CountType steps_per_pass, step, pass, count;
  TimeType time;
  BigCountType n_accept, n_reject;
  void reset(SAMPLE_MODE _sample_mode, CountType _steps_per_pass) {
    sample_mode = _sample_mode;
    steps_per_pass = _steps_per_pass;
    step = pass = count = 0;
    time = 0.0;
    n_accept = n_reject = 0;
  }
  void increment_n_accept() { ++n_accept; }
  void increment_n_reject() { ++n_reject; }
  void increment_step() {
    ++step;
    if (sample_mode == SAMPLE_MODE::BY_STEP) ++count;
    if (step == steps_per_pass) {
      ++pass;
      if (sample_mode != SAMPLE_MODE::BY_STEP) ++count;
      step = 0;
    }
  }
  void set_time(double event_time) { time = event_time; }
  /// == n_passes * steps_per_pass calls of increment_step (requires step == 0)
  void advance_passes(CountType n_passes, BigCountType d_accept, BigCountType d_reject) {
    if (step != 0) throw std::runtime_error("MonteCounter::advance_passes: not at a pass boundary");
    pass += n_passes;
    count += (sample_mode == SAMPLE_MODE::BY_STEP) ? n_passes * steps_per_pass : n_passes;
    n_accept += d_accept;
    n_reject += d_reject;
  }
  /// steps per unit of `count`
  CountType steps_per_count() const {
    return sample_mode == SAMPLE_MODE::BY_STEP ? 1 : steps_per_pass;
  }
};

/// sampling/StateSamplingFunction.hh (json form); the value is JSON text
struct jsonStateSamplingFunction {
  std::string name, description;
  std::function<std::string()> function;
  std::string operator()() const { return function(); }
};
typedef std::map<std::string, jsonStateSamplingFunction> jsonStateSamplingFunctionMap;
struct jsonSampler {
  std::vector<std::string> values;
};

struct Results;
/// run_management/ResultsAnalysisFunction.hh:21-63
struct ResultsAnalysisFunction {
  ResultsAnalysisFunction(std::string _name, std::string _description, std::vector<Index> _shape,
                          std::function<std::vector<double>(Results const &)> _function,
                          std::optional<std::vector<std::string>> _component_names = std::nullopt)
      : name(std::move(_name)), description(std::move(_description)), shape(std::move(_shape)),
        component_names(_component_names.has_value() ? *_component_names
                                                     : default_component_names(shape)),
        function(std::move(_function)) {}
  std::string name, description;
  std::vector<Index> shape;
  std::vector<std::string> component_names;
  std::function<std::vector<double>(Results const &)> function;
  std::vector<double> operator()(Results const &results) const { return function(results); }
};
typedef std::map<std::string, ResultsAnalysisFunction> ResultsAnalysisFunctionMap;

/// run_management/Results.hh:14-120
struct Results {
  Results(std::vector<std::string> _sampler_names, StateSamplingFunctionMap _sampling_functions,
          std::vector<std::string> _json_sampler_names,
          jsonStateSamplingFunctionMap _json_sampling_functions,
          ResultsAnalysisFunctionMap _analysis_functions)
      : sampler_names(std::move(_sampler_names)), sampling_functions(std::move(_sampling_functions)),
        json_sampler_names(std::move(_json_sampler_names)),
        json_sampling_functions(std::move(_json_sampling_functions)),
        analysis_functions(std::move(_analysis_functions)), sample_weight(std::vector<Index>{}) {}
  std::vector<std::string> sampler_names;
  StateSamplingFunctionMap sampling_functions;
  std::vector<std::string> json_sampler_names;
  jsonStateSamplingFunctionMap json_sampling_functions;
  ResultsAnalysisFunctionMap analysis_functions;
  std::optional<double> initial_memory_used_MiB, final_memory_used_MiB;
  std::optional<TimeType> elapsed_clocktime;
  SamplerMap samplers;
  std::map<std::string, std::shared_ptr<jsonSampler>> json_samplers;
  std::map<std::string, std::vector<double>> analysis;
  std::vector<CountType> sample_count;
  std::vector<TimeType> sample_time;
  Sampler sample_weight;
  std::vector<TimeType> sample_clocktime;
  std::vector<std::vector<int>> sample_trajectory;  // occupation vectors
  CompletionCheckResults completion_check_results;
  BigCountType n_accept = 0, n_reject = 0;

  void reset() {
    initial_memory_used_MiB.reset();
    final_memory_used_MiB.reset();
    elapsed_clocktime.reset();
    samplers.clear();
    json_samplers.clear();
    analysis.clear();
    sample_count.clear();
    sample_time.clear();
    sample_weight.clear();
    sample_clocktime.clear();
    sample_trajectory.clear();
    completion_check_results.full_reset();
    n_accept = n_reject = 0;
    for (auto const &sampler_name : sampler_names) {
      auto it = sampling_functions.find(sampler_name);
      if (it == sampling_functions.end()) {
        std::stringstream ss;
        ss << "Results::reset error." << std::endl
           << "Failed to find sampling function '" << sampler_name << "'." << std::endl;
        throw std::runtime_error(ss.str());
      }
      auto const &f = it->second;
      samplers.emplace(f.name, std::make_shared<Sampler>(f.shape, f.component_names));
    }
    for (auto const &name : json_sampler_names) {
      auto it = json_sampling_functions.find(name);
      if (it == json_sampling_functions.end()) {
        std::stringstream ss;
        ss << "Results::reset error." << std::endl
           << "Failed to find json sampling function '" << name << "'." << std::endl;
        throw std::runtime_error(ss.str());
      }
      json_samplers.emplace(it->second.name, std::make_shared<jsonSampler>());
    }
  }
};

// Results.hh:122-215: accessors
inline bool is_auto_converge_mode(Results const &r) {
  return r.completion_check_results.params.requested_precision.size() != 0;
}
inline CountType N_samples(Results const &r) { return get_n_samples(r.samplers); }
inline CountType N_samples_for_statistics(Results const &r) {
  return is_auto_converge_mode(r)
             ? r.completion_check_results.convergence_check_results.N_samples_for_statistics
             : N_samples(r);
}
inline CountType N_samples_for_all_to_equilibrate(Results const &r) {
  return r.completion_check_results.equilibration_check_results.N_samples_for_all_to_equilibrate;
}
inline bool all_equilibrated(Results const &r) {
  return r.completion_check_results.equilibration_check_results.all_equilibrated;
}
inline bool all_converged(Results const &r) {
  return r.completion_check_results.convergence_check_results.all_converged;
}
inline double acceptance_rate(Results const &r) {
  return static_cast<double>(r.n_accept) / static_cast<double>(r.n_accept + r.n_reject);
}

/// ResultsAnalysisFunction.hh:107-136: unknown names are skipped, a throwing
/// function yields NaNs (and a message on stderr)
inline std::map<std::string, std::vector<double>> make_analysis(
    Results const &results, ResultsAnalysisFunctionMap const &analysis_functions,
    std::vector<std::string> const &analysis_names) {
  std::map<std::string, std::vector<double>> analysis;
  for (auto const &name : analysis_names) {
    auto it = analysis_functions.find(name);
    if (it == analysis_functions.end()) continue;
    auto const &f = it->second;
    try {
      analysis.emplace(f.name, f(results));
    } catch (std::exception &e) {
      std::cerr << "Results analysis '" << it->first << "' failed: " << e.what() << std::endl;
      analysis.emplace(f.name, std::vector<double>(f.component_names.size(),
                                                   std::numeric_limits<double>::quiet_NaN()));
    }
  }
  return analysis;
}

/// Results.hh:217-330: statistics of one sampled quantity as they go to summary.json
struct QuantityStats {
  QuantityStats(std::string const &quantity_name, Sampler const &sampler, Results const &results)
      : shape(sampler.shape()), is_scalar(shape.size() == 0),
        component_names(sampler.component_names()) {
    auto const &calc_statistics_f = results.completion_check_results.params.calc_statistics_f;
    if (calc_statistics_f == nullptr)
      throw std::runtime_error("Error in QuantityStats: calc_statistics_f == nullptr");
    auto const &requested = results.completion_check_results.params.requested_precision;
    auto tail_stats = [&](Index component_index, CountType N_stats) {
      std::vector<double> x = sampler.component(component_index);
      std::vector<double> t(x.end() - std::min<size_t>(x.size(), static_cast<size_t>(N_stats)), x.end());
      std::vector<double> w;
      if (results.sample_weight.n_samples() != 0) {
        std::vector<double> ww = results.sample_weight.component(0);
        w.assign(ww.end() - std::min<size_t>(ww.size(), t.size()), ww.end());
      }
      return calc_statistics_f(t, w);
    };
    Index i = 0;
    for (auto const &component_name : component_names) {
      SamplerComponent key(quantity_name, i, component_name);
      const bool requested_to_converge = requested.find(key) != requested.end();
      if (is_auto_converge_mode(results)) {
        if (N_samples_for_statistics(results) == 0) {
          is_converged.push_back(requested_to_converge ? std::optional<bool>(false) : std::nullopt);
          component_stats.push_back(std::nullopt);
        } else if (requested_to_converge) {
          auto const &r =
              results.completion_check_results.convergence_check_results.individual_results.find(key)->second;
          is_converged.push_back(r.is_converged);
          component_stats.push_back(r.stats);
        } else {
          is_converged.push_back(std::nullopt);
          component_stats.push_back(tail_stats(i, N_samples_for_statistics(results)));
        }
      } else {
        is_converged.push_back(std::nullopt);
        component_stats.push_back(tail_stats(i, sampler.n_samples()));
      }
      ++i;
    }
  }
  std::vector<Index> shape;
  bool is_scalar;
  std::vector<std::string> component_names;
  std::vector<std::optional<bool>> is_converged;
  std::vector<std::optional<BasicStatistics>> component_stats;
};

/// run_management/io/ResultsIO.hh: write(results, conditions, run_index)
typedef std::function<void(Results const &, ValueMap const &, Index)> ResultsIOFunction;

/// run_management/SamplingFixture.hh:24-77
struct SamplingFixtureParams {
  SamplingFixtureParams(std::string _label, StateSamplingFunctionMap _sampling_functions,
                        jsonStateSamplingFunctionMap _json_sampling_functions,
                        ResultsAnalysisFunctionMap _analysis_functions, SamplingParams _sampling_params,
                        CompletionCheckParams _completion_check_params,
                        std::vector<std::string> _analysis_names = {},
                        ResultsIOFunction _results_io_f = nullptr, MethodLog _method_log = MethodLog())
      : label(std::move(_label)), sampling_functions(std::move(_sampling_functions)),
        json_sampling_functions(std::move(_json_sampling_functions)),
        analysis_functions(std::move(_analysis_functions)), sampling_params(std::move(_sampling_params)),
        completion_check_params(std::move(_completion_check_params)),
        analysis_names(std::move(_analysis_names)), results_io_f(std::move(_results_io_f)),
        method_log(std::move(_method_log)) {
    for (auto const &name : sampling_params.sampler_names)
      if (!sampling_functions.count(name)) {
        std::stringstream ss;
        ss << "SamplingFixtureParams constructor error: No sampling function for '" << name << "'";
        throw std::runtime_error(ss.str());
      }
    for (auto const &name : sampling_params.json_sampler_names)
      if (!json_sampling_functions.count(name)) {
        std::stringstream ss;
        ss << "SamplingFixtureParams constructor error: No sampling function for '" << name << "'";
        throw std::runtime_error(ss.str());
      }
  }
  std::string label;
  StateSamplingFunctionMap sampling_functions;
  jsonStateSamplingFunctionMap json_sampling_functions;
  ResultsAnalysisFunctionMap analysis_functions;
  SamplingParams sampling_params;
  CompletionCheckParams completion_check_params;
  std::vector<std::string> analysis_names;
  ResultsIOFunction results_io_f;
  MethodLog method_log;
};

/// run_management/SamplingFixture.hh:121-644
template <typename EngineType = default_engine_type>
class SamplingFixture {
 public:
  typedef IsingState state_type;
  SamplingFixture(SamplingFixtureParams const &_params, std::shared_ptr<EngineType> _engine)
      : m_params(_params), m_random_number_generator(_engine), m_n_samples(0), m_count(0),
        m_is_complete(false), m_next_sample_count(0), m_next_sample_time(0.0),
        m_completion_check(m_params.completion_check_params),
        m_results(m_params.sampling_params.sampler_names, m_params.sampling_functions,
                  m_params.sampling_params.json_sampler_names, m_params.json_sampling_functions,
                  m_params.analysis_functions) {}

  std::string label() const { return m_params.label; }
  SamplingFixtureParams const &params() const { return m_params; }
  MonteCounter const &counter() const { return m_counter; }
  Results const &results() const { return m_results; }
  CompletionCheck const &completion_check() const { return m_completion_check; }
  CompletionCheckResults const &completion_check_results() const { return m_completion_check.results(); }
  CountType next_sample_count() const { return m_next_sample_count; }
  TimeType next_sample_time() const { return m_next_sample_time; }

  void initialize(Index steps_per_pass) {
    m_n_samples = 0;
    m_count = 0;
    m_is_complete = false;
    m_counter.reset(m_params.sampling_params.sample_mode, steps_per_pass);
    m_completion_check.reset();
    m_results.reset();
    if (m_params.sampling_params.sample_mode == SAMPLE_MODE::BY_TIME) {
      m_next_sample_count = 0;
      m_next_sample_time = this->sample_at(static_cast<CountType>(m_results.sample_time.size()));
      if (m_next_sample_time < 0.0)
        throw std::runtime_error("Error: sampling period parameter error, next_sample_time < 0.0");
    } else {
      m_next_sample_time = 0.0;
      m_next_sample_count = static_cast<CountType>(
          std::round(this->sample_at(static_cast<CountType>(m_results.sample_count.size()))));
      if (m_next_sample_count < 0)
        throw std::runtime_error("Error: sampling period parameter error, next_sample_count < 0");
    }
    m_params.method_log.log.restart_clock();
    m_params.method_log.log.begin_lap();
  }

  bool is_complete() {
    if (m_is_complete) return true;
    LogClock &log = m_params.method_log.log;
    if (m_params.sampling_params.do_sample_time)
      m_is_complete = m_completion_check.is_complete(m_results.samplers, m_results.sample_weight,
                                                     m_counter.count, m_counter.time, log);
    else
      m_is_complete =
          m_completion_check.is_complete(m_results.samplers, m_results.sample_weight, m_counter.count, log);
    return m_is_complete;
  }

  /// {"run_index", "time", "completion_check_results"} to the fixture's log file
  void write_status(Index run_index) {
    if (m_params.method_log.logfile_path.empty()) return;
    m_params.method_log.reset();
    LogClock &log = m_params.method_log.log;
    (*log.out) << "{\"run_index\": " << run_index << ", \"time\": " << json_number(log.time_s())
               << ", \"completion_check_results\": " << to_json_text(m_completion_check.results()) << "}"
               << std::endl;
    log.begin_lap();
  }
  void write_status_if_due(Index run_index) {
    std::optional<double> &log_frequency = m_params.method_log.log_frequency;
    if (!log_frequency.has_value()) return;
    if (m_n_samples != get_n_samples(m_results.samplers) || m_count != m_counter.count) {
      m_n_samples = get_n_samples(m_results.samplers);
      m_count = m_counter.count;
      if (m_params.method_log.log.lap_time() > *log_frequency) write_status(run_index);
    }
  }

  void increment_n_accept() { m_counter.increment_n_accept(); }
  void increment_n_reject() { m_counter.increment_n_reject(); }
  void increment_step() { m_counter.increment_step(); }
  void advance_passes(CountType n_passes, BigCountType d_accept, BigCountType d_reject) {
    m_counter.advance_passes(n_passes, d_accept, d_reject);
  }
  void set_time(double event_time) { m_counter.set_time(event_time); }
  void push_back_sample_weight(double weight) { m_results.sample_weight.push_back(weight); }

  void sample_data(state_type const &state) {
    m_results.sample_count.push_back(m_counter.count);
    if (m_params.sampling_params.do_sample_time) m_results.sample_time.push_back(m_counter.time);
    m_results.sample_clocktime.push_back(m_params.method_log.log.time_s());
    if (m_params.sampling_params.do_sample_trajectory)
      m_results.sample_trajectory.push_back(state.configuration.occupation());
    for (auto const &name : m_params.sampling_params.sampler_names) {
      auto it = m_params.sampling_functions.find(name);
      if (it == m_params.sampling_functions.end()) {
        std::stringstream ss;
        ss << "Error in SamplingFixture::sample_data: did not find sampling function '" << name << "'";
        throw std::runtime_error(ss.str());
      }
      m_results.samplers.at(name)->push_back(it->second());
    }
    for (auto const &name : m_params.sampling_params.json_sampler_names) {
      auto it = m_params.json_sampling_functions.find(name);
      if (it == m_params.json_sampling_functions.end()) {
        std::stringstream ss;
        ss << "Error in SamplingFixture::sample_data: did not find json sampling function'" << name << "'";
        throw std::runtime_error(ss.str());
      }
      m_results.json_samplers.at(name)->values.push_back(it->second());
    }
    if (m_params.sampling_params.sample_mode == SAMPLE_MODE::BY_TIME) {
      m_next_sample_time = this->sample_at(static_cast<CountType>(m_results.sample_time.size()));
      if (m_next_sample_time <= m_counter.time)
        throw std::runtime_error(
            "Error: state sampling period parameter error, next_sample_time <= current time");
    } else {
      m_next_sample_count = static_cast<CountType>(
          std::round(this->sample_at(static_cast<CountType>(m_results.sample_count.size()))));
      if (m_next_sample_count <= m_counter.count)
        throw std::runtime_error(
            "Error: state sampling period parameter error, next_sample_count <= current count");
    }
  }
  void sample_data_by_count_if_due(state_type const &state) {
    if (m_params.sampling_params.sample_mode != SAMPLE_MODE::BY_TIME &&
        m_counter.count == m_next_sample_count)
      sample_data(state);
  }
  double sample_at(CountType sample_index) {
    if (m_params.sampling_params.stochastic_sample_period)
      return stochastic_sample_at(sample_index, m_params.sampling_params, m_random_number_generator,
                                  m_results.sample_count, m_results.sample_time);
    return casm_monte_b200::sample_at(sample_index, m_params.sampling_params);
  }

  void finalize(state_type const &state, Index run_index) {
    m_results.elapsed_clocktime = m_params.method_log.log.time_s();
    m_results.completion_check_results = m_completion_check.results();
    m_results.analysis = make_analysis(m_results, m_params.analysis_functions, m_params.analysis_names);
    m_results.n_accept = m_counter.n_accept;
    m_results.n_reject = m_counter.n_reject;
    if (m_params.results_io_f) m_params.results_io_f(m_results, state.conditions, run_index);
    write_status(run_index);
  }

  /// Steps from now to the next count at which this fixture does something a
  /// per-step loop would notice: its next sample, or a cutoff on `count`.
  /// One pass at most when a clock-based cutoff is set.
  CountType steps_to_next_event() const {
    const CountType unit = m_counter.steps_per_count();
    CountType best = std::numeric_limits<CountType>::max();
    auto consider = [&](CountType at_count) {
      if (at_count > m_counter.count) best = std::min(best, (at_count - m_counter.count) * unit);
    };
    if (m_params.sampling_params.sample_mode != SAMPLE_MODE::BY_TIME) consider(m_next_sample_count);
    CutoffCheckParams const &c = m_params.completion_check_params.cutoff_params;
    if (c.min_count) consider(*c.min_count);
    if (c.max_count) consider(*c.max_count);
    if (c.min_clocktime || c.max_clocktime) best = std::min(best, m_counter.steps_per_pass);
    return best;
  }

 private:
  SamplingFixtureParams m_params;
  RandomNumberGenerator<EngineType> m_random_number_generator;
  Index m_n_samples, m_count;
  bool m_is_complete;
  MonteCounter m_counter;
  CountType m_next_sample_count;
  TimeType m_next_sample_time;
  CompletionCheck m_completion_check;
  Results m_results;
};

/// run_management/RunManager.hh:20-236 (count-based parts)
template <typename EngineType = default_engine_type>
struct RunManager {
  typedef EngineType engine_type;
  typedef IsingState state_type;
  typedef SamplingFixture<EngineType> sampling_fixture_type;
  typedef std::function<bool(sampling_fixture_type const &, state_type const &)> BreakPointCheck;

  Index run_index;
  std::shared_ptr<engine_type> engine;
  std::vector<std::shared_ptr<sampling_fixture_type>> sampling_fixtures;
  bool global_cutoff;
  std::map<std::string, BreakPointCheck> break_point_checks;
  bool break_point_set;

  RunManager(std::shared_ptr<engine_type> _engine,
             std::vector<SamplingFixtureParams> const &_sampling_fixture_params, bool _global_cutoff = true)
      : run_index(0), engine(_engine), global_cutoff(_global_cutoff), break_point_set(false) {
    if (!engine) throw std::runtime_error("Error constructing RunManager: engine==nullptr");
    for (auto const &params : _sampling_fixture_params)
      sampling_fixtures.emplace_back(std::make_shared<sampling_fixture_type>(params, engine));
  }
  void initialize(Index steps_per_pass) {
    for (auto &f : sampling_fixtures) f->initialize(steps_per_pass);
    break_point_set = false;
  }
  bool is_break_point() const { return break_point_set; }
  /// every fixture is consulted (no early exit) so that status files carry the
  /// latest completion-check results, :94-112
  bool is_complete() {
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
  void write_status_if_due() {
    for (auto &f : sampling_fixtures) f->write_status_if_due(run_index);
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
  void advance_passes(CountType n_passes, BigCountType d_accept, BigCountType d_reject) {
    for (auto &f : sampling_fixtures) f->advance_passes(n_passes, d_accept, d_reject);
  }
  void sample_data_by_count_if_due(state_type const &state) {
    for (auto &fp : sampling_fixtures) {
      auto &f = *fp;
      if (f.params().sampling_params.sample_mode == SAMPLE_MODE::BY_TIME) continue;
      if (f.counter().count == f.next_sample_count()) {
        f.sample_data(state);
        auto it = break_point_checks.find(f.label());
        if (it != break_point_checks.end()) break_point_set = it->second(f, state);
      }
    }
  }
  void finalize(state_type const &final_state) {
    for (auto &f : sampling_fixtures) f->finalize(final_state, run_index);
  }
  /// sum of the convergence checks done so far (device drivers re-consult
  /// is_complete while this is still moving)
  Index n_checks() const {
    Index n = 0;
    for (auto const &f : sampling_fixtures) n += f->completion_check().n_checks();
    return n;
  }
  CountType steps_to_next_event() const {
    CountType best = std::numeric_limits<CountType>::max();
    for (auto const &f : sampling_fixtures) best = std::min(best, f->steps_to_next_event());
    return best;
  }
};

/// methods/occupation_metropolis.hh:90-154 for the Ising SGC calculator, with
/// the sequence of passes between events run on the GPU.
/// update_mode: "auto" | "checkerboard" | "serial_reference" (as
/// SemiGrandCanonicalCalculator::run); in serial_reference mode the trajectory,
/// the samples and the engine state are those of the reference's loop.
template <typename EngineType = default_engine_type>
void occupation_metropolis(SemiGrandCanonicalCalculator &mc_calculator, IsingState &state,
                           RunManager<EngineType> &run_manager, std::string update_mode = "auto") {
  static_assert(std::is_same<EngineType, default_engine_type>::value,
                "the device loop restates std::mt19937_64");
  auto &calc = mc_calculator;
  calc.state = &state;
  calc.conditions = std::make_shared<SemiGrandCanonicalConditions>(
      SemiGrandCanonicalConditions::from_values(state.conditions));
  if (calc.conditions->exchange_potential.size() != 1)
    throw std::runtime_error("Error in occupation_metropolis: exchange_potential must have 1 component");
  calc.potential.set_state(&state, calc.conditions);
  for (auto const &fp : run_manager.sampling_fixtures)
    if (fp->params().sampling_params.sample_mode == SAMPLE_MODE::BY_TIME)
      throw std::runtime_error(
          "Error in occupation_metropolis: sampling BY_TIME is not defined for Metropolis on the device");

  IsingConfiguration &config = state.configuration;
  const CountType steps_per_pass = config.n_variable_sites;
  const double n_unitcells = static_cast<double>(config.n_unitcells);

  DeviceLattice &dev = config.device();
  cmg_context *ctx = dev.ctx();
  dev.check(cmg_set_model(ctx, calc.potential.formation_energy_calculator.J,
                          calc.potential.formation_energy_calculator.lattice_type));
  dev.check(cmg_set_conditions(ctx, 0, calc.conditions->temperature, calc.conditions->exchange_potential[0]));
  dev.check(cmg_reset_counters(ctx));
  dev.check(cmg_clear_samples(ctx));

  bool even = true;
  for (int s : config.shape) even = even && (s % 2 == 0);
  int mode;
  if (update_mode == "serial_reference") mode = CMG_MODE_SERIAL_REFERENCE;
  else if (update_mode == "checkerboard") mode = CMG_MODE_CHECKERBOARD;
  else if (update_mode == "auto") mode = even ? CMG_MODE_CHECKERBOARD : CMG_MODE_SERIAL_REFERENCE;
  else throw std::runtime_error("Error in occupation_metropolis: unknown update_mode '" + update_mode + "'");

  auto push_engine = [&]() {
    uint64_t words[312];
    int pos = 0;
    engine_to_words(*run_manager.engine, words, &pos);
    dev.check(cmg_set_mt19937_64_state(ctx, 0, words, pos));
  };
  auto pull_engine = [&]() {
    uint64_t words[312];
    int pos = 0;
    dev.check(cmg_get_mt19937_64_state(ctx, 0, words, &pos));
    words_to_engine(words, pos, *run_manager.engine);
  };
  if (mode == CMG_MODE_SERIAL_REFERENCE) {
    push_engine();
  } else {
    dev.check(cmg_seed_philox(ctx, (*run_manager.engine)()));  // one draw seeds the Philox key
    dev.check(cmg_set_pass_counter(ctx, 0));
  }
  // fixtures with a stochastic sample period draw from the shared engine between
  // blocks of passes; in serial mode the device owns the stream in between
  bool any_stochastic = false;
  for (auto const &fp : run_manager.sampling_fixtures)
    any_stochastic = any_stochastic || fp->params().sampling_params.stochastic_sample_period;
  const bool share_engine = any_stochastic && mode == CMG_MODE_SERIAL_REFERENCE;

  auto set_potential_energy_property = [&]() {
    state.properties.scalar_values["potential_energy"] = calc.potential.per_supercell() / n_unitcells;
  };
  state.properties.scalar_values["potential_energy"] = 0.;
  set_potential_energy_property();

  run_manager.initialize(steps_per_pass);
  run_manager.sample_data_by_count_if_due(state);
  if (share_engine) push_engine();

  int64_t acc_prev = 0, rej_prev = 0;
  while (true) {
    // consult is_complete as the per-step loop would; repeat while scheduled
    // convergence checks are still being caught up
    bool done = false;
    while (true) {
      const Index before = run_manager.n_checks();
      done = run_manager.is_complete();
      if (done || run_manager.n_checks() == before) break;
    }
    if (done) break;
    run_manager.write_status_if_due();

    const CountType steps = run_manager.steps_to_next_event();
    if (steps == std::numeric_limits<CountType>::max())
      throw std::runtime_error("Error in occupation_metropolis: nothing is scheduled (no samples, no cutoffs)");
    if (steps % steps_per_pass != 0)
      throw std::runtime_error(
          "Error in occupation_metropolis: a BY_STEP sample or cutoff does not fall on a pass "
          "boundary; the device loop advances whole passes");
    const CountType n_run = steps / steps_per_pass;
    dev.check(cmg_run_passes(ctx, n_run, mode, 0));
    config.mark_device_modified();
    int64_t np = 0, na = 0, nr = 0;
    dev.check(cmg_counters(ctx, 0, &np, &na, &nr));
    run_manager.advance_passes(n_run, na - acc_prev, nr - rej_prev);
    acc_prev = na;
    rej_prev = nr;

    bool sample_due = false;
    for (auto const &fp : run_manager.sampling_fixtures)
      sample_due = sample_due || fp->counter().count == fp->next_sample_count();
    if (sample_due) {
      if (share_engine) pull_engine();
      set_potential_energy_property();
      bool needs_host_state = false;
      for (auto const &fp : run_manager.sampling_fixtures)
        needs_host_state = needs_host_state || fp->params().sampling_params.do_sample_trajectory ||
                           !fp->params().sampling_params.json_sampler_names.empty();
      if (needs_host_state) config.pull();
      run_manager.sample_data_by_count_if_due(state);
      if (share_engine) push_engine();
    }
  }

  config.pull();
  if (mode == CMG_MODE_SERIAL_REFERENCE) pull_engine();
  set_potential_energy_property();
  calc.last_kernel = cmg_kernel_variant(ctx);
  run_manager.finalize(state);
}

// ---------------------------------------------------------------------------
// Results analysis functions for the Ising SGC calculator.  The reference
// defines ResultsAnalysisFunction but ships no instances (libcasm-clexmonte
// does, for its own calculators); these follow SURVEY Appendix B.10:
//   heat_capacity  = N * Var(potential_energy) / (KB * T^2)      [eV/K per unit cell]
//   susceptibility = N * Var(param_composition) / (KB * T)       [1/eV per unit cell]
// with the population variance of the samples used for statistics (after
// equilibration when convergence was requested).
// ---------------------------------------------------------------------------
namespace analysis_impl {
inline double tail_variance(Results const &results, std::string const &sampler_name) {
  auto it = results.samplers.find(sampler_name);
  if (it == results.samplers.end())
    throw std::runtime_error("analysis: sampler '" + sampler_name + "' was not sampled");
  std::vector<double> x = it->second->component(0);
  CountType n = N_samples_for_statistics(results);
  if (n <= 0 || n > static_cast<CountType>(x.size())) n = static_cast<CountType>(x.size());
  if (n == 0) throw std::runtime_error("analysis: no samples");
  double mean = 0.0;
  for (auto p = x.end() - n; p != x.end(); ++p) mean += *p;
  mean /= static_cast<double>(n);
  double var = 0.0;
  for (auto p = x.end() - n; p != x.end(); ++p) var += (*p - mean) * (*p - mean);
  return var / static_cast<double>(n);
}
}  // namespace analysis_impl

inline ResultsAnalysisFunction make_heat_capacity_f(std::shared_ptr<SemiGrandCanonicalCalculator> mc_calculator) {
  if (mc_calculator == nullptr)
    throw std::runtime_error("Error in make_heat_capacity_f: mc_calculator == nullptr");
  std::weak_ptr<SemiGrandCanonicalCalculator> weak = mc_calculator;
  return ResultsAnalysisFunction(
      "heat_capacity", "Heat capacity (per unit cell) = N * Var(potential_energy) / (KB * T^2)", {},
      [weak](Results const &results) {
        auto calc = weak.lock();
        if (!calc || !calc->state) throw std::runtime_error("heat_capacity: calculator has no state");
        const double T = calc->conditions->temperature;
        const double N = static_cast<double>(calc->state->configuration.n_unitcells);
        return std::vector<double>{N * analysis_impl::tail_variance(results, "potential_energy") / (KB * T * T)};
      });
}
inline ResultsAnalysisFunction make_susceptibility_f(std::shared_ptr<SemiGrandCanonicalCalculator> mc_calculator) {
  if (mc_calculator == nullptr)
    throw std::runtime_error("Error in make_susceptibility_f: mc_calculator == nullptr");
  std::weak_ptr<SemiGrandCanonicalCalculator> weak = mc_calculator;
  return ResultsAnalysisFunction(
      "susceptibility", "Susceptibility (per unit cell) = N * Var(param_composition) / (KB * T)", {},
      [weak](Results const &results) {
        auto calc = weak.lock();
        if (!calc || !calc->state) throw std::runtime_error("susceptibility: calculator has no state");
        const double T = calc->conditions->temperature;
        const double N = static_cast<double>(calc->state->configuration.n_unitcells);
        return std::vector<double>{N * analysis_impl::tail_variance(results, "param_composition") / (KB * T)};
      });
}

}  // namespace casm_monte_b200
#endif
