// oracle_bench.cc -- timed CPU baseline (TEST / BENCH INFRASTRUCTURE ONLY).
//
// Runs the restated reference loop (serial random-site Metropolis,
// std::mt19937_64, is_complete evaluated every step, O(N) observables at every
// sample) with ONE INDEPENDENT CHAIN PER THREAD and reports the reference's own
// throughput definition, Steps/Second = n_pass * n_steps_per_pass / time_s
// (include/casm/monte/methods/basic_occupation_metropolis.hh:232-240), summed
// over chains.
//
// usage: oracle_bench n0 n1 T mu n_passes sample_period use_nlist n_threads [seed]
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "monte_oracle.hh"

using namespace monte_oracle;

struct ChainResult {
  double seconds = 0.0;
  long n_pass = 0;
  long long n_accept = 0;
  double last_x = 0.0, last_e = 0.0;
};

static void run_chain(int n0, int n1, double T, double mu, long n_passes,
                      int sample_period, bool use_nlist, uint64_t seed,
                      ChainResult *out) {
  IsingConfiguration config(std::vector<int>{n0, n1}, 1);
  ValueMap cond;
  cond.scalar_values["temperature"] = T;
  cond.vector_values["exchange_potential"] = std::vector<double>{mu};
  IsingState state(config, cond);
  auto system = std::make_shared<IsingSystem>(
      IsingFormationEnergy(0.1, 1, use_nlist), IsingParamComposition());
  auto mc = std::make_shared<SemiGrandCanonicalCalculator>(system);
  StateSamplingFunctionMap fns;
  if (sample_period > 0) {
    for (auto const &f :
         {make_parametric_composition_f(mc), make_formation_energy_f(mc),
          make_potential_energy_f(mc)})
      fns.emplace(f.name, f);
  }
  CompletionCheckParams p;
  p.cutoff_params.max_count = n_passes;
  SemiGrandCanonicalCalculator::event_generator_type gen;
  auto engine = std::make_shared<std::mt19937_64>(seed);
  std::optional<MethodLog> log = MethodLog();
  auto no_status = [](BasicOccupationMetropolisData const &, MethodLog &) {};
  int period = sample_period > 0 ? sample_period : static_cast<int>(n_passes + 1);
  auto t0 = std::chrono::steady_clock::now();
  mc->run(state, fns, p, gen, period, log, engine, no_status);
  auto t1 = std::chrono::steady_clock::now();
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  out->n_pass = mc->data->n_pass;
  out->n_accept = mc->data->n_accept;
  if (sample_period > 0 && mc->data->samplers.count("param_composition")) {
    auto const &s = *mc->data->samplers.at("param_composition");
    if (s.n_samples()) out->last_x = s.component(0).back();
    auto const &e = *mc->data->samplers.at("potential_energy");
    if (e.n_samples()) out->last_e = e.component(0).back();
  }
}

int main(int argc, char **argv) {
  if (argc < 9) {
    std::fprintf(stderr,
                 "usage: %s n0 n1 T mu n_passes sample_period use_nlist "
                 "n_threads [seed]\n",
                 argv[0]);
    return 2;
  }
  int n0 = std::atoi(argv[1]), n1 = std::atoi(argv[2]);
  double T = std::atof(argv[3]), mu = std::atof(argv[4]);
  long n_passes = std::atol(argv[5]);
  int sample_period = std::atoi(argv[6]);
  bool use_nlist = std::atoi(argv[7]) != 0;
  int n_threads = std::atoi(argv[8]);
  uint64_t seed = argc > 9 ? std::strtoull(argv[9], nullptr, 10) : 12345ull;
  if (n_threads <= 0) n_threads = std::thread::hardware_concurrency();

  std::vector<ChainResult> res(n_threads);
  std::vector<std::thread> th;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_threads; ++i)
    th.emplace_back(run_chain, n0, n1, T, mu, n_passes, sample_period,
                    use_nlist, seed + 1000003ull * i, &res[i]);
  for (auto &t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  double wall = std::chrono::duration<double>(t1 - t0).count();

  double N = static_cast<double>(n0) * n1;
  double total_steps = 0.0, sum_rate = 0.0;
  long long acc = 0;
  for (auto const &r : res) {
    total_steps += r.n_pass * N;
    sum_rate += r.n_pass * N / r.seconds;
    acc += r.n_accept;
  }
  std::printf(
      "{\"attempts_per_s\": %.6e, \"attempts_per_s_sum_of_chains\": %.6e, "
      "\"wall_s\": %.4f, \"threads\": %d, \"n0\": %d, \"n1\": %d, "
      "\"n_passes\": %ld, \"sample_period\": %d, \"use_nlist\": %d, "
      "\"acceptance\": %.6f, \"last_x\": %.8f, \"last_e_pot\": %.8f}\n",
      total_steps / wall, sum_rate, wall, n_threads, n0, n1, n_passes,
      sample_period, use_nlist ? 1 : 0,
      static_cast<double>(acc) / (total_steps > 0 ? total_steps : 1.0),
      res[0].last_x, res[0].last_e);
  return 0;
}
