// monte_oracle.hh -- CPU restatement of libcasm-monte's Ising SGC Metropolis path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (casmcode_monte_b200/,
// include/) may include, link or call this file.  It is used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs, as the checker and as the timed CPU baseline.
//
// The reference itself cannot be compiled in this image (it needs Eigen and
// the CASMcode_global / CASMcode_crystallography headers, none of which are
// present and there is no network), so this is a dependency-free C++17
// restatement.  Each block cites the reference file:line it follows
// (paths relative to the reference root).
//
// PARITY PINNING STATUS
//   pinned   : dE / energy / composition / potential known answers
//              (tests/unit/monte/Ising_basic_semigrand_canonical_test.cpp:113-267),
//              Sampler layout and growth (python/tests/sampling/test_Sampler.py),
//              completion at max_count (python/tests/sampling/test_CompletionCheck.py:5-37),
//              engine reproducibility (python/tests/test_RandomNumberGeneratory.py).
//   UNPINNED : seeded trajectories and ensemble averages (the reference's run
//              tests assert nothing about values), the last-bit numerics of the
//              statistics (Eigen's vectorised reductions), the value of KB
//              (lives in CASMcode_global), the unit-cell ordering inside
//              Conversions (lives in CASMcode_crystallography).  "parity
//              unpinned" for those; closed-form Onsager results are used as an
//              independent anchor in tests/.
//
// Arithmetic that must be held bit-for-bit is written with the reference's
// exact association; build with -ffp-contract=off (see oracle/Makefile).
#ifndef CASM_MONTE_B200_ORACLE_HH
#define CASM_MONTE_B200_ORACLE_HH

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace monte_oracle {

// include/casm/monte/definitions.hh:25-27 ; CMakeLists.txt:169 (Index == long)
using Index = long;
using CountType = long;
using BigCountType = long long;
using TimeType = double;
// include/casm/monte/definitions.hh:17
using default_engine_type = std::mt19937_64;

// Boltzmann constant, eV/K.  Defined in CASMcode_global
// (casm/global/definitions.hh), which is NOT part of the reference tree:
// value recalled (CODATA-2014), flagged unpinned.  Used at
// include/casm/monte/methods/basic_occupation_metropolis.hh:361.
constexpr double KB = 8.6173303E-05;

// include/casm/monte/definitions.hh:107-113
template <typename P>
P throw_if_null(P ptr, std::string const &what) {
  if (ptr == nullptr) throw std::runtime_error(what);
  return ptr;
}

// ---------------------------------------------------------------------------
// ValueMap  (include/casm/monte/ValueMap.hh:13-74)
// ---------------------------------------------------------------------------
struct ValueMap {
  std::map<std::string, bool> boolean_values;
  std::map<std::string, double> scalar_values;
  std::map<std::string, std::vector<double>> vector_values;
  // matrices kept as (rows, cols, column-major data)
  struct Mat {
    Index rows = 0, cols = 0;
    std::vector<double> data;
  };
  std::map<std::string, Mat> matrix_values;
};

inline bool is_mismatched(ValueMap const &A, ValueMap const &B) {
  for (auto const &kv : B.boolean_values)
    if (!A.boolean_values.count(kv.first)) return true;
  for (auto const &kv : B.scalar_values)
    if (!A.scalar_values.count(kv.first)) return true;
  for (auto const &kv : B.vector_values)
    if (!A.vector_values.count(kv.first)) return true;
  for (auto const &kv : B.matrix_values)
    if (!A.matrix_values.count(kv.first)) return true;
  return false;
}

inline ValueMap make_incremented_values(ValueMap values, ValueMap const &inc,
                                        double n_increment) {
  for (auto const &kv : inc.scalar_values)
    values.scalar_values.at(kv.first) += kv.second * n_increment;
  for (auto const &kv : inc.vector_values) {
    auto &dst = values.vector_values.at(kv.first);
    for (size_t i = 0; i < dst.size(); ++i) dst[i] += kv.second[i] * n_increment;
  }
  for (auto const &kv : inc.matrix_values) {
    auto &dst = values.matrix_values.at(kv.first).data;
    for (size_t i = 0; i < dst.size(); ++i)
      dst[i] += kv.second.data[i] * n_increment;
  }
  return values;
}

// ---------------------------------------------------------------------------
// RandomNumberGenerator  (include/casm/monte/RandomNumberGenerator.hh:15-42)
// libstdc++'s distributions are used directly, so the implementation-defined
// algorithms (Lemire for ints, generate_canonical for reals) come for free.
// ---------------------------------------------------------------------------
template <typename EngineType = default_engine_type>
struct RandomNumberGenerator {
  std::shared_ptr<EngineType> engine;

  explicit RandomNumberGenerator(
      std::shared_ptr<EngineType> _engine = std::shared_ptr<EngineType>())
      : engine(_engine) {
    if (engine == nullptr) {
      engine = std::make_shared<EngineType>();
      std::random_device device;
      engine->seed(device());
    }
  }
  template <typename IntType>
  IntType random_int(IntType maximum_value) {
    return std::uniform_int_distribution<IntType>(0, maximum_value)(*engine);
  }
  template <typename RealType>
  RealType random_real(RealType maximum_value) {
    return std::uniform_real_distribution<RealType>(0, maximum_value)(*engine);
  }
};

// ---------------------------------------------------------------------------
// OccEvent (include/casm/monte/events/OccEvent.hh:57-73), Ising subset
// ---------------------------------------------------------------------------
struct OccEvent {
  std::vector<Index> linear_site_index;
  std::vector<int> new_occ;
};

// ---------------------------------------------------------------------------
// IsingConfiguration (include/casm/monte/ising_cpp/model.hh:19-110)
// 2-d in the reference; a 3-d extension (l = i + n0*(j + n1*k)) is available
// when allow_3d is passed -- the reference throws for anything but 2-d.
// Sizes are Index (long) here; the reference computes shape[0]*shape[1] in int.
// ---------------------------------------------------------------------------
class IsingConfiguration {
 public:
  IsingConfiguration() : IsingConfiguration(std::vector<int>{0, 0}, 1) {}

  explicit IsingConfiguration(std::vector<int> _shape, int fill_value = 1,
                              bool allow_3d = false)
      : shape(std::move(_shape)) {
    if (!(shape.size() == 2 || (allow_3d && shape.size() == 3))) {
      throw std::runtime_error("IsingConfiguration only supports 2d");
    }
    Index n = 1;
    for (int s : shape) n *= static_cast<Index>(s);
    m_occupation.assign(static_cast<size_t>(n), fill_value);
    n_sites = n;
    n_variable_sites = n;
    n_unitcells = n;
  }

  std::vector<int> shape;
  Index n_sites = 0;
  Index n_variable_sites = 0;
  Index n_unitcells = 0;

  std::vector<int> const &occupation() const { return m_occupation; }

  void set_occupation(std::vector<int> const &occupation) {
    if (m_occupation.size() != occupation.size()) {
      throw std::runtime_error("Error in set_occupation: size mismatch");
    }
    m_occupation = occupation;
  }
  int occ(Index l) const { return m_occupation[l]; }
  void set_occ(Index l, int v) { m_occupation[l] = v; }

  // model.hh:73-79 (floor-mod)
  Index within(Index index, int dim) const {
    Index r = index % shape[dim];
    if (r < 0) r += shape[dim];
    return r;
  }
  // model.hh:82-90 ; column-major unrolling
  std::vector<int> from_linear_site_index(Index l) const {
    std::vector<int> mi(shape.size());
    if (shape.size() == 2) {
      mi[0] = static_cast<int>(l % shape[0]);
      mi[1] = static_cast<int>(l / shape[0]);
    } else {
      mi[0] = static_cast<int>(l % shape[0]);
      Index r = l / shape[0];
      mi[1] = static_cast<int>(r % shape[1]);
      mi[2] = static_cast<int>(r / shape[1]);
    }
    return mi;
  }
  // model.hh:93-109
  Index to_linear_site_index(std::vector<int> const &mi) const {
    if (shape.size() == 2) return static_cast<Index>(shape[0]) * mi[1] + mi[0];
    return mi[0] +
           static_cast<Index>(shape[0]) *
               (mi[1] + static_cast<Index>(shape[1]) * mi[2]);
  }
  Index to_linear_site_index(Index row, Index col) const {
    return static_cast<Index>(shape[0]) * col + row;
  }
  Index to_linear_site_index(Index i, Index j, Index k) const {
    return i + static_cast<Index>(shape[0]) *
                   (j + static_cast<Index>(shape[1]) * k);
  }

 private:
  std::vector<int> m_occupation;
};

// model.hh:141-157
class IsingState {
 public:
  IsingState(IsingConfiguration _configuration, ValueMap _conditions,
             ValueMap _properties = ValueMap())
      : configuration(std::move(_configuration)),
        conditions(std::move(_conditions)),
        properties(std::move(_properties)) {}
  IsingConfiguration configuration;
  ValueMap conditions;
  ValueMap properties;
};

// ---------------------------------------------------------------------------
// IsingFormationEnergy (model.hh:164-380)
// ---------------------------------------------------------------------------
class IsingFormationEnergy {
 public:
  typedef IsingState state_type;

  // model.hh:168-181.  The reference's null-check tests the *member* (always
  // nullptr at that point) so the ctor's state argument is ignored; same here.
  IsingFormationEnergy(double _J = 1.0, int _lattice_type = 1,
                       bool _use_nlist = true,
                       state_type const * /*_state*/ = nullptr)
      : J(_J), lattice_type(_lattice_type), state(nullptr),
        m_use_nlist(_use_nlist) {
    if (lattice_type != 1) throw std::runtime_error("Unsupported lattice_type");
  }

  double J;
  int lattice_type;
  state_type const *state;

  // model.hh:210-256.  Neighbour order of the "flower" list:
  // (i+1,j), (i,j+1), (i-1,j), (i,j-1) ; 3-d appends (k+1) to both lists after
  // (j+1) ... see below: bond list = +i, +j, +k ; flower = +i,+j,+k,-i,-j,-k.
  void set_state(state_type const *_state) {
    state = throw_if_null(
        _state, "Error in IsingFormationEnergy::set_state: _state==nullptr");
    if (!m_use_nlist) return;
    IsingConfiguration const &config = state->configuration;
    const int dim = static_cast<int>(config.shape.size());
    m_nlist.clear();
    m_nlist.resize(config.n_sites);
    m_flower_nlist.clear();
    m_flower_nlist.resize(config.n_sites);
    for (Index l = 0; l < config.n_sites; ++l) {
      std::vector<int> mi = config.from_linear_site_index(l);
      for (int d = 0; d < dim; ++d) {
        std::vector<int> nb = mi;
        nb[d] = static_cast<int>(config.within(mi[d] + 1, d));
        Index ln = config.to_linear_site_index(nb);
        m_nlist[l].push_back(ln);
        m_flower_nlist[l].push_back(ln);
      }
      for (int d = 0; d < dim; ++d) {
        std::vector<int> nb = mi;
        nb[d] = static_cast<int>(config.within(mi[d] - 1, d));
        m_flower_nlist[l].push_back(config.to_linear_site_index(nb));
      }
    }
  }

  // model.hh:259-290
  double per_supercell() const {
    IsingConfiguration const &config = state->configuration;
    std::vector<int> const &occ = config.occupation();
    const int dim = static_cast<int>(config.shape.size());
    if (m_use_nlist) {
      double e_formation = 0.0;
      for (Index l = 0; l < config.n_sites; ++l) {
        int nb = 0;
        for (Index ln : m_nlist[l]) nb += occ[ln];
        e_formation += occ[l] * nb;  // int product, accumulated in double
      }
      e_formation *= -J;
      return e_formation;
    }
    // use_nlist == false: sum over lines of (-J * integer dot product), first
    // along dimension 0 ("rows"), then dimension 1 ("cols") [, then 2].
    double e_formation = 0.0;
    if (dim == 2) {
      Index rows = config.shape[0], cols = config.shape[1];
      for (Index i = 0; i < rows; ++i) {
        Index in = config.within(i + 1, 0);
        long dot = 0;
        for (Index j = 0; j < cols; ++j)
          dot += occ[i + rows * j] * occ[in + rows * j];
        e_formation += -J * static_cast<double>(dot);
      }
      for (Index j = 0; j < cols; ++j) {
        Index jn = config.within(j + 1, 1);
        long dot = 0;
        for (Index i = 0; i < rows; ++i)
          dot += occ[i + rows * j] * occ[i + rows * jn];
        e_formation += -J * static_cast<double>(dot);
      }
      return e_formation;
    }
    // 3-d (extension): one plane-dot per index value and direction
    Index n0 = config.shape[0], n1 = config.shape[1], n2 = config.shape[2];
    for (int d = 0; d < 3; ++d) {
      Index nd = config.shape[d];
      for (Index a = 0; a < nd; ++a) {
        Index an = config.within(a + 1, d);
        long dot = 0;
        for (Index k = 0; k < n2; ++k)
          for (Index j = 0; j < n1; ++j)
            for (Index i = 0; i < n0; ++i) {
              Index c[3] = {i, j, k};
              if (c[d] != a) continue;
              Index cn[3] = {i, j, k};
              cn[d] = an;
              dot += occ[c[0] + n0 * (c[1] + n1 * c[2])] *
                     occ[cn[0] + n0 * (cn[1] + n1 * cn[2])];
            }
        e_formation += -J * static_cast<double>(dot);
      }
    }
    return e_formation;
  }

  // model.hh:293-295
  double per_unitcell() const {
    return per_supercell() / state->configuration.n_unitcells;
  }

  // model.hh:305-345.  Evaluation order ((-J) * ds) * sum, left to right.
  double _single_occ_delta_per_supercell(Index l, int new_occ) const {
    IsingConfiguration const &config = state->configuration;
    std::vector<int> const &occ = config.occupation();
    if (m_use_nlist) {
      int nb = 0;
      for (Index ln : m_flower_nlist[l]) nb += occ[ln];
      return -J * (new_occ - occ[l]) * nb;
    }
    std::vector<int> mi = config.from_linear_site_index(l);
    const int dim = static_cast<int>(mi.size());
    double ds = new_occ - occ[l];
    if (dim == 2) {  // model.hh:315-340 (one multi-index allocation, as there)
      Index rows = config.shape[0];
      int i = mi[0], j = mi[1];
      int nb2 = occ[i + rows * config.within(j - 1, 1)] +
                occ[i + rows * config.within(j + 1, 1)] +
                occ[config.within(i - 1, 0) + rows * j] +
                occ[config.within(i + 1, 0) + rows * j];
      return -J * ds * nb2;
    }
    int nb = 0;
    for (int d = dim - 1; d >= 0; --d) {
      std::vector<int> m = mi, p = mi;
      m[d] = static_cast<int>(config.within(mi[d] - 1, d));
      p[d] = static_cast<int>(config.within(mi[d] + 1, d));
      nb += occ[config.to_linear_site_index(m)];
      nb += occ[config.to_linear_site_index(p)];
    }
    return -J * ds * nb;
  }

  // model.hh:354-379
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) const {
    auto &config = const_cast<IsingConfiguration &>(state->configuration);
    if (linear_site_index.size() == 1) {
      return _single_occ_delta_per_supercell(linear_site_index[0], new_occ[0]);
    }
    double dE = 0.0;
    m_original_value.clear();
    for (size_t i = 0; i < linear_site_index.size(); ++i) {
      Index index = linear_site_index[i];
      dE += _single_occ_delta_per_supercell(index, new_occ[i]);
      m_original_value.push_back(config.occ(index));
      config.set_occ(index, new_occ[i]);
    }
    for (size_t i = 0; i < m_original_value.size(); ++i)
      config.set_occ(linear_site_index[i], m_original_value[i]);
    return dE;
  }

 private:
  mutable std::vector<int> m_original_value;
  bool m_use_nlist = true;
  std::vector<std::vector<Index>> m_nlist;
  std::vector<std::vector<Index>> m_flower_nlist;
};

// ---------------------------------------------------------------------------
// IsingParamComposition (model.hh:388-436)
// ---------------------------------------------------------------------------
class IsingParamComposition {
 public:
  typedef IsingState state_type;
  explicit IsingParamComposition(state_type const * /*_state*/ = nullptr)
      : state(nullptr) {}
  state_type const *state;

  void set_state(state_type const *_state) {
    state = throw_if_null(
        _state, "Error in IsingParamComposition::set_state: _state==nullptr");
  }
  Index n_independent_compositions() const { return 1; }

  // model.hh:412-417 (int32 sum() in the reference; int64 here)
  std::vector<double> per_supercell() const {
    std::vector<int> const &occ = state->configuration.occupation();
    long sum = 0;
    for (int v : occ) sum += v;
    std::vector<double> r(1);
    r[0] = static_cast<double>(static_cast<long>(occ.size()) + sum) / 2.0;
    return r;
  }
  // model.hh:420-422
  std::vector<double> per_unitcell() const {
    std::vector<double> r = per_supercell();
    r[0] = r[0] / static_cast<double>(state->configuration.n_unitcells);
    return r;
  }
  // model.hh:425-435 (allocates a size-1 vector per call, as the reference)
  std::vector<double> occ_delta_per_supercell(
      std::vector<Index> const &linear_site_index,
      std::vector<int> const &new_occ) const {
    auto const &config = state->configuration;
    std::vector<double> Ndx(1);
    Ndx[0] = 0.0;
    for (size_t i = 0; i < linear_site_index.size(); ++i)
      Ndx[0] += (new_occ[i] - config.occ(linear_site_index[i])) / 2.0;
    return Ndx;
  }
};

// model.hh:439-452
class IsingSystem {
 public:
  typedef IsingState state_type;
  typedef IsingFormationEnergy formation_energy_f_type;
  typedef IsingParamComposition param_composition_f_type;
  IsingSystem(formation_energy_f_type f, param_composition_f_type c)
      : formation_energy_calculator(std::move(f)),
        param_composition_calculator(std::move(c)) {}
  formation_energy_f_type formation_energy_calculator;
  param_composition_f_type param_composition_calculator;
};

// ---------------------------------------------------------------------------
// SemiGrandCanonicalConditions
// (include/casm/monte/ising_cpp/basic_semigrand_canonical.hh:36-80)
// ---------------------------------------------------------------------------
class SemiGrandCanonicalConditions {
 public:
  SemiGrandCanonicalConditions() : temperature(0.0) {}
  SemiGrandCanonicalConditions(double T, std::vector<double> mu)
      : temperature(T), exchange_potential(std::move(mu)) {}
  double temperature;
  std::vector<double> exchange_potential;

  static SemiGrandCanonicalConditions from_values(ValueMap const &values) {
    if (!values.scalar_values.count("temperature"))
      throw std::runtime_error("Missing required condition: \"temperature\"");
    if (!values.vector_values.count("exchange_potential"))
      throw std::runtime_error(
          "Missing required condition: \"exchange_potential\"");
    return SemiGrandCanonicalConditions(
        values.scalar_values.at("temperature"),
        values.vector_values.at("exchange_potential"));
  }
  ValueMap to_values() const {
    ValueMap v;
    v.scalar_values["temperature"] = temperature;
    v.vector_values["exchange_potential"] = exchange_potential;
    return v;
  }
};

inline double dot(std::vector<double> const &a, std::vector<double> const &b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}

// ---------------------------------------------------------------------------
// SemiGrandCanonicalPotential (basic_semigrand_canonical.hh:105-199)
// ---------------------------------------------------------------------------
class SemiGrandCanonicalPotential {
 public:
  typedef IsingSystem system_type;
  typedef IsingState state_type;

  explicit SemiGrandCanonicalPotential(std::shared_ptr<system_type> _system)
      : system(throw_if_null(_system,
                             "Error constructing SemiGrandCanonicalPotential: "
                             "_system==nullptr")),
        state(nullptr), conditions(nullptr),
        formation_energy_calculator(system->formation_energy_calculator),
        param_composition_calculator(system->param_composition_calculator) {}

  std::shared_ptr<system_type> system;
  state_type const *state;
  std::shared_ptr<SemiGrandCanonicalConditions> conditions;
  IsingFormationEnergy formation_energy_calculator;    // copies (by value)
  IsingParamComposition param_composition_calculator;  // :125-126

  void set_state(state_type const *_state,
                 std::shared_ptr<SemiGrandCanonicalConditions> _conditions) {
    state = throw_if_null(
        _state,
        "Error in SemiGrandCanonicalPotential::set_state: _state is nullptr");
    conditions = throw_if_null(
        _conditions,
        "Error in SemiGrandCanonicalPotential::set_state: "
        "_conditions is nullptr");
    formation_energy_calculator.set_state(_state);
    param_composition_calculator.set_state(_state);
  }
  // :165-169
  double per_supercell() {
    return formation_energy_calculator.per_supercell() -
           dot(conditions->exchange_potential,
               param_composition_calculator.per_supercell());
  }
  // :172-174
  double per_unitcell() {
    return per_supercell() / state->configuration.n_unitcells;
  }
  // :178-192
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) const {
    double dE_f = formation_energy_calculator.occ_delta_per_supercell(
        linear_site_index, new_occ);
    std::vector<double> Ndx =
        param_composition_calculator.occ_delta_per_supercell(linear_site_index,
                                                             new_occ);
    return dE_f - dot(conditions->exchange_potential, Ndx);
  }
  double occ_delta_per_supercell(OccEvent const &e) const {
    return occ_delta_per_supercell(e.linear_site_index, e.new_occ);
  }
};

// ---------------------------------------------------------------------------
// metropolis_acceptance (include/casm/monte/methods/metropolis.hh:26-35)
// ---------------------------------------------------------------------------
template <typename GeneratorType>
bool metropolis_acceptance(double delta_potential_energy, double beta,
                           GeneratorType &rng) {
  if (delta_potential_energy < 0.0) return true;
  double rand = rng.random_real(1.0);
  double prob = std::exp(-delta_potential_energy * beta);
  return rand < prob;
}

// ---------------------------------------------------------------------------
// Sampler (include/casm/monte/sampling/Sampler.hh:30-112, 233-330)
// Row-per-sample, column-major matrix with capacity growth.
// ---------------------------------------------------------------------------
inline std::vector<std::string> colmajor_component_names(Index n_rows,
                                                         Index n_cols) {
  std::vector<std::string> r;
  for (Index c = 0; c < n_cols; ++c)
    for (Index w = 0; w < n_rows; ++w)
      r.push_back(std::to_string(w) + "," + std::to_string(c));
  return r;
}
// Sampler.hh:333-352
inline std::vector<std::string> default_component_names(
    std::vector<Index> const &shape) {
  if (shape.empty()) return {"0"};
  if (shape.size() == 1) {
    std::vector<std::string> r;
    for (Index i = 0; i < shape[0]; ++i) r.push_back(std::to_string(i));
    return r;
  }
  if (shape.size() == 2) return colmajor_component_names(shape[0], shape[1]);
  throw std::runtime_error(
      "Error constructing sampler component names: >2 dimensions is not "
      "supported");
}

class Sampler {
 public:
  explicit Sampler(std::vector<Index> _shape, CountType _capacity_increment = 1000)
      : m_component_names(default_component_names(_shape)), m_shape(_shape),
        m_n_samples(0), m_capacity_increment(_capacity_increment) {
    m_n_components = 1;
    for (Index x : _shape) m_n_components *= x;
    clear();
  }
  Sampler(std::vector<Index> _shape, std::vector<std::string> const &names,
          CountType _capacity_increment = 1000)
      : m_n_components(static_cast<Index>(names.size())),
        m_component_names(names), m_shape(_shape), m_n_samples(0),
        m_capacity_increment(_capacity_increment) {
    clear();
  }

  void push_back(double value) {
    grow_if_full();
    m_cols[0][m_n_samples] = value;
    ++m_n_samples;
  }
  void push_back(std::vector<double> const &v) {
    if (static_cast<Index>(v.size()) != m_n_components)
      throw std::runtime_error("Error in Sampler::push_back: size mismatch");
    grow_if_full();
    for (Index c = 0; c < m_n_components; ++c) m_cols[c][m_n_samples] = v[c];
    ++m_n_samples;
  }
  // values given as rows (n_samples x n_components, row-major list of rows)
  void set_values(std::vector<std::vector<double>> const &rows) {
    m_capacity = static_cast<CountType>(rows.size());
    m_n_samples = m_capacity;
    for (Index c = 0; c < m_n_components; ++c) {
      m_cols[c].assign(m_capacity, 0.0);
      for (CountType r = 0; r < m_capacity; ++r) m_cols[c][r] = rows[r][c];
    }
  }
  void clear() {
    m_capacity = m_capacity_increment;
    m_cols.assign(m_n_components, std::vector<double>(m_capacity, 0.0));
    m_n_samples = 0;
  }
  void set_sample_capacity(CountType cap) {
    m_capacity = cap;
    for (auto &c : m_cols) c.resize(cap, 0.0);
    if (m_n_samples > cap) m_n_samples = cap;
  }
  void set_capacity_increment(CountType inc) { m_capacity_increment = inc; }
  std::vector<std::string> const &component_names() const {
    return m_component_names;
  }
  std::vector<Index> const &shape() const { return m_shape; }
  Index n_components() const { return m_n_components; }
  CountType n_samples() const { return m_n_samples; }
  CountType sample_capacity() const { return m_capacity; }
  // one component = one contiguous column, first n_samples entries
  std::vector<double> component(Index c) const {
    if (m_n_components == 0 || m_cols.empty()) return {};
    return std::vector<double>(m_cols[c].begin(),
                               m_cols[c].begin() + m_n_samples);
  }
  double const *component_data(Index c) const { return m_cols[c].data(); }
  std::vector<double> sample(CountType r) const {
    std::vector<double> v(m_n_components);
    for (Index c = 0; c < m_n_components; ++c) v[c] = m_cols[c][r];
    return v;
  }

 private:
  void grow_if_full() {
    if (m_n_samples == m_capacity)
      set_sample_capacity(m_capacity + m_capacity_increment);
  }
  Index m_n_components;
  std::vector<std::string> m_component_names;
  std::vector<Index> m_shape;
  Index m_n_samples;
  CountType m_capacity_increment;
  CountType m_capacity = 0;
  std::vector<std::vector<double>> m_cols;
};

typedef std::map<std::string, std::shared_ptr<Sampler>> SamplerMap;

// Sampler.hh:128-149, 365-370
struct SamplerComponent {
  SamplerComponent(std::string s, Index i, std::string n)
      : sampler_name(std::move(s)), component_index(i),
        component_name(std::move(n)) {}
  std::string sampler_name;
  Index component_index = 0;
  std::string component_name;
  bool operator<(SamplerComponent const &o) const {
    if (sampler_name == o.sampler_name)
      return component_index < o.component_index;
    return sampler_name < o.sampler_name;
  }
};

// Sampler.hh:151-180
struct RequestedPrecision {
  bool abs_convergence_is_required = false;
  double abs_precision = 0.0;
  bool rel_convergence_is_required = false;
  double rel_precision = 0.0;
  static RequestedPrecision abs_and_rel(double a, double r) {
    RequestedPrecision x;
    x.abs_convergence_is_required = true;
    x.abs_precision = a;
    x.rel_convergence_is_required = true;
    x.rel_precision = r;
    return x;
  }
  static RequestedPrecision abs(double v) {
    RequestedPrecision x;
    x.abs_convergence_is_required = true;
    x.abs_precision = v;
    return x;
  }
  static RequestedPrecision rel(double v) {
    RequestedPrecision x;
    x.rel_convergence_is_required = true;
    x.rel_precision = v;
    return x;
  }
};
typedef std::map<SamplerComponent, RequestedPrecision> RequestedPrecisionMap;

// Sampler.hh:373-408
inline SamplerMap::const_iterator find_or_throw(SamplerMap const &samplers,
                                                SamplerComponent const &key) {
  auto it = samplers.find(key.sampler_name);
  if (it == samplers.end()) {
    std::stringstream msg;
    msg << "Error finding sampler component: Sampler '" << key.sampler_name
        << "' not found." << std::endl;
    throw std::runtime_error(msg.str());
  }
  if (key.component_index >= it->second->n_components()) {
    std::stringstream msg;
    msg << "Error finding sampler component: Requested component index "
        << key.component_index << ", but '" << key.sampler_name << "' has "
        << it->second->n_components() << "components." << std::endl;
    throw std::runtime_error(msg.str());
  }
  return it;
}
// Sampler.hh:410-416 (reads the FIRST map entry)
inline CountType get_n_samples(SamplerMap const &samplers) {
  if (samplers.size()) return samplers.begin()->second->n_samples();
  return CountType(0);
}

// include/casm/monte/sampling/StateSamplingFunction.hh:20-54
struct StateSamplingFunction {
  StateSamplingFunction(std::string _name, std::string _description,
                        std::vector<Index> _shape,
                        std::function<std::vector<double>()> _function)
      : name(std::move(_name)), description(std::move(_description)),
        shape(std::move(_shape)),
        component_names(default_component_names(shape)),
        function(std::move(_function)) {}
  std::string name;
  std::string description;
  std::vector<Index> shape;
  std::vector<std::string> component_names;
  std::function<std::vector<double>()> function;
  std::vector<double> operator()() const { return function(); }
};
typedef std::map<std::string, StateSamplingFunction> StateSamplingFunctionMap;

// ---------------------------------------------------------------------------
// misc/math.hh:21-72   (views are (pointer, length))
// ---------------------------------------------------------------------------
inline double mean_of(double const *x, Index n) {
  // Eigen's .mean() is a vectorised tree reduction; a sequential sum is used
  // here, so the last bits may differ from a given Eigen build (unpinned).
  double s = 0.0;
  for (Index i = 0; i < n; ++i) s += x[i];
  return s / static_cast<double>(n);
}
inline double covariance(double const *x, double const *y, Index n,
                         double mean) {
  double cov = 0.0;
  for (Index i = 0; i < n; ++i) cov += (x[i] - mean) * (y[i] - mean);
  return cov / n;
}
inline double variance(double const *x, Index n, double x_mean) {
  double cov = 0.0;
  for (Index i = 0; i < n; ++i) {
    double d = x[i] - x_mean;
    cov += d * d;
  }
  return cov / n;
}
inline double weighted_variance(double const *x, Index n, double x_mean,
                                double const *w, double w_sum) {
  double cov = 0.0;
  for (Index i = 0; i < n; ++i) {
    double d = x[i] - x_mean;
    cov += w[i] * d * d;
  }
  return cov / w_sum;
}
// Winitzki's approximation, a = 0.147
inline double approx_erf_inv(double x) {
  const double one = 1.0;
  const double PI = 3.141592653589793238463;
  const double a = 0.147;
  double sgn = (x < 0.0) ? -one : one;
  double b = std::log((one - x) * (one + x));
  double c = 2.0 / (PI * a) + b * 0.5;
  double d = b / a;
  return sgn * std::sqrt(std::sqrt(c * c - d) - c);
}

// ---------------------------------------------------------------------------
// BasicStatistics (include/casm/monte/BasicStatistics.hh:24-79,
//                  src/casm/monte/BasicStatistics.cc:24-188)
// ---------------------------------------------------------------------------
struct BasicStatistics {
  double mean = 0.0;
  double calculated_precision = std::numeric_limits<double>::max();
};
inline double get_calculated_precision(BasicStatistics const &s) {
  return s.calculated_precision;
}
inline double get_calculated_relative_precision(BasicStatistics const &s) {
  return std::abs(s.calculated_precision / s.mean);
}

// BasicStatistics.cc:24-48
inline double autocorrelation_factor(double const *obs, Index N,
                                     double increment = 1.0,
                                     Index *k_star = nullptr) {
  double mean = mean_of(obs, N);
  double CoVar0 = variance(obs, N, mean);
  if (k_star) *k_star = 0;
  if (std::abs(CoVar0 / mean) < 1e-8 || CoVar0 == 0.0) return 1.0;
  for (CountType i = 1; i < N; ++i) {
    CountType range = N - i;
    double cov = covariance(obs, obs + i, range, mean);
    if (std::abs(cov / CoVar0) <= 0.5) {
      double rho = std::pow(2.0, (-1.0 / (i * increment)));
      if (k_star) *k_star = i;
      return (1.0 + rho) / (1.0 - rho);
    }
  }
  if (k_star) *k_star = -1;
  return std::numeric_limits<double>::max();
}

// BasicStatistics.cc:50-73
inline std::vector<double> resample(std::vector<double> const &obs,
                                    std::vector<double> const &weight,
                                    double weight_sum, Index n_equally_spaced) {
  double increment = weight_sum / n_equally_spaced;
  std::vector<double> out(n_equally_spaced);
  Index j = 0;
  double W_j = 0.0;
  for (Index i = 0; i < n_equally_spaced; ++i) {
    double W_target = i * increment;
    while (W_j + weight[j] < W_target) {
      W_j += weight[j];
      ++j;
    }
    out[i] = obs[j];
  }
  return out;
}

struct BasicStatisticsCalculator {
  explicit BasicStatisticsCalculator(double _confidence = 0.95, Index _method = 1,
                                     Index _n_resamples = 10000)
      : confidence(_confidence), method(_method), n_resamples(_n_resamples) {}
  double confidence;
  Index method;
  Index n_resamples;

  // BasicStatistics.cc:114-131
  BasicStatistics operator()(std::vector<double> const &obs) const {
    if (obs.empty())
      throw std::runtime_error(
          "Error in BasicStatisticsCalculator: observations.size()==0");
    CountType N = static_cast<CountType>(obs.size());
    BasicStatistics stats;
    stats.mean = mean_of(obs.data(), N);
    double CoVar0 = variance(obs.data(), N, stats.mean);
    double f_autocorr = autocorrelation_factor(obs.data(), N);
    double f_confidence = std::sqrt(2.0) * approx_erf_inv(confidence);
    stats.calculated_precision =
        f_confidence * std::sqrt(f_autocorr * CoVar0 / N);
    return stats;
  }
  // BasicStatistics.cc:144-188
  BasicStatistics operator()(std::vector<double> const &obs,
                             std::vector<double> const &weight) const {
    if (obs.empty())
      throw std::runtime_error(
          "Error in BasicStatisticsCalculator: observations.size()==0");
    if (weight.empty()) return (*this)(obs);
    if (obs.size() != weight.size())
      throw std::runtime_error(
          "Error in BasicStatisticsCalculator: observations.size() != "
          "sample_weight.size()");
    double W = 0.0;
    for (double w : weight) W += w;
    double increment = W / n_resamples;
    std::vector<double> eq = resample(obs, weight, W, n_resamples);
    if (method == 1) {
      BasicStatistics stats;
      double d = 0.0;
      for (size_t i = 0; i < obs.size(); ++i) d += obs[i] * weight[i];
      stats.mean = d / W;
      double wvar = weighted_variance(obs.data(), obs.size(), stats.mean,
                                      weight.data(), W);
      double f_autocorr =
          autocorrelation_factor(eq.data(), eq.size(), increment);
      double f_confidence = std::sqrt(2.0) * approx_erf_inv(confidence);
      stats.calculated_precision =
          f_confidence * std::sqrt(f_autocorr * wvar / W);
      return stats;
    } else if (method == 2) {
      return (*this)(eq);
    }
    throw std::runtime_error(
        "Error in BasicStatisticsCalculator: invalid method");
  }
};

typedef std::function<BasicStatistics(std::vector<double> const &,
                                      std::vector<double> const &)>
    CalcStatisticsFunction;

// ---------------------------------------------------------------------------
// Equilibration check (src/casm/monte/checks/EquilibrationCheck.cc:50-225,
//                      include/casm/monte/checks/EquilibrationCheck.hh)
// ---------------------------------------------------------------------------
struct IndividualEquilibrationCheckResult {
  bool is_equilibrated = false;
  CountType N_samples_for_equilibration = 0;
};

// EquilibrationCheck.cc:50-117
inline IndividualEquilibrationCheckResult _default_equilibration_check(
    std::vector<double> const &x, double prec) {
  if (x.empty())
    throw std::runtime_error(
        "Error in equilibration_check: observations.size()==0");
  IndividualEquilibrationCheckResult result;
  CountType N = static_cast<CountType>(x.size());
  double eps = (x[0] == 0.0) ? 1e-8 : std::abs(x[0]) * 1e-8;
  bool is_even = ((N % 2) == 0);

  bool all_same = true;
  for (CountType i = 0; i < N; ++i)
    if (std::abs(x[i] - x[0]) > eps) {
      all_same = false;
      break;
    }
  if (all_same) {
    result.is_equilibrated = true;
    result.N_samples_for_equilibration = 0;
    return result;
  }

  CountType start1 = 0;
  CountType start2 = is_even ? N / 2 : (N / 2) + 1;
  double sum1 = 0.0, sum2 = 0.0;
  for (CountType i = 0; i < start2; ++i) sum1 += x[i];
  for (CountType i = start2; i < N; ++i) sum2 += x[i];

  while (std::abs((sum1 / (start2 - start1)) - (sum2 / (N - start2))) > prec &&
         start1 < N - 2) {
    if (is_even) {
      sum1 -= x[start1];
      sum1 += x[start2];
      sum2 -= x[start2];
      start2++;
    } else {
      sum1 -= x[start1];
    }
    start1++;
    is_even = !is_even;
  }

  double mean_tot = (sum1 + sum2) / (N - start1);
  if (x[start1] < mean_tot) {
    while (x[start1] < mean_tot && start1 < N - 1) start1++;
  } else {
    while (x[start1] > mean_tot && start1 < N - 1) start1++;
  }
  result.is_equilibrated = (start1 < N - 1);
  result.N_samples_for_equilibration = start1;
  return result;
}

// EquilibrationCheck.cc:119-162
inline IndividualEquilibrationCheckResult default_equilibration_check(
    std::vector<double> const &obs, std::vector<double> const &weight,
    RequestedPrecision req) {
  double prec;
  if (req.abs_convergence_is_required) {
    prec = req.abs_precision;
  } else if (req.rel_convergence_is_required) {
    prec = std::abs(mean_of(obs.data(), obs.size()) * req.rel_precision);
  } else {
    IndividualEquilibrationCheckResult r;
    r.is_equilibrated = true;
    r.N_samples_for_equilibration = 0;
    return r;
  }
  if (weight.empty()) return _default_equilibration_check(obs, prec);
  if (weight.size() != obs.size())
    throw std::runtime_error(
        "Error in equilibration_check: sample_weight.size() != "
        "observations.size()");
  Index N = static_cast<Index>(weight.size());
  double W = 0.0;
  for (double w : weight) W += w;
  double weight_factor = N / W;
  std::vector<double> wobs = obs;
  for (size_t i = 0; i < wobs.size(); ++i) wobs[i] *= weight_factor * weight[i];
  return _default_equilibration_check(wobs, prec);
}

typedef std::function<IndividualEquilibrationCheckResult(
    std::vector<double> const &, std::vector<double> const &,
    RequestedPrecision)>
    EquilibrationCheckFunction;

struct EquilibrationCheckResults {
  bool all_equilibrated = false;
  CountType N_samples_for_all_to_equilibrate = 0;
  std::map<SamplerComponent, IndividualEquilibrationCheckResult>
      individual_results;
};

// EquilibrationCheck.cc:180-225
inline EquilibrationCheckResults equilibration_check(
    EquilibrationCheckFunction f, RequestedPrecisionMap const &requested,
    SamplerMap const &samplers, Sampler const &sample_weight, bool check_all) {
  if (f == nullptr)
    throw std::runtime_error(
        "Error in equilibration_check: equilibration_check_f == nullptr");
  EquilibrationCheckResults results;
  if (!requested.size()) return results;
  results.all_equilibrated = true;
  for (auto const &p : requested) {
    SamplerComponent const &key = p.first;
    Sampler const &sampler = *find_or_throw(samplers, key)->second;
    IndividualEquilibrationCheckResult cur =
        f(sampler.component(key.component_index), sample_weight.component(0),
          p.second);
    results.N_samples_for_all_to_equilibrate =
        std::max(results.N_samples_for_all_to_equilibrate,
                 cur.N_samples_for_equilibration);
    results.all_equilibrated &= cur.is_equilibrated;
    results.individual_results.emplace(key, cur);
    if (!check_all && !results.all_equilibrated) break;
  }
  return results;
}

// ---------------------------------------------------------------------------
// Convergence check (include/casm/monte/checks/ConvergenceCheck.hh:13-184)
// ---------------------------------------------------------------------------
struct IndividualConvergenceCheckResult {
  bool is_converged = false;
  RequestedPrecision requested_precision;
  BasicStatistics stats;
};
struct ConvergenceCheckResults {
  bool all_converged = false;
  CountType N_samples_for_statistics = 0;
  std::map<SamplerComponent, IndividualConvergenceCheckResult>
      individual_results;
};

// ConvergenceCheck.hh:74-90
inline IndividualConvergenceCheckResult convergence_check(
    BasicStatistics const &stats, RequestedPrecision const &req) {
  IndividualConvergenceCheckResult r;
  r.stats = stats;
  r.requested_precision = req;
  r.is_converged = true;
  if (req.abs_convergence_is_required)
    r.is_converged &= get_calculated_precision(stats) < req.abs_precision;
  if (req.rel_convergence_is_required)
    r.is_converged &=
        get_calculated_relative_precision(stats) < req.rel_precision;
  return r;
}

inline std::vector<double> tail_of(std::vector<double> const &v, CountType n) {
  return std::vector<double>(v.end() - n, v.end());
}

// ConvergenceCheck.hh:94-119
inline IndividualConvergenceCheckResult component_convergence_check(
    Sampler const &sampler, Sampler const &sample_weight,
    SamplerComponent const &key, RequestedPrecision const &req,
    CountType N_stats, CalcStatisticsFunction calc_statistics_f) {
  if (calc_statistics_f == nullptr)
    throw std::runtime_error(
        "Error in component_convergence_check: calc_statistics_f == nullptr");
  if (sample_weight.n_samples() != 0) {
    return convergence_check(
        calc_statistics_f(tail_of(sampler.component(key.component_index), N_stats),
                          tail_of(sample_weight.component(0), N_stats)),
        req);
  }
  static const std::vector<double> empty_weight;
  return convergence_check(
      calc_statistics_f(tail_of(sampler.component(key.component_index), N_stats),
                        empty_weight),
      req);
}

// ConvergenceCheck.hh:139-184
inline ConvergenceCheckResults convergence_check(
    SamplerMap const &samplers, Sampler const &sample_weight,
    RequestedPrecisionMap const &requested, CountType N_equil,
    CalcStatisticsFunction calc_statistics_f) {
  ConvergenceCheckResults results;
  CountType N_samples = get_n_samples(samplers);
  if (!requested.size()) {
    results.N_samples_for_statistics = N_samples;
    return results;
  }
  if (N_equil >= N_samples) return results;
  results.N_samples_for_statistics = N_samples - N_equil;
  results.all_converged = true;
  for (auto const &p : requested) {
    SamplerComponent const &key = p.first;
    Sampler const &sampler = *find_or_throw(samplers, key)->second;
    IndividualConvergenceCheckResult cur = component_convergence_check(
        sampler, sample_weight, key, p.second,
        results.N_samples_for_statistics, calc_statistics_f);
    results.all_converged &= cur.is_converged;
    results.individual_results.emplace(key, cur);
  }
  return results;
}

// ---------------------------------------------------------------------------
// Cutoff check (include/casm/monte/checks/CutoffCheck.hh:14-94)
// ---------------------------------------------------------------------------
struct CutoffCheckParams {
  std::optional<CountType> min_count;
  std::optional<TimeType> min_time;
  std::optional<CountType> min_sample;
  std::optional<TimeType> min_clocktime;
  std::optional<CountType> max_count;
  std::optional<TimeType> max_time;
  std::optional<CountType> max_sample;
  std::optional<TimeType> max_clocktime;
};

inline bool all_minimums_met(CutoffCheckParams const &p,
                             std::optional<CountType> count,
                             std::optional<TimeType> time, CountType n_samples,
                             TimeType clocktime) {
  if (p.min_sample.has_value() && n_samples < p.min_sample.value()) return false;
  if (p.min_count.has_value() && count.has_value() &&
      count.value() < p.min_count.value())
    return false;
  if (p.min_time.has_value() && time.has_value() &&
      time.value() < p.min_time.value())
    return false;
  if (p.min_clocktime.has_value() && clocktime < p.min_clocktime.value())
    return false;
  return true;
}
inline bool any_maximum_met(CutoffCheckParams const &p,
                            std::optional<CountType> count,
                            std::optional<TimeType> time, CountType n_samples,
                            TimeType clocktime) {
  if (p.max_sample.has_value() && n_samples >= p.max_sample.value()) return true;
  if (p.max_count.has_value() && count.has_value() &&
      count.value() >= p.max_count.value())
    return true;
  if (p.max_time.has_value() && time.has_value() &&
      time.value() >= p.max_time.value())
    return true;
  if (p.max_clocktime.has_value() && clocktime >= p.max_clocktime.value())
    return true;
  return false;
}

// ---------------------------------------------------------------------------
// Wall clock standing in for CASM::Log's timer (casm/casm_io/Log.hh, external)
// and MethodLog (include/casm/monte/MethodLog.hh:13-38).
// ---------------------------------------------------------------------------
struct Clock {
  using clk = std::chrono::steady_clock;
  clk::time_point t0 = clk::now();
  clk::time_point lap0 = clk::now();
  void restart_clock() { t0 = clk::now(); }
  void begin_lap() { lap0 = clk::now(); }
  double time_s() const {
    return std::chrono::duration<double>(clk::now() - t0).count();
  }
  double lap_time() const {
    return std::chrono::duration<double>(clk::now() - lap0).count();
  }
};
struct MethodLog {
  std::string logfile_path;
  Clock log;
  std::optional<double> log_frequency;
};

// ---------------------------------------------------------------------------
// Completion check (include/casm/monte/checks/CompletionCheck.hh:20-376)
// ---------------------------------------------------------------------------
struct CompletionCheckParams {
  CompletionCheckParams()
      : equilibration_check_f(default_equilibration_check),
        calc_statistics_f(BasicStatisticsCalculator()) {}
  CutoffCheckParams cutoff_params;
  EquilibrationCheckFunction equilibration_check_f;
  CalcStatisticsFunction calc_statistics_f;
  RequestedPrecisionMap requested_precision;
  bool log_spacing = false;
  CountType check_begin = 100;
  CountType check_period = 100;
  double check_base = 10.0;
  double check_shift = 2.0;
  CountType check_period_max = 10000;

  CountType sample_check_linear(Index n) const {
    return check_begin + check_period * n;
  }
  CountType sample_check_log(Index n) const {
    return check_begin + static_cast<CountType>(std::round(
                             std::pow(check_base, (n + check_shift))));
  }
  Index find_n_begin_linear() const {
    Index n_begin_linear = 0;
    auto check_delta = [&](Index n) {
      return sample_check_log(n) - sample_check_log(n - 1);
    };
    while (check_delta(n_begin_linear + 1) <= check_period_max)
      n_begin_linear += 1;
    return n_begin_linear;
  }
  CountType sample_check_log(Index n, Index n_begin_linear) const {
    if (n <= n_begin_linear) return sample_check_log(n);
    return sample_check_log(n_begin_linear) +
           check_period_max * (n - n_begin_linear);
  }
};

struct CompletionCheckResults {
  CompletionCheckParams params;
  std::optional<CountType> count;
  std::optional<TimeType> time;
  TimeType clocktime = 0.0;
  CountType n_samples = 0;
  bool has_all_minimums_met = false;
  bool has_any_maximum_met = false;
  std::optional<CountType> n_samples_at_convergence_check;
  EquilibrationCheckResults equilibration_check_results;
  ConvergenceCheckResults convergence_check_results;
  bool is_complete = false;

  void partial_reset(std::optional<CountType> _count = std::nullopt,
                     std::optional<TimeType> _time = std::nullopt,
                     TimeType _clocktime = 0.0, CountType _n_samples = 0) {
    count = _count;
    time = _time;
    clocktime = _clocktime;
    n_samples = _n_samples;
    has_all_minimums_met = false;
    has_any_maximum_met = false;
    is_complete = false;
  }
  void full_reset(std::optional<CountType> _count = std::nullopt,
                  std::optional<TimeType> _time = std::nullopt,
                  TimeType _clocktime = 0.0, CountType _n_samples = 0) {
    partial_reset(_count, _time, _clocktime, _n_samples);
    n_samples_at_convergence_check = std::nullopt;
    equilibration_check_results = EquilibrationCheckResults();
    convergence_check_results = ConvergenceCheckResults();
  }
};

class CompletionCheck {
 public:
  explicit CompletionCheck(CompletionCheckParams params)
      : m_params(std::move(params)),
        m_n_begin_linear(m_params.find_n_begin_linear()) {
    m_results.params = m_params;
    m_results.is_complete = false;
    if (m_params.equilibration_check_f == nullptr)
      throw std::runtime_error(
          "Error constructing CompletionCheck: params.equilibration_check_f == "
          "nullptr");
    if (m_params.calc_statistics_f == nullptr)
      throw std::runtime_error(
          "Error constructing CompletionCheck: params.calc_statistics_f == "
          "nullptr");
  }
  CompletionCheckParams const &params() const { return m_params; }
  void reset() {
    m_results.full_reset();
    m_n_checks = 0;
    m_last_n_samples = 0;
    m_last_clocktime = 0.0;
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, Clock &log) {
    return _is_complete(s, w, std::nullopt, std::nullopt, log);
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, CountType count,
                   Clock &log) {
    return _is_complete(s, w, count, std::nullopt, log);
  }
  bool is_complete_time(SamplerMap const &s, Sampler const &w, TimeType time,
                        Clock &log) {
    return _is_complete(s, w, std::nullopt, time, log);
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, CountType count,
                   TimeType time, Clock &log) {
    return _is_complete(s, w, count, time, log);
  }
  CompletionCheckResults const &results() const { return m_results; }
  Index n_checks() const { return m_n_checks; }

 private:
  // CompletionCheck.hh:273-328
  bool _is_complete(SamplerMap const &samplers, Sampler const &sample_weight,
                    std::optional<CountType> count,
                    std::optional<TimeType> time, Clock &log) {
    CountType n_samples = get_n_samples(samplers);
    TimeType clocktime = m_last_clocktime;
    if (n_samples != m_last_n_samples) {
      clocktime = log.time_s();
      m_last_n_samples = n_samples;
      m_last_clocktime = clocktime;
    }
    m_results.partial_reset(count, time, clocktime, n_samples);
    m_results.has_all_minimums_met = all_minimums_met(
        m_params.cutoff_params, count, time, n_samples, clocktime);
    if (!m_results.has_all_minimums_met) return false;

    m_results.has_any_maximum_met = any_maximum_met(
        m_params.cutoff_params, count, time, n_samples, clocktime);
    if (m_results.has_any_maximum_met) {
      m_results.is_complete = true;
      if (!(m_results.n_samples_at_convergence_check.has_value() &&
            n_samples == m_results.n_samples_at_convergence_check.value())) {
        _check_convergence(samplers, sample_weight, n_samples);
      }
      return true;
    }
    Index check_at;
    if (m_params.log_spacing)
      check_at = m_params.sample_check_log(m_n_checks, m_n_begin_linear);
    else
      check_at = m_params.sample_check_linear(m_n_checks);
    if (n_samples >= check_at) {
      m_n_checks += 1;
      _check_convergence(samplers, sample_weight, n_samples);
    }
    if (m_results.convergence_check_results.all_converged)
      m_results.is_complete = true;
    return m_results.is_complete;
  }
  // CompletionCheck.hh:351-376
  void _check_convergence(SamplerMap const &samplers,
                          Sampler const &sample_weight, CountType n_samples) {
    if (m_params.requested_precision.size()) {
      m_results.n_samples_at_convergence_check = n_samples;
      bool check_all = false;
      m_results.equilibration_check_results = equilibration_check(
          m_params.equilibration_check_f, m_params.requested_precision,
          samplers, sample_weight, check_all);
      if (m_results.equilibration_check_results.all_equilibrated) {
        m_results.convergence_check_results = convergence_check(
            samplers, sample_weight, m_params.requested_precision,
            m_results.equilibration_check_results
                .N_samples_for_all_to_equilibrate,
            m_params.calc_statistics_f);
      } else {
        m_results.convergence_check_results = ConvergenceCheckResults();
      }
    }
  }

  CompletionCheckParams m_params;
  CompletionCheckResults m_results;
  Index m_n_checks = 0;
  Index m_n_begin_linear = 0;
  Index m_last_n_samples = 0;
  double m_last_clocktime = 0.0;
};

// ---------------------------------------------------------------------------
// BasicOccupationMetropolisData + main loop
// (include/casm/monte/methods/basic_occupation_metropolis.hh:19-115, 354-425)
// ---------------------------------------------------------------------------
struct BasicOccupationMetropolisData {
  BasicOccupationMetropolisData(StateSamplingFunctionMap const &_fns,
                                CountType _n_steps_per_pass,
                                CompletionCheckParams const &_cc_params)
      : sampling_functions(_fns), sample_weight(std::vector<Index>{}),
        n_steps_per_pass(_n_steps_per_pass), completion_check(_cc_params) {
    for (auto const &kv : sampling_functions) {
      auto const &f = kv.second;
      samplers.emplace(f.name,
                       std::make_shared<Sampler>(f.shape, f.component_names));
    }
    n_pass = 0;
    n_accept = 0;
    n_reject = 0;
  }
  StateSamplingFunctionMap sampling_functions;
  SamplerMap samplers;
  Sampler sample_weight;
  CountType n_pass;
  CountType n_steps_per_pass;
  BigCountType n_accept;
  BigCountType n_reject;
  CompletionCheck completion_check;

  double acceptance_rate() const {
    double a = static_cast<double>(n_accept), r = static_cast<double>(n_reject);
    return a / (a + r);
  }
  double rejection_rate() const {
    double a = static_cast<double>(n_accept), r = static_cast<double>(n_reject);
    return r / (a + r);
  }
  void reset() {
    for (auto &kv : samplers) kv.second->clear();
    sample_weight.clear();
    n_pass = 0;
    n_accept = 0;
    n_reject = 0;
    completion_check.reset();
  }
};

// basic_occupation_metropolis.hh:354-425.  JSON samplers and the status
// writer are callbacks outside the arithmetic; a no-op writer is the default.
template <typename EngineType, typename DPotentialF, typename ProposeF,
          typename ApplyF, typename WriteStatusF>
void basic_occupation_metropolis(BasicOccupationMetropolisData &data,
                                 double temperature, DPotentialF dpotential_f,
                                 ProposeF propose_event_f, ApplyF apply_event_f,
                                 int sample_period,
                                 std::optional<MethodLog> method_log,
                                 std::shared_ptr<EngineType> random_engine,
                                 WriteStatusF write_status_f) {
  double beta = 1.0 / (KB * temperature);
  RandomNumberGenerator<EngineType> rng(random_engine);
  if (!method_log.has_value()) {
    method_log = MethodLog();
    method_log->logfile_path = "status.json";
    method_log->log_frequency = 600.0;
  }
  method_log->log.restart_clock();
  method_log->log.begin_lap();

  Index n_pass_next_sample = sample_period;
  CountType n_step = 0;
  double delta_potential_energy;

  while (!data.completion_check.is_complete(data.samplers, data.sample_weight,
                                            data.n_pass, method_log->log)) {
    OccEvent const &event = propose_event_f(rng);
    delta_potential_energy = dpotential_f(event);
    if (metropolis_acceptance(delta_potential_energy, beta, rng)) {
      data.n_accept++;
      apply_event_f(event);
    } else {
      data.n_reject++;
    }
    n_step++;
    if (n_step == data.n_steps_per_pass) {
      n_step = 0;
      data.n_pass += 1;
    }
    if (data.n_pass == n_pass_next_sample) {
      n_pass_next_sample += sample_period;
      for (auto const &kv : data.sampling_functions) {
        auto const &f = kv.second;
        data.samplers.at(f.name)->push_back(f());
      }
      if (method_log->log_frequency.has_value() &&
          method_log->log.lap_time() >= method_log->log_frequency.value()) {
        write_status_f(data, *method_log);
      }
    }
  }
  write_status_f(data, *method_log);
}

// ---------------------------------------------------------------------------
// Event generator (basic_semigrand_canonical.hh:268-321)
// ---------------------------------------------------------------------------
template <typename EngineType = default_engine_type>
class SemiGrandCanonicalEventGenerator {
 public:
  typedef IsingState state_type;
  typedef EngineType engine_type;
  typedef RandomNumberGenerator<engine_type> random_number_generator_type;

  SemiGrandCanonicalEventGenerator() : state(nullptr), m_max_l(0) {
    occ_event.linear_site_index.assign(1, 0);
    occ_event.new_occ.assign(1, 1);
  }
  state_type *state;
  OccEvent occ_event;

  void set_state(state_type *_state) {
    state = throw_if_null(_state,
                          "Error in SemiGrandCanonicalEventGenerator::set_state: "
                          "_state==nullptr");
    m_max_l = state->configuration.n_sites - 1;
  }
  OccEvent const &propose(random_number_generator_type &rng) {
    occ_event.linear_site_index[0] = rng.random_int(m_max_l);
    occ_event.new_occ[0] =
        -state->configuration.occ(occ_event.linear_site_index[0]);
    return occ_event;
  }
  void apply(OccEvent const &e) {
    state->configuration.set_occ(e.linear_site_index[0], e.new_occ[0]);
  }

 private:
  Index m_max_l;
};

// ---------------------------------------------------------------------------
// SemiGrandCanonicalCalculator (basic_semigrand_canonical.hh:324-470) and the
// default sampling functions (:486-590)
// ---------------------------------------------------------------------------
class SemiGrandCanonicalCalculator {
 public:
  typedef IsingSystem system_type;
  typedef IsingState state_type;
  typedef SemiGrandCanonicalEventGenerator<default_engine_type>
      event_generator_type;
  typedef default_engine_type engine_type;

  explicit SemiGrandCanonicalCalculator(std::shared_ptr<system_type> _system)
      : system(throw_if_null(_system,
                             "Error constructing SemiGrandCanonicalCalculator: "
                             "_system==nullptr")),
        state(nullptr), conditions(nullptr), potential(_system),
        formation_energy_calculator(&potential.formation_energy_calculator),
        param_composition_calculator(&potential.param_composition_calculator) {}

  std::shared_ptr<system_type> system;
  state_type *state;
  std::shared_ptr<SemiGrandCanonicalConditions> conditions;
  SemiGrandCanonicalPotential potential;
  IsingFormationEnergy *formation_energy_calculator;
  IsingParamComposition *param_composition_calculator;
  std::shared_ptr<BasicOccupationMetropolisData> data;

  template <typename WriteStatusF>
  std::shared_ptr<BasicOccupationMetropolisData> run(
      state_type &_state, StateSamplingFunctionMap const &sampling_functions,
      CompletionCheckParams const &completion_check_params,
      event_generator_type event_generator, int sample_period,
      std::optional<MethodLog> method_log,
      std::shared_ptr<engine_type> random_engine, WriteStatusF write_status_f) {
    state = &_state;
    conditions = std::make_shared<SemiGrandCanonicalConditions>(
        SemiGrandCanonicalConditions::from_values(state->conditions));
    double temperature = conditions->temperature;
    CountType n_steps_per_pass = state->configuration.n_variable_sites;
    potential.set_state(state, conditions);
    auto dpotential_f = [this](OccEvent const &e) {
      return potential.occ_delta_per_supercell(e);
    };
    event_generator.set_state(state);
    auto propose_f =
        [&](event_generator_type::random_number_generator_type &rng)
        -> OccEvent const & { return event_generator.propose(rng); };
    auto apply_f = [&](OccEvent const &e) { event_generator.apply(e); };
    data = std::make_shared<BasicOccupationMetropolisData>(
        sampling_functions, n_steps_per_pass, completion_check_params);
    basic_occupation_metropolis(*data, temperature, dpotential_f, propose_f,
                                apply_f, sample_period, method_log,
                                random_engine, write_status_f);
    return data;
  }
};

inline StateSamplingFunction make_parametric_composition_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc) {
  if (mc == nullptr)
    throw std::runtime_error(
        "Error in parametric_composition sampling function: "
        "mc_calculator == nullptr");
  std::vector<Index> shape;
  shape.push_back(
      mc->system->param_composition_calculator.n_independent_compositions());
  auto f = [mc]() -> std::vector<double> {
    if (mc->param_composition_calculator->state == nullptr)
      throw std::runtime_error(
          "Error in parametric_composition sampling function: "
          "mc_calculator->param_composition_calculator->state == nullptr");
    return mc->param_composition_calculator->per_unitcell();
  };
  return StateSamplingFunction("param_composition", "Parametric composition",
                               shape, f);
}
inline StateSamplingFunction make_formation_energy_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc) {
  auto f = [mc]() -> std::vector<double> {
    if (mc->formation_energy_calculator->state == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: "
          "mc_calculator->formation_energy_calculator->state == nullptr");
    return std::vector<double>{mc->formation_energy_calculator->per_unitcell()};
  };
  return StateSamplingFunction("formation_energy", "Intensive formation energy",
                               {}, f);
}
inline StateSamplingFunction make_potential_energy_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc) {
  auto f = [mc]() -> std::vector<double> {
    if (mc->potential.state == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: "
          "mc_calculator->potential.state == nullptr");
    return std::vector<double>{mc->potential.per_unitcell()};
  };
  return StateSamplingFunction("potential_energy", "Intensive potential energy",
                               {}, f);
}

// ===========================================================================
// Everything below is NOT in the reference: it states, on the CPU, the
// production (checkerboard) update order that the GPU path uses, so the CUDA
// kernels can be checked bit-for-bit against a scalar restatement.  The
// per-site physics (dE, acceptance rule) is the reference's; only the visiting
// order and the random stream differ.
// ===========================================================================

// Philox4x32-10 (Salmon et al., SC'11), the published algorithm.
struct Philox4x32 {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  static std::array<uint32_t, 4> generate(std::array<uint32_t, 4> c,
                                          std::array<uint32_t, 2> k,
                                          int rounds = 10) {
    for (int r = 0; r < rounds; ++r) {
      uint64_t p0 = static_cast<uint64_t>(M0) * c[0];
      uint64_t p1 = static_cast<uint64_t>(M1) * c[2];
      std::array<uint32_t, 4> n;
      n[0] = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k[0];
      n[1] = static_cast<uint32_t>(p1);
      n[2] = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k[1];
      n[3] = static_cast<uint32_t>(p0);
      c = n;
      k[0] += W0;
      k[1] += W1;
    }
    return c;
  }
};

// dE and acceptance tables shared by both modes: index [s>0][n_up] where n_up
// is the number of +1 neighbours (0..2*dim).  Built with the reference's exact
// expression order (model.hh:312-314, :430-433;
// basic_semigrand_canonical.hh:185-191; metropolis.hh:28-34).
struct AcceptTable {
  int dim = 2;
  double dE[2][7];
  double prob[2][7];       // exp(-dE*beta)  (unused when dE<0)
  uint32_t thr_m1[2][7];   // checkerboard mode: accept iff r32 <= thr_m1 ...
  bool never[2][7];        // ... unless exp(-dE*beta) == 0: `rand < prob` (metropolis.hh:33) never accepts
  double beta = 0.0;
  bool accept_u32(int sp, int n_up, uint32_t r) const {
    return !never[sp][n_up] && r <= thr_m1[sp][n_up];
  }
};
inline AcceptTable make_accept_table(int dim, double J, double T, double mu) {
  AcceptTable t;
  t.dim = dim;
  t.beta = 1.0 / (KB * T);
  const int z = 2 * dim;
  for (int sp = 0; sp < 2; ++sp) {
    int s = sp ? 1 : -1;
    int new_occ = -s;
    for (int nu = 0; nu <= z; ++nu) {
      int nb_sum = 2 * nu - z;
      double dE_f = -J * (new_occ - s) * nb_sum;
      double Ndx = 0.0;
      Ndx += (new_occ - s) / 2.0;
      double dE = dE_f - mu * Ndx;
      t.dE[sp][nu] = dE;
      double p = std::exp(-dE * t.beta);
      t.prob[sp][nu] = p;
      uint32_t thr;
      t.never[sp][nu] = false;
      if (dE < 0.0 || p >= 1.0) {
        thr = 0xFFFFFFFFu;
      } else if (!(p > 0.0)) {
        thr = 0u;
        t.never[sp][nu] = true;
      } else {
        double scaled = std::ceil(p * 4294967296.0);  // in [0, 2^32]
        if (scaled < 1.0) scaled = 1.0;
        if (scaled > 4294967296.0) scaled = 4294967296.0;
        thr = static_cast<uint32_t>(static_cast<uint64_t>(scaled) - 1u);
      }
      t.thr_m1[sp][nu] = thr;
    }
  }
  return t;
}

// Checkerboard lattice update, scalar restatement of the production kernels.
// Colour c = (i + j [+ k]) & 1.  Sites of one colour are numbered by the
// "plane index" q = (i >> 1) + (n0/2) * (j + n1 * k).  The acceptance uniform
// of site q in pass t, colour c, chain ch is the 32-bit integer
//     R = (rotl16(r16, 1) << 16) | r16'
// (the rotation puts the low 15 bits of the lane on top, which is what the
// kernels' packed 15-bit first-stage compare looks at)
// where r16 is 16-bit lane (q & 7) of
//   Philox4x32-10(counter = {lo32(q>>3), (hi32(q>>3)&0xff) | ch<<8, lo32(t),
//                            (hi32(t)<<2) | c},          key = {lo32(seed), hi32(seed)})
// (lane l = bits [16*(l&1), 16*(l&1)+16) of output word l>>1) and r16' is the
// same lane of the call with counter word 3 | 2 ("refinement" stream; the
// kernels only evaluate it when r16 & 0x7FFF ties with the top 15 bits of the threshold).
// The site is flipped iff R <= thr_m1[b][n_up] (and never when exp(-dE*beta) == 0).
// Requires even extents.  One pass = colour 0 half-sweep then colour 1.
struct CheckerboardResult {
  long long n_accept = 0;
};
inline uint32_t checkerboard_uniform(uint64_t q, uint32_t chain, uint64_t pass_index,
                                     int colour, std::array<uint32_t, 2> key, int rounds = 10) {
  const uint64_t g = q >> 3;
  const int lane = static_cast<int>(q & 7);
  std::array<uint32_t, 4> ctr = {
      static_cast<uint32_t>(g),
      (static_cast<uint32_t>(g >> 32) & 0xffu) | (chain << 8),
      static_cast<uint32_t>(pass_index),
      (static_cast<uint32_t>(pass_index >> 32) << 2) | static_cast<uint32_t>(colour)};
  const uint32_t w0 = Philox4x32::generate(ctr, key, rounds)[lane >> 1];
  ctr[3] |= 2u;
  const uint32_t w1 = Philox4x32::generate(ctr, key, rounds)[lane >> 1];
  const uint32_t r16 = (w0 >> (16 * (lane & 1))) & 0xffffu;
  const uint32_t r16b = (w1 >> (16 * (lane & 1))) & 0xffffu;
  const uint32_t lead = ((r16 << 1) | (r16 >> 15)) & 0xffffu;  // rotl16(r16, 1)
  return (lead << 16) | r16b;
}
inline void checkerboard_pass(std::vector<int> &occ, std::vector<int> const &shape,
                              AcceptTable const &tab, uint64_t seed,
                              uint32_t chain, uint64_t pass_index,
                              CheckerboardResult &res, int rounds = 10) {
  const int dim = static_cast<int>(shape.size());
  const long n0 = shape[0], n1 = shape[1], n2 = (dim == 3) ? shape[2] : 1;
  const long h = n0 / 2;
  std::array<uint32_t, 2> key = {static_cast<uint32_t>(seed),
                                 static_cast<uint32_t>(seed >> 32)};
  for (int colour = 0; colour < 2; ++colour) {
    for (long k = 0; k < n2; ++k)
      for (long j = 0; j < n1; ++j)
        for (long p = 0; p < h; ++p) {
          long i = 2 * p + ((j + k + colour) & 1);
          long l = i + n0 * (j + n1 * k);
          uint64_t q = static_cast<uint64_t>(p) +
                       static_cast<uint64_t>(h) *
                           (static_cast<uint64_t>(j) +
                            static_cast<uint64_t>(n1) * static_cast<uint64_t>(k));
          uint32_t r = checkerboard_uniform(q, chain, pass_index, colour, key, rounds);
          long ip = (i + 1) % n0, im = (i + n0 - 1) % n0;
          long jp = (j + 1) % n1, jm = (j + n1 - 1) % n1;
          int n_up = (occ[ip + n0 * (j + n1 * k)] > 0) +
                     (occ[im + n0 * (j + n1 * k)] > 0) +
                     (occ[i + n0 * (jp + n1 * k)] > 0) +
                     (occ[i + n0 * (jm + n1 * k)] > 0);
          if (dim == 3) {
            long kp = (k + 1) % n2, km = (k + n2 - 1) % n2;
            n_up += (occ[i + n0 * (j + n1 * kp)] > 0) +
                    (occ[i + n0 * (j + n1 * km)] > 0);
          }
          int sp = occ[l] > 0 ? 1 : 0;
          if (tab.accept_u32(sp, n_up, r)) {
            occ[l] = -occ[l];
            res.n_accept++;
          }
        }
  }
}

// One coloured half-sweep of a column slab (2-d): the slab holds global columns
// [col_begin, col_begin + n_cols) in `occ` (n0 x n_cols, column-major) and the
// two neighbouring columns in halo_lo / halo_hi (n0 entries each).  Same
// uniforms as checkerboard_pass (keyed on GLOBAL plane indices), so stitching
// the slabs back together reproduces the undecomposed trajectory bit for bit.
inline long long checkerboard_half_sweep_slab(std::vector<int> &occ, std::vector<int> const &halo_lo,
                                              std::vector<int> const &halo_hi, long n0,
                                              long col_begin, long n_cols,
                                              AcceptTable const &tab, uint64_t seed, uint32_t chain,
                                              uint64_t pass_index, int colour) {
  const long h = n0 / 2;
  std::array<uint32_t, 2> key = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  long long n_accept = 0;
  for (long jl = 0; jl < n_cols; ++jl) {
    const long j = col_begin + jl;
    for (long p = 0; p < h; ++p) {
      const long i = 2 * p + ((j + colour) & 1);
      const uint64_t q = static_cast<uint64_t>(p) + static_cast<uint64_t>(h) * static_cast<uint64_t>(j);
      const uint32_t r = checkerboard_uniform(q, chain, pass_index, colour, key);
      const long ip = (i + 1) % n0, im = (i + n0 - 1) % n0;
      const int left = (jl == 0) ? halo_lo[i] : occ[i + n0 * (jl - 1)];
      const int right = (jl == n_cols - 1) ? halo_hi[i] : occ[i + n0 * (jl + 1)];
      const int n_up = (occ[ip + n0 * jl] > 0) + (occ[im + n0 * jl] > 0) + (left > 0) + (right > 0);
      const long l = i + n0 * jl;
      const int sp = occ[l] > 0 ? 1 : 0;
      if (tab.accept_u32(sp, n_up, r)) {
        occ[l] = -occ[l];
        ++n_accept;
      }
    }
  }
  return n_accept;
}

// Integer observables: S = sum_l s_l ; B = sum_l s_l*(s_{+i} + s_{+j} [+ s_{+k}])
inline void integer_observables(std::vector<int> const &occ,
                                std::vector<int> const &shape, long long &S,
                                long long &B) {
  const int dim = static_cast<int>(shape.size());
  const long n0 = shape[0], n1 = shape[1], n2 = (dim == 3) ? shape[2] : 1;
  S = 0;
  B = 0;
  for (long k = 0; k < n2; ++k)
    for (long j = 0; j < n1; ++j)
      for (long i = 0; i < n0; ++i) {
        long l = i + n0 * (j + n1 * k);
        int s = occ[l];
        S += s;
        int nb = occ[(i + 1) % n0 + n0 * (j + n1 * k)] +
                 occ[i + n0 * ((j + 1) % n1 + n1 * k)];
        if (dim == 3) nb += occ[i + n0 * (j + n1 * ((k + 1) % n2))];
        B += s * nb;
      }
}

// The three default observables from (S, B), with the reference's expression
// order (model.hh:266-270, :412-422; basic_semigrand_canonical.hh:165-174).
struct IntensiveObservables {
  double param_composition, formation_energy, potential_energy;
};
inline IntensiveObservables observables_from_sums(long long S, long long B,
                                                  long long N, double J,
                                                  double mu) {
  IntensiveObservables o;
  double e_formation = static_cast<double>(B);
  e_formation *= -J;
  double Nx = static_cast<double>(N + S) / 2.0;
  o.param_composition = Nx / static_cast<double>(N);
  o.formation_energy = e_formation / static_cast<double>(N);
  double e_pot = e_formation - mu * Nx;
  o.potential_energy = e_pot / static_cast<double>(N);
  return o;
}

// Derived thermodynamics (not in the reference):
// C = N*Var(e_pot)/(KB*T^2), chi = N*Var(x)/(KB*T), population variance.
inline double heat_capacity(std::vector<double> const &e_pot, long long N,
                            double T) {
  double m = mean_of(e_pot.data(), e_pot.size());
  return N * variance(e_pot.data(), e_pot.size(), m) / (KB * T * T);
}
inline double susceptibility(std::vector<double> const &x, long long N,
                             double T) {
  double m = mean_of(x.data(), x.size());
  return N * variance(x.data(), x.size(), m) / (KB * T);
}

// ---------------------------------------------------------------------------
// Conversions, index arithmetic only (include/casm/monte/Conversions.hh:43-135,
// src/casm/monte/Conversions.cc:181-229).  The arithmetic is delegated by the
// reference to xtal::UnitCellCoordIndexConverter (CASMcode_crystallography,
// absent).  Pinned by python/tests/events/test_Conversions.py:
// l = b * n_unitcells + unitcell_index, periodic wrap of ijk.  The order of
// unit cells inside unitcell_index is UNPINNED; for a diagonal transformation
// matrix diag(n0,n1,n2) it is restated as first-index-fastest.
// ---------------------------------------------------------------------------
struct DiagonalConversions {
  long n[3];
  long n_basis;
  long n_unitcells() const { return n[0] * n[1] * n[2]; }
  long l_size() const { return n_basis * n_unitcells(); }
  long l_to_b(long l) const { return l / n_unitcells(); }
  void l_to_ijk(long l, long ijk[3]) const {
    long u = l % n_unitcells();
    ijk[0] = u % n[0];
    ijk[1] = (u / n[0]) % n[1];
    ijk[2] = u / (n[0] * n[1]);
  }
  static long wrap(long a, long m) {
    long r = a % m;
    return r < 0 ? r + m : r;
  }
  long bijk_to_l(long b, long i, long j, long k) const {
    return b * n_unitcells() + wrap(i, n[0]) +
           n[0] * (wrap(j, n[1]) + n[1] * wrap(k, n[2]));
  }
};

}  // namespace monte_oracle

#endif
