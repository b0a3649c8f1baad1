// casm_monte_b200/monte.hh -- host-side C++ mirror of the libcasm-monte interface
// for the Ising semi-grand-canonical Metropolis path, implemented above the C
// ABI of include/casm_monte_gpu.h (all arithmetic of the path runs on the GPU).
//
// Same class names, method names, argument meaning and error behaviour
// (std::runtime_error) as the reference headers cited at each class, with
// std::vector in place of Eigen (Eigen is an external dependency of the
// reference and is not needed at this boundary).  Header-only; link against
// libcasm_monte_b200.so.
//
// What differs from the reference, by design:
//  * IsingConfiguration keeps a device-resident copy of the occupation next to
//    its host mirror; the calculators evaluate on the device.
//  * SemiGrandCanonicalCalculator::run drives the device loop at pass
//    granularity (the reference loop can only terminate at a pass boundary,
//    see SURVEY 3.2) and has one extra knob, `update_mode`:
//      "serial_reference": the reference's serial random-site order on the
//                          reference's mt19937_64 stream (trajectory-exact);
//      "checkerboard":     production order (needs even extents);
//      "auto" (default):   checkerboard when the extents are even, else serial.
#ifndef CASM_MONTE_B200_MONTE_HH
#define CASM_MONTE_B200_MONTE_HH

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../casm_monte_gpu.h"
#include "snf.hh"

namespace casm_monte_b200 {

// include/casm/monte/definitions.hh:17,25-27
typedef long Index;
typedef long CountType;
typedef long long BigCountType;
typedef double TimeType;
typedef std::mt19937_64 default_engine_type;
constexpr double KB = CMG_KB;

// include/casm/monte/definitions.hh:107-113
template <typename PtrType>
PtrType throw_if_null(PtrType ptr, std::string const &what) {
  if (ptr == nullptr) throw std::runtime_error(what);
  return ptr;
}

inline void cmg_check(int rc, cmg_context const *ctx = nullptr) {
  if (rc != CMG_OK) {
    const char *m = ctx ? cmg_last_error(ctx) : cmg_last_global_error();
    throw std::runtime_error(std::string(m ? m : "casm_monte_gpu error"));
  }
}

/// RAII owner of one C-ABI context (one lattice, one chain)
class DeviceLattice {
 public:
  DeviceLattice(std::vector<int> const &shape, int device = 0) {
    int64_t sh[3] = {1, 1, 1};
    for (size_t d = 0; d < shape.size() && d < 3; ++d) sh[d] = shape[d];
    cmg_check(cmg_create(static_cast<int>(shape.size()), sh, 1, device, &m_ctx));
  }
  ~DeviceLattice() { cmg_destroy(m_ctx); }
  DeviceLattice(DeviceLattice const &) = delete;
  DeviceLattice &operator=(DeviceLattice const &) = delete;
  cmg_context *ctx() const { return m_ctx; }
  void check(int rc) const { cmg_check(rc, m_ctx); }

 private:
  cmg_context *m_ctx = nullptr;
};

// ---------------------------------------------------------------------------
// ValueMap (include/casm/monte/ValueMap.hh:13-74)
// ---------------------------------------------------------------------------
struct MatrixValue {
  Index rows = 0, cols = 0;
  std::vector<double> data;  // column-major
};
struct ValueMap {
  std::map<std::string, bool> boolean_values;
  std::map<std::string, double> scalar_values;
  std::map<std::string, std::vector<double>> vector_values;
  std::map<std::string, MatrixValue> matrix_values;
};
inline bool is_mismatched(ValueMap const &A, ValueMap const &B) {
  for (auto const &p : B.boolean_values)
    if (!A.boolean_values.count(p.first)) return true;
  for (auto const &p : B.scalar_values)
    if (!A.scalar_values.count(p.first)) return true;
  for (auto const &p : B.vector_values)
    if (!A.vector_values.count(p.first)) return true;
  for (auto const &p : B.matrix_values)
    if (!A.matrix_values.count(p.first)) return true;
  return false;
}
inline ValueMap make_incremented_values(ValueMap values, ValueMap const &increment,
                                        double n_increment) {
  for (auto const &p : increment.scalar_values)
    values.scalar_values.at(p.first) += p.second * n_increment;
  for (auto const &p : increment.vector_values) {
    auto &v = values.vector_values.at(p.first);
    for (size_t i = 0; i < v.size(); ++i) v[i] += p.second[i] * n_increment;
  }
  for (auto const &p : increment.matrix_values) {
    auto &v = values.matrix_values.at(p.first).data;
    for (size_t i = 0; i < v.size(); ++i) v[i] += p.second.data[i] * n_increment;
  }
  return values;
}

// ---------------------------------------------------------------------------
// MethodLog (include/casm/monte/MethodLog.hh:13-38) with the wall clock that
// CASM::Log provides in the reference
// ---------------------------------------------------------------------------
struct LogClock {
  typedef std::chrono::steady_clock clk;
  clk::time_point t0 = clk::now(), lap0 = clk::now();
  std::ostream *out = &std::cout;
  void restart_clock() { t0 = clk::now(); }
  void begin_lap() { lap0 = clk::now(); }
  double time_s() const { return std::chrono::duration<double>(clk::now() - t0).count(); }
  double lap_time() const { return std::chrono::duration<double>(clk::now() - lap0).count(); }
};
struct MethodLog {
  std::string logfile_path;
  std::shared_ptr<std::ofstream> fout;
  LogClock log;
  std::optional<double> log_frequency;
  void reset() {
    if (!logfile_path.empty()) {
      fout = std::make_shared<std::ofstream>(logfile_path);
      log.out = fout.get();
    }
  }
  void reset_to_stdout() {
    fout.reset();
    log.out = &std::cout;
  }
};

// ---------------------------------------------------------------------------
// RandomNumberGenerator (include/casm/monte/RandomNumberGenerator.hh:15-42)
// ---------------------------------------------------------------------------
template <typename EngineType = default_engine_type>
struct RandomNumberGenerator {
  std::shared_ptr<EngineType> engine;
  explicit RandomNumberGenerator(std::shared_ptr<EngineType> _engine = std::shared_ptr<EngineType>())
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

/// engine state <-> the 312 words + position that operator<< prints
inline void engine_to_words(default_engine_type const &e, uint64_t *words312, int *pos) {
  std::stringstream ss;
  ss << e;
  for (int i = 0; i < 312; ++i) ss >> words312[i];
  ss >> *pos;
}
inline void words_to_engine(uint64_t const *words312, int pos, default_engine_type &e) {
  std::stringstream ss;
  for (int i = 0; i < 312; ++i) ss << words312[i] << ' ';
  ss << pos;
  ss >> e;
}

// ---------------------------------------------------------------------------
// OccEvent (include/casm/monte/events/OccEvent.hh:33-73).  atom_traj (species
// trajectories for kinetic Monte Carlo) is out of scope.
// ---------------------------------------------------------------------------
struct OccTransform {
  Index l = 0;             ///< Config occupant that is being transformed
  Index mol_id = 0;        ///< Location in OccLocation.m_mol
  Index asym = 0;          ///< Asym index
  Index from_species = 0;  ///< Species index before transformation
  Index to_species = 0;    ///< Species index after transformation
};
struct OccEvent {
  std::vector<Index> linear_site_index;
  std::vector<int> new_occ;
  /// used to update the occupant tracking of OccLocation (events.hh)
  std::vector<OccTransform> occ_transform;
};

// ---------------------------------------------------------------------------
// IsingConfiguration (include/casm/monte/ising_cpp/model.hh:19-110)
// ---------------------------------------------------------------------------
class IsingConfiguration {
 public:
  IsingConfiguration() : IsingConfiguration(std::vector<int>{0, 0}, 1) {}

  explicit IsingConfiguration(std::vector<int> _shape, int fill_value = 1)
      : shape(std::move(_shape)) {
    if (shape.size() != 2 && shape.size() != 3)  // 3-d is this implementation's extension
      throw std::runtime_error("IsingConfiguration only supports 2d");
    Index n = 1;
    for (int s : shape) n *= static_cast<Index>(s);
    m_occupation.assign(static_cast<size_t>(n), fill_value);
    n_sites = n_variable_sites = n_unitcells = n;
    if (fill_value == 1 || fill_value == -1) m_uniform = fill_value;  // filled on the device, not uploaded
  }
  // deep copy of the host mirror; the copy gets its own device lattice on demand
  IsingConfiguration(IsingConfiguration const &o)
      : shape(o.shape), n_sites(o.n_sites), n_variable_sites(o.n_variable_sites),
        n_unitcells(o.n_unitcells), device_index(o.device_index) {
    o.pull();
    m_occupation = o.m_occupation;
    m_uniform = o.m_uniform;
  }
  IsingConfiguration &operator=(IsingConfiguration const &o) {
    if (this == &o) return *this;
    o.pull();
    shape = o.shape;
    n_sites = o.n_sites;
    n_variable_sites = o.n_variable_sites;
    n_unitcells = o.n_unitcells;
    device_index = o.device_index;
    m_occupation = o.m_occupation;
    m_uniform = o.m_uniform;
    m_dev.reset();
    m_host_valid = true;
    m_dev_valid = false;
    return *this;
  }

  std::vector<int> shape;
  Index n_sites = 0, n_variable_sites = 0, n_unitcells = 0;
  int device_index = 0;

  std::vector<int> const &occupation() const {
    pull();
    return m_occupation;
  }
  void set_occupation(std::vector<int> const &occupation) {
    if (m_occupation.size() != occupation.size())
      throw std::runtime_error("Error in set_occupation: size mismatch");
    m_occupation = occupation;
    m_uniform = 0;
    m_host_valid = true;
    m_dev_valid = false;
  }
  int occ(Index l) const {
    pull();
    return m_occupation[l];
  }
  void set_occ(Index l, int new_occ) {
    pull();
    m_occupation[l] = new_occ;
    m_uniform = 0;
    if (m_dev && m_dev_valid) m_dev->check(cmg_set_occ(m_dev->ctx(), 0, l, new_occ));
  }
  Index within(Index index, int dim) const {
    Index r = index % shape[dim];
    if (r < 0) r += shape[dim];
    return r;
  }
  std::vector<int> from_linear_site_index(Index l) const {
    std::vector<int> mi(shape.size());
    mi[0] = static_cast<int>(l % shape[0]);
    Index r = l / shape[0];
    if (shape.size() == 2) {
      mi[1] = static_cast<int>(r);
    } else {
      mi[1] = static_cast<int>(r % shape[1]);
      mi[2] = static_cast<int>(r / shape[1]);
    }
    return mi;
  }
  Index to_linear_site_index(std::vector<int> const &mi) const {
    if (shape.size() == 2) return static_cast<Index>(shape[0]) * mi[1] + mi[0];
    return mi[0] + static_cast<Index>(shape[0]) * (mi[1] + static_cast<Index>(shape[1]) * mi[2]);
  }
  Index to_linear_site_index(Index row, Index col) const {
    return static_cast<Index>(shape[0]) * col + row;
  }

  // ---- device side (not in the reference) ----
  /// the device lattice, holding the current occupation
  DeviceLattice &device() const {
    if (!m_dev) {
      m_dev = std::make_shared<DeviceLattice>(shape, device_index);
      m_dev_valid = false;
    }
    if (!m_dev_valid) {
      if (m_occupation.empty()) throw std::runtime_error("empty configuration");
      static_assert(sizeof(int) == sizeof(int32_t), "the host mirror is the reference's VectorXi");
      if (m_uniform != 0) {
        // a freshly constructed configuration: fill on the device instead of copying 4 B per site
        m_dev->check(cmg_fill_occupation(m_dev->ctx(), 0, m_uniform));
      } else {
        m_dev->check(cmg_upload_occupation_i32(m_dev->ctx(), 0, reinterpret_cast<int32_t const *>(m_occupation.data()),
                                               static_cast<int64_t>(m_occupation.size())));
      }
      m_dev_valid = true;
    }
    return *m_dev;
  }
  /// the device copy was modified by a kernel: the host mirror is stale
  void mark_device_modified() const {
    m_host_valid = false;
    m_uniform = 0;
  }
  /// refresh the host mirror from the device if needed (every accessor does)
  void pull() const {
    if (m_host_valid) return;
    m_dev->check(cmg_download_occupation_i32(m_dev->ctx(), 0, reinterpret_cast<int32_t *>(m_occupation.data()),
                                             static_cast<int64_t>(m_occupation.size())));
    m_host_valid = true;
  }

 private:
  mutable std::vector<int> m_occupation;
  mutable std::shared_ptr<DeviceLattice> m_dev;
  mutable bool m_host_valid = true;
  mutable bool m_dev_valid = false;
  mutable int m_uniform = 0;  // +1 / -1: every site holds this value (as constructed); 0: unknown
};

// model.hh:141-157
class IsingState {
 public:
  IsingState(IsingConfiguration _configuration, ValueMap _conditions,
             ValueMap _properties = ValueMap())
      : configuration(std::move(_configuration)), conditions(std::move(_conditions)),
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
  IsingFormationEnergy(double _J = 1.0, int _lattice_type = 1, bool _use_nlist = true,
                       state_type const * /*_state*/ = nullptr)
      : J(_J), lattice_type(_lattice_type), state(nullptr), m_use_nlist(_use_nlist) {
    if (lattice_type != 1) throw std::runtime_error("Unsupported lattice_type");
  }
  double J;
  int lattice_type;
  state_type const *state;
  bool use_nlist() const { return m_use_nlist; }

  void set_state(state_type const *_state) {
    state = throw_if_null(_state, "Error in IsingFormationEnergy::set_state: _state==nullptr");
  }
  /// model.hh:259-290.  nlist form: -J * double(B); non-nlist form: the sum over
  /// rows then columns of (-J * integer line dot product), in that order.
  double per_supercell() const {
    DeviceLattice &dev = ready();
    if (m_use_nlist || state->configuration.shape.size() != 2) {
      int64_t S = 0, B = 0;
      dev.check(cmg_sample_now(dev.ctx(), 0, &S, &B));
      double e_formation = static_cast<double>(B);
      e_formation *= -J;
      return e_formation;
    }
    std::vector<int64_t> rows(state->configuration.shape[0]), cols(state->configuration.shape[1]);
    dev.check(cmg_line_dots(dev.ctx(), 0, rows.data(), cols.data()));
    double e_formation = 0.0;
    for (int64_t d : rows) e_formation += -J * static_cast<double>(d);
    for (int64_t d : cols) e_formation += -J * static_cast<double>(d);
    return e_formation;
  }
  double per_unitcell() const { return per_supercell() / state->configuration.n_unitcells; }
  /// model.hh:354-379
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) const {
    DeviceLattice &dev = ready();
    std::vector<int64_t> ls(linear_site_index.begin(), linear_site_index.end());
    std::vector<int32_t> no(new_occ.begin(), new_occ.end());
    double dE = 0.0, dNx = 0.0;
    dev.check(cmg_event_delta(dev.ctx(), 0, static_cast<int>(ls.size()), ls.data(), no.data(), &dE, &dNx));
    return dE;
  }

 private:
  DeviceLattice &ready() const {
    if (state == nullptr) throw std::runtime_error("IsingFormationEnergy: state==nullptr");
    DeviceLattice &dev = state->configuration.device();
    dev.check(cmg_set_model(dev.ctx(), J, lattice_type));
    return dev;
  }
  bool m_use_nlist = true;
};

// ---------------------------------------------------------------------------
// IsingParamComposition (model.hh:388-436)
// ---------------------------------------------------------------------------
class IsingParamComposition {
 public:
  typedef IsingState state_type;
  explicit IsingParamComposition(state_type const * /*_state*/ = nullptr) : state(nullptr) {}
  state_type const *state;
  void set_state(state_type const *_state) {
    state = throw_if_null(_state, "Error in IsingParamComposition::set_state: _state==nullptr");
  }
  Index n_independent_compositions() const { return 1; }
  std::vector<double> per_supercell() const {
    if (state == nullptr) throw std::runtime_error("IsingParamComposition: state==nullptr");
    DeviceLattice &dev = state->configuration.device();
    int64_t S = 0, B = 0;
    dev.check(cmg_sample_now(dev.ctx(), 0, &S, &B));
    std::vector<double> r(1);
    r[0] = static_cast<double>(state->configuration.n_sites + S) / 2.0;
    return r;
  }
  std::vector<double> per_unitcell() const {
    std::vector<double> r = per_supercell();
    r[0] = r[0] / static_cast<double>(state->configuration.n_unitcells);
    return r;
  }
  std::vector<double> occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                              std::vector<int> const &new_occ) const {
    if (state == nullptr) throw std::runtime_error("IsingParamComposition: state==nullptr");
    DeviceLattice &dev = state->configuration.device();
    std::vector<int64_t> ls(linear_site_index.begin(), linear_site_index.end());
    std::vector<int32_t> no(new_occ.begin(), new_occ.end());
    double dE = 0.0, dNx = 0.0;
    dev.check(cmg_event_delta(dev.ctx(), 0, static_cast<int>(ls.size()), ls.data(), no.data(), &dE, &dNx));
    return std::vector<double>{dNx};
  }
};

// model.hh:439-452
class IsingSystem {
 public:
  typedef IsingState state_type;
  typedef IsingFormationEnergy formation_energy_f_type;
  typedef IsingParamComposition param_composition_f_type;
  IsingSystem(formation_energy_f_type f, param_composition_f_type c)
      : formation_energy_calculator(std::move(f)), param_composition_calculator(std::move(c)) {}
  formation_energy_f_type formation_energy_calculator;
  param_composition_f_type param_composition_calculator;
};

// ---------------------------------------------------------------------------
// Sampler & friends (include/casm/monte/sampling/Sampler.hh)
// ---------------------------------------------------------------------------
inline std::vector<std::string> colmajor_component_names(Index n_rows, Index n_cols) {
  std::vector<std::string> r;
  for (Index c = 0; c < n_cols; ++c)
    for (Index w = 0; w < n_rows; ++w) r.push_back(std::to_string(w) + "," + std::to_string(c));
  return r;
}
inline std::vector<std::string> default_component_names(std::vector<Index> const &shape) {
  if (shape.empty()) return {"0"};
  if (shape.size() == 1) {
    std::vector<std::string> r;
    for (Index i = 0; i < shape[0]; ++i) r.push_back(std::to_string(i));
    return r;
  }
  if (shape.size() == 2) return colmajor_component_names(shape[0], shape[1]);
  throw std::runtime_error(
      "Error constructing sampler component names: >2 dimensions is not supported");
}

/// Row-per-sample storage, one contiguous column per component (the reference's
/// column-major Eigen::MatrixXd), capacity grown by `capacity_increment` rows.
class Sampler {
 public:
  explicit Sampler(std::vector<Index> _shape, CountType _capacity_increment = 1000)
      : m_component_names(default_component_names(_shape)), m_shape(_shape),
        m_capacity_increment(_capacity_increment) {
    m_n_components = 1;
    for (Index x : _shape) m_n_components *= x;
    clear();
  }
  Sampler(std::vector<Index> _shape, std::vector<std::string> const &names,
          CountType _capacity_increment = 1000)
      : m_n_components(static_cast<Index>(names.size())), m_component_names(names),
        m_shape(_shape), m_capacity_increment(_capacity_increment) {
    clear();
  }
  void push_back(double value) {
    grow();
    m_cols[0][m_n_samples++] = value;
  }
  void push_back(std::vector<double> const &v) {
    if (static_cast<Index>(v.size()) != m_n_components)
      throw std::runtime_error("Error in Sampler::push_back: vector size != n_components");
    grow();
    for (Index c = 0; c < m_n_components; ++c) m_cols[c][m_n_samples] = v[c];
    ++m_n_samples;
  }
  /// append `n` samples of a 1-component quantity at once (device series -> host)
  void append_column(double const *x, CountType n) {
    if (m_n_components != 1) throw std::runtime_error("append_column needs 1 component");
    while (m_n_samples + n > m_capacity) set_sample_capacity(m_capacity + m_capacity_increment);
    std::copy(x, x + n, m_cols[0].begin() + m_n_samples);
    m_n_samples += n;
  }
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
  std::vector<std::string> const &component_names() const { return m_component_names; }
  std::vector<Index> const &shape() const { return m_shape; }
  Index n_components() const { return m_n_components; }
  CountType n_samples() const { return m_n_samples; }
  CountType sample_capacity() const { return m_capacity; }
  std::vector<double> component(Index c) const {
    if (m_cols.empty()) return {};
    return std::vector<double>(m_cols[c].begin(), m_cols[c].begin() + m_n_samples);
  }
  double const *component_data(Index c) const { return m_cols[c].data(); }
  std::vector<double> sample(CountType r) const {
    std::vector<double> v(m_n_components);
    for (Index c = 0; c < m_n_components; ++c) v[c] = m_cols[c][r];
    return v;
  }

 private:
  void grow() {
    if (m_n_samples == m_capacity) set_sample_capacity(m_capacity + m_capacity_increment);
  }
  Index m_n_components = 1;
  std::vector<std::string> m_component_names;
  std::vector<Index> m_shape;
  Index m_n_samples = 0;
  CountType m_capacity_increment = 1000;
  CountType m_capacity = 0;
  std::vector<std::vector<double>> m_cols;
};
typedef std::map<std::string, std::shared_ptr<Sampler>> SamplerMap;

struct SamplerComponent {
  SamplerComponent(std::string s, Index i, std::string n)
      : sampler_name(std::move(s)), component_index(i), component_name(std::move(n)) {}
  std::string sampler_name;
  Index component_index = 0;
  std::string component_name;
  bool operator<(SamplerComponent const &o) const {
    if (sampler_name == o.sampler_name) return component_index < o.component_index;
    return sampler_name < o.sampler_name;
  }
};
struct RequestedPrecision {
  bool abs_convergence_is_required = false;
  double abs_precision = 0.0;
  bool rel_convergence_is_required = false;
  double rel_precision = 0.0;
  static RequestedPrecision abs_and_rel(double a, double r) {
    RequestedPrecision x;
    x.abs_convergence_is_required = x.rel_convergence_is_required = true;
    x.abs_precision = a;
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

inline SamplerMap::const_iterator find_or_throw(SamplerMap const &samplers,
                                                SamplerComponent const &key) {
  auto it = samplers.find(key.sampler_name);
  if (it == samplers.end()) {
    std::stringstream msg;
    msg << "Error finding sampler component: Sampler '" << key.sampler_name << "' not found."
        << std::endl;
    throw std::runtime_error(msg.str());
  }
  if (key.component_index >= it->second->n_components()) {
    std::stringstream msg;
    msg << "Error finding sampler component: Requested component index " << key.component_index
        << ", but '" << key.sampler_name << "' has " << it->second->n_components()
        << "components." << std::endl;
    throw std::runtime_error(msg.str());
  }
  return it;
}
inline CountType get_n_samples(SamplerMap const &samplers) {
  if (samplers.size()) return samplers.begin()->second->n_samples();
  return CountType(0);
}

/// include/casm/monte/sampling/StateSamplingFunction.hh:20-54.  `builtin` is
/// this implementation's tag for the three default functions (CMG_Q_*), which
/// the device loop samples itself; -1 for user functions.
struct StateSamplingFunction {
  StateSamplingFunction(std::string _name, std::string _description, std::vector<Index> _shape,
                        std::function<std::vector<double>()> _function,
                        std::optional<std::vector<std::string>> _component_names = std::nullopt)
      : name(std::move(_name)), description(std::move(_description)), shape(std::move(_shape)),
        component_names(_component_names.has_value() ? *_component_names
                                                     : default_component_names(shape)),
        function(std::move(_function)) {}
  std::string name, description;
  std::vector<Index> shape;
  std::vector<std::string> component_names;
  std::function<std::vector<double>()> function;
  int builtin = -1;
  void const *builtin_owner = nullptr;
  std::vector<double> operator()() const { return function(); }
};
typedef std::map<std::string, StateSamplingFunction> StateSamplingFunctionMap;

// ---------------------------------------------------------------------------
// BasicStatistics (include/casm/monte/BasicStatistics.hh, src/.../BasicStatistics.cc)
// evaluated on the device (cmg_host_series_stats)
// ---------------------------------------------------------------------------
struct BasicStatistics {
  double mean = 0.0;
  double calculated_precision = std::numeric_limits<double>::max();
};
inline double get_calculated_precision(BasicStatistics const &s) { return s.calculated_precision; }
inline double get_calculated_relative_precision(BasicStatistics const &s) {
  return std::abs(s.calculated_precision / s.mean);
}
struct BasicStatisticsCalculator {
  explicit BasicStatisticsCalculator(double _confidence = 0.95, Index _method = 1,
                                     Index _n_resamples = 10000, int _device = 0)
      : confidence(_confidence), method(_method), n_resamples(_n_resamples), device(_device) {}
  double confidence;
  Index method;
  Index n_resamples;
  int device;
  BasicStatistics operator()(std::vector<double> const &observations) const {
    if (observations.empty())
      throw std::runtime_error("Error in BasicStatisticsCalculator: observations.size()==0");
    BasicStatistics s;
    double var = 0.0;
    int64_t k = 0;
    cmg_check(cmg_host_series_stats(device, observations.data(),
                                    static_cast<int64_t>(observations.size()), confidence, &s.mean,
                                    &s.calculated_precision, &var, &k));
    return s;
  }
  BasicStatistics operator()(std::vector<double> const &observations,
                             std::vector<double> const &sample_weight) const {
    if (observations.empty())
      throw std::runtime_error("Error in BasicStatisticsCalculator: observations.size()==0");
    if (sample_weight.empty()) return (*this)(observations);
    if (observations.size() != sample_weight.size())
      throw std::runtime_error(
          "Error in BasicStatisticsCalculator: observations.size() != sample_weight.size()");
    // BasicStatistics.cc:163-187, evaluated on the device
    if (method != 1 && method != 2)
      throw std::runtime_error("Error in BasicStatisticsCalculator: invalid method");
    BasicStatistics s;
    double var = 0.0, W = 0.0;
    int64_t k = 0;
    cmg_check(cmg_host_series_stats_weighted(
        device, observations.data(), sample_weight.data(),
        static_cast<int64_t>(observations.size()), confidence, static_cast<int>(method),
        static_cast<int64_t>(n_resamples), &s.mean, &s.calculated_precision, &var, &W, &k));
    return s;
  }
};
/// BasicStatistics.cc:50-73: weighted observations -> n_equally_spaced observations
inline std::vector<double> resample(std::vector<double> const &observations,
                                    std::vector<double> const &sample_weight,
                                    double sample_weight_sum, Index n_equally_spaced,
                                    int device = 0) {
  if (observations.empty() || observations.size() != sample_weight.size())
    throw std::runtime_error("Error in resample: observations.size() != sample_weight.size()");
  std::vector<double> out(static_cast<size_t>(n_equally_spaced));
  cmg_check(cmg_host_series_resample(device, observations.data(), sample_weight.data(),
                                     static_cast<int64_t>(observations.size()),
                                     sample_weight_sum, static_cast<int64_t>(n_equally_spaced),
                                     out.data()));
  return out;
}
typedef std::function<BasicStatistics(std::vector<double> const &, std::vector<double> const &)>
    CalcStatisticsFunction;

// ---------------------------------------------------------------------------
// Equilibration / convergence / cutoff / completion checks
// (src/casm/monte/checks/EquilibrationCheck.cc, include/casm/monte/checks/*.hh)
// ---------------------------------------------------------------------------
struct IndividualEquilibrationCheckResult {
  bool is_equilibrated = false;
  CountType N_samples_for_equilibration = 0;
};
/// EquilibrationCheck.cc:119-162; the scan itself runs on the device
inline IndividualEquilibrationCheckResult default_equilibration_check(
    std::vector<double> const &observations, std::vector<double> const &sample_weight,
    RequestedPrecision requested_precision) {
  IndividualEquilibrationCheckResult result;
  double prec;
  if (requested_precision.abs_convergence_is_required) {
    prec = requested_precision.abs_precision;
  } else if (requested_precision.rel_convergence_is_required) {
    if (observations.empty())
      throw std::runtime_error("Error in equilibration_check: observations.size()==0");
    double mean = 0, p = 0, v = 0;
    int64_t k = 0;
    cmg_check(cmg_host_series_stats(0, observations.data(), static_cast<int64_t>(observations.size()),
                                    0.95, &mean, &p, &v, &k));
    prec = std::abs(mean * requested_precision.rel_precision);
  } else {
    result.is_equilibrated = true;
    result.N_samples_for_equilibration = 0;
    return result;
  }
  if (observations.empty())
    throw std::runtime_error("Error in equilibration_check: observations.size()==0");
  int is_eq = 0;
  int64_t n_eq = 0;
  if (!sample_weight.empty()) {
    // EquilibrationCheck.cc:137-161
    if (sample_weight.size() != observations.size())
      throw std::runtime_error(
          "Error in equilibration_check: sample_weight.size() != observations.size()");
    cmg_check(cmg_host_series_equilibration_weighted(
        0, observations.data(), sample_weight.data(), static_cast<int64_t>(observations.size()),
        prec, &is_eq, &n_eq));
  } else {
    cmg_check(cmg_host_series_equilibration(0, observations.data(),
                                            static_cast<int64_t>(observations.size()), prec,
                                            &is_eq, &n_eq));
  }
  result.is_equilibrated = is_eq != 0;
  result.N_samples_for_equilibration = static_cast<CountType>(n_eq);
  return result;
}
typedef std::function<IndividualEquilibrationCheckResult(
    std::vector<double> const &, std::vector<double> const &, RequestedPrecision)>
    EquilibrationCheckFunction;

struct EquilibrationCheckResults {
  bool all_equilibrated = false;
  CountType N_samples_for_all_to_equilibrate = 0;
  std::map<SamplerComponent, IndividualEquilibrationCheckResult> individual_results;
};
inline EquilibrationCheckResults equilibration_check(
    EquilibrationCheckFunction equilibration_check_f, RequestedPrecisionMap const &requested_precision,
    SamplerMap const &samplers, Sampler const &sample_weight, bool check_all) {
  if (equilibration_check_f == nullptr)
    throw std::runtime_error("Error in equilibration_check: equilibration_check_f == nullptr");
  EquilibrationCheckResults results;
  if (!requested_precision.size()) return results;
  results.all_equilibrated = true;
  for (auto const &p : requested_precision) {
    Sampler const &sampler = *find_or_throw(samplers, p.first)->second;
    IndividualEquilibrationCheckResult current = equilibration_check_f(
        sampler.component(p.first.component_index), sample_weight.component(0), p.second);
    results.N_samples_for_all_to_equilibrate = std::max(results.N_samples_for_all_to_equilibrate,
                                                        current.N_samples_for_equilibration);
    results.all_equilibrated &= current.is_equilibrated;
    results.individual_results.emplace(p.first, current);
    if (!check_all && !results.all_equilibrated) break;
  }
  return results;
}

struct IndividualConvergenceCheckResult {
  bool is_converged = false;
  RequestedPrecision requested_precision;
  BasicStatistics stats;
};
struct ConvergenceCheckResults {
  bool all_converged = false;
  CountType N_samples_for_statistics = 0;
  std::map<SamplerComponent, IndividualConvergenceCheckResult> individual_results;
};
inline IndividualConvergenceCheckResult convergence_check(BasicStatistics const &stats,
                                                          RequestedPrecision const &req) {
  IndividualConvergenceCheckResult r;
  r.stats = stats;
  r.requested_precision = req;
  r.is_converged = true;
  if (req.abs_convergence_is_required)
    r.is_converged &= get_calculated_precision(stats) < req.abs_precision;
  if (req.rel_convergence_is_required)
    r.is_converged &= get_calculated_relative_precision(stats) < req.rel_precision;
  return r;
}
inline IndividualConvergenceCheckResult component_convergence_check(
    Sampler const &sampler, Sampler const &sample_weight, SamplerComponent const &key,
    RequestedPrecision const &req, CountType N_samples_for_statistics,
    CalcStatisticsFunction calc_statistics_f) {
  if (calc_statistics_f == nullptr)
    throw std::runtime_error("Error in component_convergence_check: calc_statistics_f == nullptr");
  std::vector<double> col = sampler.component(key.component_index);
  std::vector<double> tail(col.end() - N_samples_for_statistics, col.end());
  std::vector<double> w;
  if (sample_weight.n_samples() != 0) {
    std::vector<double> wc = sample_weight.component(0);
    w.assign(wc.end() - N_samples_for_statistics, wc.end());
  }
  return convergence_check(calc_statistics_f(tail, w), req);
}
inline ConvergenceCheckResults convergence_check(SamplerMap const &samplers,
                                                 Sampler const &sample_weight,
                                                 RequestedPrecisionMap const &requested_precision,
                                                 CountType N_samples_for_equilibration,
                                                 CalcStatisticsFunction calc_statistics_f) {
  ConvergenceCheckResults results;
  CountType N_samples = get_n_samples(samplers);
  if (!requested_precision.size()) {
    results.N_samples_for_statistics = N_samples;
    return results;
  }
  if (N_samples_for_equilibration >= N_samples) return results;
  results.N_samples_for_statistics = N_samples - N_samples_for_equilibration;
  results.all_converged = true;
  for (auto const &p : requested_precision) {
    Sampler const &sampler = *find_or_throw(samplers, p.first)->second;
    IndividualConvergenceCheckResult current = component_convergence_check(
        sampler, sample_weight, p.first, p.second, results.N_samples_for_statistics,
        calc_statistics_f);
    results.all_converged &= current.is_converged;
    results.individual_results.emplace(p.first, current);
  }
  return results;
}

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
inline bool all_minimums_met(CutoffCheckParams const &p, std::optional<CountType> count,
                             std::optional<TimeType> time, CountType n_samples,
                             TimeType clocktime) {
  if (p.min_sample.has_value() && n_samples < p.min_sample.value()) return false;
  if (p.min_count.has_value() && count.has_value() && count.value() < p.min_count.value())
    return false;
  if (p.min_time.has_value() && time.has_value() && time.value() < p.min_time.value())
    return false;
  if (p.min_clocktime.has_value() && clocktime < p.min_clocktime.value()) return false;
  return true;
}
inline bool any_maximum_met(CutoffCheckParams const &p, std::optional<CountType> count,
                            std::optional<TimeType> time, CountType n_samples,
                            TimeType clocktime) {
  if (p.max_sample.has_value() && n_samples >= p.max_sample.value()) return true;
  if (p.max_count.has_value() && count.has_value() && count.value() >= p.max_count.value())
    return true;
  if (p.max_time.has_value() && time.has_value() && time.value() >= p.max_time.value())
    return true;
  if (p.max_clocktime.has_value() && clocktime >= p.max_clocktime.value()) return true;
  return false;
}

/// include/casm/monte/checks/CompletionCheck.hh:20-100
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
  CountType sample_check_linear(Index n) const { return check_begin + check_period * n; }
  CountType sample_check_log(Index n) const {
    return check_begin +
           static_cast<CountType>(std::round(std::pow(check_base, (n + check_shift))));
  }
  Index find_n_begin_linear() const {
    Index n = 0;
    while (sample_check_log(n + 1) - sample_check_log(n) <= check_period_max) n += 1;
    return n;
  }
  CountType sample_check_log(Index n, Index n_begin_linear) const {
    if (n <= n_begin_linear) return sample_check_log(n);
    return sample_check_log(n_begin_linear) + check_period_max * (n - n_begin_linear);
  }
};

/// CompletionCheck.hh:102-173
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
                     std::optional<TimeType> _time = std::nullopt, TimeType _clocktime = 0.0,
                     CountType _n_samples = 0) {
    count = _count;
    time = _time;
    clocktime = _clocktime;
    n_samples = _n_samples;
    has_all_minimums_met = false;
    has_any_maximum_met = false;
    is_complete = false;
  }
  void full_reset(std::optional<CountType> _count = std::nullopt,
                  std::optional<TimeType> _time = std::nullopt, TimeType _clocktime = 0.0,
                  CountType _n_samples = 0) {
    partial_reset(_count, _time, _clocktime, _n_samples);
    n_samples_at_convergence_check = std::nullopt;
    equilibration_check_results = EquilibrationCheckResults();
    convergence_check_results = ConvergenceCheckResults();
  }
};

/// This implementation only: the scheduled checks of a CompletionCheck can be
/// evaluated on the device-resident sample series (one call per check, no copy of
/// the series) instead of on the host samplers.  `check` fills, for the requested
/// components in map order, the equilibration results and -- if all of them
/// equilibrated -- the statistics of the last n_stats samples.
struct DeviceSeriesCheck {
  std::function<CountType()> n_samples;
  std::function<void(RequestedPrecisionMap const &, CountType count,
                     std::vector<IndividualEquilibrationCheckResult> &eq, CountType &n_stats,
                     std::vector<BasicStatistics> &stats)>
      check;
};

/// CompletionCheck.hh:175-376
class CompletionCheck {
 public:
  explicit CompletionCheck(CompletionCheckParams params)
      : m_params(std::move(params)), m_n_begin_linear(m_params.find_n_begin_linear()) {
    m_results.params = m_params;
    m_results.is_complete = false;
    if (m_params.equilibration_check_f == nullptr)
      throw std::runtime_error(
          "Error constructing CompletionCheck: params.equilibration_check_f == nullptr");
    if (m_params.calc_statistics_f == nullptr)
      throw std::runtime_error(
          "Error constructing CompletionCheck: params.calc_statistics_f == nullptr");
  }
  CompletionCheckParams const &params() const { return m_params; }
  void reset() {
    m_results.full_reset();
    m_n_checks = 0;
    m_last_n_samples = 0;
    m_last_clocktime = 0.0;
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, LogClock &log) {
    return _is_complete(s, w, std::nullopt, std::nullopt, log);
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, CountType count, LogClock &log) {
    return _is_complete(s, w, count, std::nullopt, log);
  }
  bool is_complete_time(SamplerMap const &s, Sampler const &w, TimeType time, LogClock &log) {
    return _is_complete(s, w, std::nullopt, time, log);
  }
  bool is_complete(SamplerMap const &s, Sampler const &w, CountType count, TimeType time,
                   LogClock &log) {
    return _is_complete(s, w, count, time, log);
  }
  CompletionCheckResults const &results() const { return m_results; }
  Index n_checks() const { return m_n_checks; }
  /// sample count of the next scheduled check
  CountType next_check_at() const {
    return m_params.log_spacing ? m_params.sample_check_log(m_n_checks, m_n_begin_linear)
                                : m_params.sample_check_linear(m_n_checks);
  }
  void set_device_series(std::shared_ptr<DeviceSeriesCheck> d) { m_device = std::move(d); }

  /// Pass-granular drivers: the pass count at which is_complete could next
  /// change anything (a cutoff on count, or the sample count reaching the next
  /// scheduled check / a sample cutoff).  Calling is_complete earlier is
  /// harmless; this only lets a device loop run several passes between calls.
  /// Returns `n_pass + 1` when a clock-based cutoff is set.
  /// `extra_checks`: number of scheduled checks assumed to have been made by then (for a
  /// driver that plans one block ahead of the decision).
  CountType next_decision_pass(CountType n_pass, CountType n_samples, CountType sample_period,
                               Index extra_checks = 0) const {
    CutoffCheckParams const &c = m_params.cutoff_params;
    if (c.min_clocktime || c.max_clocktime || c.min_time || c.max_time) return n_pass + 1;
    CountType best = std::numeric_limits<CountType>::max();
    auto consider = [&](CountType pass) {
      if (pass > n_pass) best = std::min(best, pass);
    };
    if (c.max_count) consider(*c.max_count);
    if (c.min_count) consider(*c.min_count);
    CountType s_target = std::numeric_limits<CountType>::max();
    if (m_params.requested_precision.size()) {
      CountType check_at = m_params.log_spacing
                               ? m_params.sample_check_log(m_n_checks + extra_checks, m_n_begin_linear)
                               : m_params.sample_check_linear(m_n_checks + extra_checks);
      s_target = std::max<CountType>(check_at, c.min_sample ? *c.min_sample : 0);
    }
    if (c.max_sample) s_target = std::min(s_target, std::max<CountType>(*c.max_sample, c.min_sample ? *c.min_sample : 0));
    if (s_target != std::numeric_limits<CountType>::max()) {
      if (s_target <= n_samples) s_target = n_samples + 1;
      consider(s_target * sample_period);
    }
    if (best == std::numeric_limits<CountType>::max()) {
      // nothing can ever stop this run except the caller: advance sample by sample
      best = (n_samples + 1) * sample_period;
      if (best <= n_pass) best = n_pass + 1;
    }
    return best;
  }

 private:
  bool _is_complete(SamplerMap const &samplers, Sampler const &sample_weight,
                    std::optional<CountType> count, std::optional<TimeType> time, LogClock &log) {
    CountType n_samples = m_device ? m_device->n_samples() : get_n_samples(samplers);
    TimeType clocktime = m_last_clocktime;
    if (n_samples != m_last_n_samples) {
      clocktime = log.time_s();
      m_last_n_samples = n_samples;
      m_last_clocktime = clocktime;
    }
    m_results.partial_reset(count, time, clocktime, n_samples);
    m_results.has_all_minimums_met =
        all_minimums_met(m_params.cutoff_params, count, time, n_samples, clocktime);
    if (!m_results.has_all_minimums_met) return false;
    m_results.has_any_maximum_met =
        any_maximum_met(m_params.cutoff_params, count, time, n_samples, clocktime);
    if (m_results.has_any_maximum_met) {
      m_results.is_complete = true;
      if (!(m_results.n_samples_at_convergence_check.has_value() &&
            n_samples == *m_results.n_samples_at_convergence_check))
        _check_convergence(samplers, sample_weight, n_samples);
      return true;
    }
    Index check_at = m_params.log_spacing ? m_params.sample_check_log(m_n_checks, m_n_begin_linear)
                                          : m_params.sample_check_linear(m_n_checks);
    if (n_samples >= check_at) {
      m_n_checks += 1;
      _check_convergence(samplers, sample_weight, n_samples);
    }
    if (m_results.convergence_check_results.all_converged) m_results.is_complete = true;
    return m_results.is_complete;
  }
  void _check_convergence(SamplerMap const &samplers, Sampler const &sample_weight,
                          CountType n_samples) {
    if (!m_params.requested_precision.size()) return;
    m_results.n_samples_at_convergence_check = n_samples;
    if (m_device) {
      _check_convergence_on_device(n_samples);
      return;
    }
    m_results.equilibration_check_results =
        equilibration_check(m_params.equilibration_check_f, m_params.requested_precision, samplers,
                            sample_weight, false);
    if (m_results.equilibration_check_results.all_equilibrated) {
      m_results.convergence_check_results = convergence_check(
          samplers, sample_weight, m_params.requested_precision,
          m_results.equilibration_check_results.N_samples_for_all_to_equilibrate,
          m_params.calc_statistics_f);
    } else {
      m_results.convergence_check_results = ConvergenceCheckResults();
    }
  }
  /// Same results as the host path (equilibration_check with check_all = false,
  /// EquilibrationCheck.cc:203-224, then convergence_check, ConvergenceCheck.hh:139-184),
  /// from one device call.
  void _check_convergence_on_device(CountType n_samples) {
    std::vector<IndividualEquilibrationCheckResult> eq;
    std::vector<BasicStatistics> stats;
    CountType n_stats = 0;
    m_device->check(m_params.requested_precision, n_samples, eq, n_stats, stats);
    EquilibrationCheckResults er;
    er.all_equilibrated = true;
    size_t i = 0;
    for (auto const &p : m_params.requested_precision) {
      IndividualEquilibrationCheckResult const &current = eq.at(i++);
      er.N_samples_for_all_to_equilibrate =
          std::max(er.N_samples_for_all_to_equilibrate, current.N_samples_for_equilibration);
      er.all_equilibrated &= current.is_equilibrated;
      er.individual_results.emplace(p.first, current);
      if (!er.all_equilibrated) break;
    }
    m_results.equilibration_check_results = er;
    ConvergenceCheckResults cr;
    if (er.all_equilibrated && er.N_samples_for_all_to_equilibrate < n_samples) {
      cr.N_samples_for_statistics = n_samples - er.N_samples_for_all_to_equilibrate;
      if (cr.N_samples_for_statistics != n_stats)
        throw std::runtime_error("Error in CompletionCheck: device statistics window mismatch");
      cr.all_converged = true;
      i = 0;
      for (auto const &p : m_params.requested_precision) {
        IndividualConvergenceCheckResult current = convergence_check(stats.at(i++), p.second);
        cr.all_converged &= current.is_converged;
        cr.individual_results.emplace(p.first, current);
      }
    }
    m_results.convergence_check_results = cr;
  }
  CompletionCheckParams m_params;
  CompletionCheckResults m_results;
  std::shared_ptr<DeviceSeriesCheck> m_device;
  Index m_n_checks = 0;
  Index m_n_begin_linear = 0;
  Index m_last_n_samples = 0;
  double m_last_clocktime = 0.0;
};

// ---------------------------------------------------------------------------
// BasicOccupationMetropolisData
// (include/casm/monte/methods/basic_occupation_metropolis.hh:19-115)
// ---------------------------------------------------------------------------
struct BasicOccupationMetropolisData {
  BasicOccupationMetropolisData(StateSamplingFunctionMap const &_sampling_functions,
                                CountType _n_steps_per_pass,
                                CompletionCheckParams const &_completion_check_params)
      : sampling_functions(_sampling_functions), sample_weight(std::vector<Index>{}),
        n_steps_per_pass(_n_steps_per_pass), completion_check(_completion_check_params) {
    for (auto const &pair : sampling_functions) {
      auto const &f = pair.second;
      samplers.emplace(f.name, std::make_shared<Sampler>(f.shape, f.component_names));
    }
  }
  StateSamplingFunctionMap sampling_functions;
  SamplerMap samplers;
  Sampler sample_weight;
  CountType n_pass = 0;
  CountType n_steps_per_pass;
  BigCountType n_accept = 0;
  BigCountType n_reject = 0;
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
    for (auto &pair : samplers) pair.second->clear();
    sample_weight.clear();
    n_pass = 0;
    n_accept = 0;
    n_reject = 0;
    completion_check.reset();
  }
};

// ---------------------------------------------------------------------------
// JSON text of the results, same keys as
// include/casm/monte/checks/io/json/CompletionCheck_json_io.hh:421-438,
// src/casm/monte/checks/io/json/EquilibrationCheck_json_io.cc:15-47 and
// include/casm/monte/checks/io/json/ConvergenceCheck_json_io.hh:35-62
// (what default_finish_write_status writes to status.json)
// ---------------------------------------------------------------------------
inline std::string json_number(double v) {
  if (!std::isfinite(v)) return "null";
  std::ostringstream ss;
  ss.precision(17);
  ss << v;
  return ss.str();
}
inline std::string json_string(std::string const &s) {
  std::string out = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') out += '\\';
    out += c;
  }
  return out + "\"";
}
inline std::string to_json_text(RequestedPrecision const &r) {
  std::string s = "{";
  bool first = true;
  if (r.abs_convergence_is_required) {
    s += "\"abs_precision\": " + json_number(r.abs_precision);
    first = false;
  }
  if (r.rel_convergence_is_required) s += std::string(first ? "" : ", ") + "\"rel_precision\": " + json_number(r.rel_precision);
  return s + "}";
}
inline std::string to_json_text(EquilibrationCheckResults const &v) {
  std::ostringstream s;
  s << "{\"all_equilibrated\": " << (v.all_equilibrated ? "true" : "false") << ", ";
  if (v.all_equilibrated)
    s << "\"N_samples_for_all_to_equilibrate\": " << v.N_samples_for_all_to_equilibrate;
  else
    s << "\"N_samples_for_equilibration\": \"did_not_equilibrate\"";
  s << ", \"individual_results\": [";
  bool first = true;
  for (auto const &p : v.individual_results) {
    s << (first ? "" : ", ") << "{\"is_equilibrated\": " << (p.second.is_equilibrated ? "true" : "false")
      << ", \"N_samples_for_equilibration\": ";
    if (p.second.is_equilibrated) s << p.second.N_samples_for_equilibration;
    else s << "\"did_not_equilibrate\"";
    s << ", \"sampler_name\": " << json_string(p.first.sampler_name) << ", \"component_name\": "
      << json_string(p.first.component_name) << ", \"component_index\": " << p.first.component_index << "}";
    first = false;
  }
  s << "]}";
  return s.str();
}
inline std::string to_json_text(ConvergenceCheckResults const &v) {
  std::ostringstream s;
  s << "{\"all_converged\": " << (v.all_converged ? "true" : "false")
    << ", \"N_samples_for_statistics\": " << v.N_samples_for_statistics << ", \"individual_results\": [";
  bool first = true;
  for (auto const &p : v.individual_results) {
    s << (first ? "" : ", ") << "{\"is_converged\": " << (p.second.is_converged ? "true" : "false")
      << ", \"requested_precision\": " << to_json_text(p.second.requested_precision)
      << ", \"stats\": {\"mean\": " << json_number(p.second.stats.mean) << ", \"calculated_precision\": "
      << json_number(p.second.stats.calculated_precision) << "}, \"sampler_name\": "
      << json_string(p.first.sampler_name) << ", \"component_name\": " << json_string(p.first.component_name)
      << ", \"component_index\": " << p.first.component_index << "}";
    first = false;
  }
  s << "]}";
  return s.str();
}
inline std::string to_json_text(CompletionCheckResults const &v) {
  std::ostringstream s;
  s << "{\"has_all_minimums_met\": " << (v.has_all_minimums_met ? "true" : "false")
    << ", \"has_any_maximum_met\": " << (v.has_any_maximum_met ? "true" : "false") << ", \"count\": ";
  if (v.count.has_value()) s << *v.count; else s << "null";
  s << ", \"time\": " << (v.time.has_value() ? json_number(*v.time) : std::string("null"))
    << ", \"clocktime\": " << json_number(v.clocktime) << ", \"n_samples\": " << v.n_samples
    << ", \"is_complete\": " << (v.is_complete ? "true" : "false");
  if (v.n_samples_at_convergence_check.has_value())
    s << ", \"n_samples_at_convergence_check\": " << *v.n_samples_at_convergence_check
      << ", \"equilibration_check_results\": " << to_json_text(v.equilibration_check_results)
      << ", \"convergence_check_results\": " << to_json_text(v.convergence_check_results);
  s << "}";
  return s.str();
}
/// basic_occupation_metropolis.hh:168-180
inline std::string to_json_text(BasicOccupationMetropolisData const &d) {
  std::ostringstream s;
  s << "{\"completion_check_results\": " << to_json_text(d.completion_check.results())
    << ", \"n_pass\": " << d.n_pass << ", \"n_steps_per_pass\": " << d.n_steps_per_pass
    << ", \"n_accept\": " << static_cast<long>(d.n_accept) << ", \"n_reject\": " << static_cast<long>(d.n_reject)
    << ", \"acceptance_rate\": " << json_number(d.acceptance_rate()) << ", \"rejection_rate\": "
    << json_number(d.rejection_rate()) << "}";
  return s.str();
}
/// basic_occupation_metropolis.hh:248-260: results JSON to the log file, lap restarted
inline void default_finish_write_status(BasicOccupationMetropolisData const &data, MethodLog &method_log) {
  method_log.reset();
  if (method_log.fout) {
    (*method_log.fout) << to_json_text(data.completion_check.results()) << std::endl;
    method_log.fout->flush();
  }
  method_log.log.begin_lap();
}

/// basic_occupation_metropolis.hh:224-241
inline void default_write_run_status(BasicOccupationMetropolisData const &data,
                                     MethodLog &method_log, std::ostream &sout) {
  double steps = static_cast<double>(data.n_pass) * static_cast<double>(data.n_steps_per_pass);
  double time_s = method_log.log.time_s();
  sout << "Passes=" << data.n_pass << ", ";
  sout << "Samples=" << get_n_samples(data.samplers) << ", ";
  sout << "ClockTime(s)=" << time_s << ", ";
  sout << "Steps/Second=" << steps / time_s << ", ";
  sout << "Seconds/Step=" << time_s / steps << std::endl;
}

// ---------------------------------------------------------------------------
// Semi-grand canonical pieces
// (include/casm/monte/ising_cpp/basic_semigrand_canonical.hh)
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
      throw std::runtime_error("Missing required condition: \"exchange_potential\"");
    return SemiGrandCanonicalConditions(values.scalar_values.at("temperature"),
                                        values.vector_values.at("exchange_potential"));
  }
  ValueMap to_values() const {
    ValueMap v;
    v.scalar_values["temperature"] = temperature;
    v.vector_values["exchange_potential"] = exchange_potential;
    return v;
  }
};

inline double dot1(std::vector<double> const &a, std::vector<double> const &b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size() && i < b.size(); ++i) s += a[i] * b[i];
  return s;
}

class SemiGrandCanonicalPotential {
 public:
  typedef IsingSystem system_type;
  typedef IsingState state_type;
  explicit SemiGrandCanonicalPotential(std::shared_ptr<system_type> _system)
      : system(throw_if_null(
            _system, "Error constructing SemiGrandCanonicalPotential: _system==nullptr")),
        state(nullptr), conditions(nullptr),
        formation_energy_calculator(system->formation_energy_calculator),
        param_composition_calculator(system->param_composition_calculator) {}
  std::shared_ptr<system_type> system;
  state_type const *state;
  std::shared_ptr<SemiGrandCanonicalConditions> conditions;
  IsingFormationEnergy formation_energy_calculator;
  IsingParamComposition param_composition_calculator;

  void set_state(state_type const *_state, std::shared_ptr<SemiGrandCanonicalConditions> _conditions) {
    state = throw_if_null(
        _state, "Error in SemiGrandCanonicalPotential::set_state: _state is nullptr");
    conditions = throw_if_null(
        _conditions, "Error in SemiGrandCanonicalPotential::set_state: _conditions is nullptr");
    formation_energy_calculator.set_state(_state);
    param_composition_calculator.set_state(_state);
  }
  double per_supercell() {
    return formation_energy_calculator.per_supercell() -
           dot1(conditions->exchange_potential, param_composition_calculator.per_supercell());
  }
  double per_unitcell() { return per_supercell() / state->configuration.n_unitcells; }
  double occ_delta_per_supercell(std::vector<Index> const &linear_site_index,
                                 std::vector<int> const &new_occ) const {
    double dE_f = formation_energy_calculator.occ_delta_per_supercell(linear_site_index, new_occ);
    std::vector<double> Ndx =
        param_composition_calculator.occ_delta_per_supercell(linear_site_index, new_occ);
    return dE_f - dot1(conditions->exchange_potential, Ndx);
  }
  double occ_delta_per_supercell(OccEvent const &e) const {
    return occ_delta_per_supercell(e.linear_site_index, e.new_occ);
  }
};

template <typename EngineType = default_engine_type>
class SemiGrandCanonicalEventGenerator {
 public:
  typedef IsingState state_type;
  typedef EngineType engine_type;
  typedef RandomNumberGenerator<engine_type> random_number_generator_type;
  SemiGrandCanonicalEventGenerator() : state(nullptr), m_max_linear_site_index(0) {
    occ_event.linear_site_index.assign(1, 0);
    occ_event.new_occ.assign(1, 1);
  }
  state_type *state;
  OccEvent occ_event;
  void set_state(state_type *_state) {
    state = throw_if_null(
        _state, "Error in SemiGrandCanonicalEventGenerator::set_state: _state==nullptr");
    m_max_linear_site_index = state->configuration.n_sites - 1;
  }
  OccEvent const &propose(random_number_generator_type &rng) {
    occ_event.linear_site_index[0] = rng.random_int(m_max_linear_site_index);
    occ_event.new_occ[0] = -state->configuration.occ(occ_event.linear_site_index[0]);
    return occ_event;
  }
  void apply(OccEvent const &e) {
    state->configuration.set_occ(e.linear_site_index[0], e.new_occ[0]);
  }

 private:
  Index m_max_linear_site_index;
};

typedef BasicOccupationMetropolisData SemiGrandCanonicalData;

/// methods/metropolis.hh:26-35 (host form, for host-driven loops)
template <typename GeneratorType>
bool metropolis_acceptance(double delta_potential_energy, double beta, GeneratorType &rng) {
  if (delta_potential_energy < 0.0) return true;
  double rand = rng.random_real(1.0);
  double prob = std::exp(-delta_potential_energy * beta);
  return rand < prob;
}

/// basic_semigrand_canonical.hh:324-470
class SemiGrandCanonicalCalculator {
 public:
  typedef IsingSystem system_type;
  typedef IsingState state_type;
  typedef SemiGrandCanonicalPotential potential_type;
  typedef SemiGrandCanonicalEventGenerator<default_engine_type> event_generator_type;
  typedef default_engine_type engine_type;
  typedef std::function<void(SemiGrandCanonicalCalculator const &, MethodLog &)> write_status_type;

  explicit SemiGrandCanonicalCalculator(std::shared_ptr<system_type> _system)
      : system(throw_if_null(
            _system, "Error constructing SemiGrandCanonicalCalculator: _system==nullptr")),
        state(nullptr), conditions(nullptr), potential(_system),
        formation_energy_calculator(&potential.formation_energy_calculator),
        param_composition_calculator(&potential.param_composition_calculator) {}

  std::shared_ptr<system_type> system;
  state_type *state;
  std::shared_ptr<SemiGrandCanonicalConditions> conditions;
  potential_type potential;
  IsingFormationEnergy *formation_energy_calculator;
  IsingParamComposition *param_composition_calculator;
  std::shared_ptr<SemiGrandCanonicalData> data;

  /// "auto" | "checkerboard" | "serial_reference"  (not in the reference)
  std::string update_mode = "auto";
  /// name of the kernel that ran the last `run` (introspection)
  std::string last_kernel;
  /// (not in the reference) Sweep on behind every completion check: the check that is due goes
  /// into the stream (cmg_series_check_prefetch), the next block of passes is enqueued behind a
  /// restore point (cmg_mark) before the verdict is known, and the verdict is collected while
  /// that block runs; a "complete" verdict -- or a decision that asks for a different block --
  /// rolls the block back (cmg_rollback).  Results are identical either way
  /// (test_run_with_overlapped_checks_gives_the_same_results; the whole GPU test suite also
  /// passes with CMG_OVERLAP_CHECKS=1).  The device never idles between a block and the host's
  /// decision; costs one device-to-device copy of the occupation per check.
  /// Measured on 4096^2 with a check every 100 passes it changes nothing (0.124 s either way:
  /// what a check costs there is its kernels, not the host round trip), so it stays an opt-in
  /// for runs whose blocks are short: this attribute, or CMG_OVERLAP_CHECKS=1 in the environment.
  bool overlap_checks = default_overlap_checks();
  static bool default_overlap_checks() {
    const char *e = std::getenv("CMG_OVERLAP_CHECKS");
    return e && e[0] == '1';
  }

  /// `json_sample_hook`, if set, is called after every sample with the host
  /// mirror of the configuration current (JSON samplers live in the binding).
  std::shared_ptr<SemiGrandCanonicalData> run(
      state_type &_state, StateSamplingFunctionMap const &sampling_functions,
      CompletionCheckParams const &completion_check_params, event_generator_type event_generator,
      int sample_period = 1, std::optional<MethodLog> method_log = std::nullopt,
      std::shared_ptr<engine_type> random_engine = nullptr,
      write_status_type write_status_f = nullptr, std::function<void()> json_sample_hook = nullptr) {
    if (sample_period < 1) throw std::runtime_error("Error in run: sample_period < 1");
    // ### setup, as basic_semigrand_canonical.hh:427-462
    state = &_state;
    conditions = std::make_shared<SemiGrandCanonicalConditions>(
        SemiGrandCanonicalConditions::from_values(state->conditions));
    if (conditions->exchange_potential.size() != 1)
      throw std::runtime_error("Error in run: exchange_potential must have 1 component");
    const double temperature = conditions->temperature;
    const double mu = conditions->exchange_potential[0];
    CountType n_steps_per_pass = state->configuration.n_variable_sites;
    potential.set_state(state, conditions);
    event_generator.set_state(state);
    data = std::make_shared<SemiGrandCanonicalData>(sampling_functions, n_steps_per_pass,
                                                   completion_check_params);

    // ### method log / clock, as basic_occupation_metropolis.hh:366-373
    if (!method_log.has_value()) {
      method_log = MethodLog();
      method_log->logfile_path = "status.json";
      method_log->log_frequency = 600.0;
    }
    method_log->log.restart_clock();
    method_log->log.begin_lap();

    // ### device setup
    IsingConfiguration &config = state->configuration;
    DeviceLattice &dev = config.device();
    cmg_context *ctx = dev.ctx();
    const double J = potential.formation_energy_calculator.J;
    dev.check(cmg_set_model(ctx, J, potential.formation_energy_calculator.lattice_type));
    dev.check(cmg_set_conditions(ctx, 0, temperature, mu));
    dev.check(cmg_reset_counters(ctx));
    dev.check(cmg_clear_samples(ctx));

    bool even = true;
    for (int s : config.shape) even = even && (s % 2 == 0);
    int mode;
    if (update_mode == "serial_reference") mode = CMG_MODE_SERIAL_REFERENCE;
    else if (update_mode == "checkerboard") mode = CMG_MODE_CHECKERBOARD;
    else if (update_mode == "auto") mode = even ? CMG_MODE_CHECKERBOARD : CMG_MODE_SERIAL_REFERENCE;
    else throw std::runtime_error("Error in run: unknown update_mode '" + update_mode + "'");

    // random numbers: a null engine is seeded from std::random_device, as
    // RandomNumberGenerator does (RandomNumberGenerator.hh:23-27)
    RandomNumberGenerator<engine_type> rng(random_engine);
    if (mode == CMG_MODE_SERIAL_REFERENCE) {
      uint64_t words[312];
      int pos = 0;
      engine_to_words(*rng.engine, words, &pos);
      dev.check(cmg_set_mt19937_64_state(ctx, 0, words, pos));
    } else {
      dev.check(cmg_seed_philox(ctx, (*rng.engine)()));  // one draw seeds the Philox key
      dev.check(cmg_set_pass_counter(ctx, 0));
    }

    // which sampling functions can the device loop evaluate itself?
    bool all_builtin = true;
    bool needs_potential_in_nonlist_form =
        !potential.formation_energy_calculator.use_nlist() && config.shape.size() == 2;
    // the row/column energy form (model.hh:273-285) is sampled on the device in
    // checkerboard mode where the library supports it; otherwise through the
    // calculators on the downloaded state
    dev.check(cmg_set_energy_form(ctx, 1));
    bool nonlist_on_device = false;
    if (needs_potential_in_nonlist_form && mode == CMG_MODE_CHECKERBOARD)
      nonlist_on_device = cmg_set_energy_form(ctx, 0) == CMG_OK;
    for (auto const &pair : data->sampling_functions) {
      auto const &f = pair.second;
      if (f.builtin < 0 || f.builtin_owner != static_cast<void const *>(this)) all_builtin = false;
      if (needs_potential_in_nonlist_form && !nonlist_on_device &&
          f.builtin != CMG_Q_PARAM_COMPOSITION)
        all_builtin = false;
    }
    const bool host_state_each_sample = !all_builtin || static_cast<bool>(json_sample_hook);
    const bool device_samples = all_builtin;

    // Can the scheduled checks run on the device-resident series?  Yes if every
    // requested component is one of the built-in observables with an absolute
    // precision only, and the check functions are this library's defaults.
    bool device_checks = device_samples && !host_state_each_sample;
    std::vector<int> check_quantity;
    std::vector<double> check_abs;
    double check_confidence = 0.95;
    if (device_checks) {
      typedef IndividualEquilibrationCheckResult (*eq_fn)(std::vector<double> const &,
                                                          std::vector<double> const &, RequestedPrecision);
      eq_fn const *eq_target = completion_check_params.equilibration_check_f.target<eq_fn>();
      BasicStatisticsCalculator const *calc =
          completion_check_params.calc_statistics_f.target<BasicStatisticsCalculator>();
      device_checks = eq_target && *eq_target == &default_equilibration_check && calc != nullptr &&
                      completion_check_params.requested_precision.size() <= 3;
      if (calc) check_confidence = calc->confidence;
      for (auto const &p : completion_check_params.requested_precision) {
        auto it = data->sampling_functions.find(p.first.sampler_name);
        if (it == data->sampling_functions.end() || p.first.component_index != 0 ||
            !p.second.abs_convergence_is_required || p.second.rel_convergence_is_required) {
          device_checks = false;
          break;
        }
        check_quantity.push_back(it->second.builtin);
        check_abs.push_back(p.second.abs_precision);
      }
    }
    // samples of the passes that are decided (a speculative block may be in flight beyond them);
    // negative: whatever the device has
    auto committed_samples = std::make_shared<CountType>(-1);
    if (device_checks) {
      auto hook = std::make_shared<DeviceSeriesCheck>();
      hook->n_samples = [ctx, committed_samples]() {
        if (*committed_samples >= 0) return *committed_samples;
        int64_t n = 0;
        cmg_check(cmg_n_samples(ctx, &n), ctx);
        return static_cast<CountType>(n);
      };
      hook->check = [ctx, check_quantity, check_abs, check_confidence](
                        RequestedPrecisionMap const &req, CountType count,
                        std::vector<IndividualEquilibrationCheckResult> &eq, CountType &n_stats,
                        std::vector<BasicStatistics> &stats) {
        const int n = static_cast<int>(req.size());
        if (n == 0) return;
        int is_eq[3] = {0, 0, 0};
        int64_t n_eq[3] = {0, 0, 0}, ns = 0;
        double mean[3] = {0, 0, 0}, prec[3] = {0, 0, 0};
        cmg_check(cmg_series_check(ctx, 0, n, check_quantity.data(), check_abs.data(),
                                   static_cast<int64_t>(count), check_confidence, is_eq, n_eq, &ns,
                                   mean, prec),
                  ctx);
        eq.resize(n);
        stats.resize(n);
        for (int i = 0; i < n; ++i) {
          eq[i].is_equilibrated = is_eq[i] != 0;
          eq[i].N_samples_for_equilibration = static_cast<CountType>(n_eq[i]);
          stats[i].mean = mean[i];
          stats[i].calculated_precision = prec[i];
        }
        n_stats = static_cast<CountType>(ns);
      };
      data->completion_check.set_device_series(hook);
    }

    CountType n_pass_dev = 0;       // passes done on the device
    CountType n_fetched = 0;        // device samples already appended to the host samplers
    auto fetch_device_samples = [&]() {
      if (!device_samples) return;
      int64_t n_dev = 0;
      dev.check(cmg_n_samples(ctx, &n_dev));
      if (n_dev <= n_fetched) return;
      std::vector<double> buf(static_cast<size_t>(n_dev - n_fetched));
      for (auto const &pair : data->sampling_functions) {
        auto const &f = pair.second;
        dev.check(cmg_read_samples(ctx, 0, f.builtin, n_fetched, n_dev - n_fetched, buf.data()));
        data->samplers.at(f.name)->append_column(buf.data(), static_cast<CountType>(buf.size()));
      }
      n_fetched = static_cast<CountType>(n_dev);
    };
    auto refresh_counters = [&]() {
      int64_t np = 0, na = 0, nr = 0;
      dev.check(cmg_counters(ctx, 0, &np, &na, &nr));
      data->n_pass = static_cast<CountType>(np);
      data->n_accept = na;
      data->n_reject = nr;
    };

    auto current_n_samples = [&]() -> CountType {
      if (!device_checks) return get_n_samples(data->samplers);
      if (*committed_samples >= 0) return *committed_samples;
      int64_t n = 0;
      dev.check(cmg_n_samples(ctx, &n));
      return static_cast<CountType>(n);
    };
    // Sweeping on while a check is evaluated: with the checks on the device series, the block
    // after the next decision is enqueued BEFORE that decision is taken (cmg_mark keeps a
    // restore point), the check runs on a second stream next to it, and a "complete" verdict --
    // or a decision that asks for a different block -- rolls the speculative block back.  The
    // results are those of the loop that waits for every decision.
    const bool can_speculate = overlap_checks && device_checks && mode == CMG_MODE_CHECKERBOARD && even && !nonlist_on_device;
    bool have_spec = false;
    CountType spec_n_run = 0;
    auto drop_speculation = [&]() {
      if (!have_spec) return;
      dev.check(cmg_rollback(ctx));
      have_spec = false;
      *committed_samples = -1;
    };

    // ### main loop at pass granularity: is_complete is consulted
    // at every pass boundary at which its answer could change, and re-consulted
    // while it is still catching up on scheduled checks.
    while (true) {
      bool done = false;
      while (true) {
        Index checks_before = data->completion_check.n_checks();
        done = data->completion_check.is_complete(data->samplers, data->sample_weight,
                                                  data->n_pass, method_log->log);
        if (done || data->completion_check.n_checks() == checks_before) break;
      }
      if (done) {
        drop_speculation();
        break;
      }

      CountType target;
      if (host_state_each_sample) {
        // stop at the next sample (or earlier at a count cutoff)
        target = (n_pass_dev / sample_period + 1) * sample_period;
        CountType nd = data->completion_check.next_decision_pass(
            n_pass_dev, current_n_samples(), sample_period);
        target = std::min(target, nd);
      } else {
        target = data->completion_check.next_decision_pass(
            n_pass_dev, current_n_samples(), sample_period);
      }
      CountType n_run = std::max<CountType>(1, target - n_pass_dev);
      if (have_spec && spec_n_run == n_run) {
        have_spec = false;  // the block in flight is the block asked for
        *committed_samples = -1;
      } else {
        drop_speculation();
        dev.check(cmg_run_passes(ctx, n_run, mode, device_samples ? sample_period : 0));
      }
      config.mark_device_modified();
      n_pass_dev += n_run;
      data->n_pass = n_pass_dev;

      const bool sample_due = (n_pass_dev % sample_period) == 0;
      // with the checks on the device series the samplers are filled once, at the end
      if (device_samples && !device_checks) fetch_device_samples();
      if (sample_due && host_state_each_sample) {
        refresh_counters();
        config.pull();
        if (!device_samples)
          for (auto const &pair : data->sampling_functions) {
            auto const &f = pair.second;
            data->samplers.at(f.name)->push_back(f());
          }
        if (json_sample_hook) json_sample_hook();
      }
      const bool status_due = sample_due && write_status_f && method_log->log_frequency.has_value() &&
                              method_log->log.lap_time() >= method_log->log_frequency.value();
      if (status_due) {
        refresh_counters();
        fetch_device_samples();
        write_status_f(*this, *method_log);
      } else if (can_speculate) {
        // the block the loop will ask for if the decision at n_pass_dev is "go on"
        const CountType s_now = current_n_samples();
        CompletionCheck const &cc = data->completion_check;
        const Index extra = (completion_check_params.requested_precision.size() && s_now >= cc.next_check_at()) ? 1 : 0;
        const CountType t2 = cc.next_decision_pass(n_pass_dev, s_now, sample_period, extra);
        spec_n_run = std::max<CountType>(1, t2 - n_pass_dev);
        // the check the loop is about to ask for goes into the stream AHEAD of the speculative
        // block; its verdict is collected (cmg_series_check, same arguments) while the block runs
        if (extra && !check_quantity.empty())
          dev.check(cmg_series_check_prefetch(ctx, 0, static_cast<int>(check_quantity.size()), check_quantity.data(),
                                              check_abs.data(), static_cast<int64_t>(s_now), check_confidence));
        dev.check(cmg_mark(ctx));
        dev.check(cmg_run_passes(ctx, spec_n_run, mode, sample_period));
        *committed_samples = s_now;
        have_spec = true;
      }
    }
    fetch_device_samples();
    data->completion_check.set_device_series(nullptr);

    // ### finish: counters; the final occupation is visible in the caller's state through
    // every accessor of its configuration (the host mirror is refreshed on first use, 4 B
    // per site are not moved unless somebody looks); engine advanced exactly as the
    // reference would leave it (serial mode)
    refresh_counters();
    if (mode == CMG_MODE_SERIAL_REFERENCE) {
      uint64_t words[312];
      int pos = 0;
      dev.check(cmg_get_mt19937_64_state(ctx, 0, words, &pos));
      words_to_engine(words, pos, *rng.engine);
    }
    last_kernel = cmg_kernel_variant(ctx);
    if (write_status_f) write_status_f(*this, *method_log);
    return data;
  }
};

/// basic_semigrand_canonical.hh:245-264
inline void default_write_status(SemiGrandCanonicalCalculator const &mc_calculator,
                                 MethodLog &method_log) {
  std::ostream &sout = std::cout;
  default_write_run_status(*mc_calculator.data, method_log, sout);
  sout << "  ParametricComposition=" << mc_calculator.param_composition_calculator->per_unitcell()[0]
       << ", FormationEnergy=" << mc_calculator.formation_energy_calculator->per_unitcell()
       << std::endl;
  auto const &results = mc_calculator.data->completion_check.results();
  sout << "  AllEquilibrated=" << results.equilibration_check_results.all_equilibrated << std::endl;
  if (results.equilibration_check_results.all_equilibrated)
    sout << "  AllConverged=" << results.convergence_check_results.all_converged << std::endl;
  default_finish_write_status(*mc_calculator.data, method_log);
}

/// The generic driver with caller-supplied callbacks
/// (methods/basic_occupation_metropolis.hh:354-425; Python form
/// python/src/monte_methods.cpp:196-263).  It only sequences the callbacks --
/// the arithmetic is in whatever they call (e.g. the device-backed calculators
/// above).  SemiGrandCanonicalCalculator::run does NOT go through this; it
/// drives the device loop directly.
template <typename EngineType = default_engine_type>
void basic_occupation_metropolis(
    BasicOccupationMetropolisData &data, double temperature,
    std::function<double(OccEvent const &)> potential_occ_delta_per_supercell_f,
    std::function<OccEvent const &(RandomNumberGenerator<EngineType> &)> propose_event_f,
    std::function<void(OccEvent const &)> apply_event_f, int sample_period = 1,
    std::optional<MethodLog> method_log = std::nullopt,
    std::shared_ptr<EngineType> random_engine = nullptr,
    std::function<void(BasicOccupationMetropolisData const &, MethodLog &)> write_status_f = nullptr) {
  double beta = 1.0 / (KB * temperature);
  RandomNumberGenerator<EngineType> random_number_generator(random_engine);
  if (!method_log.has_value()) {
    method_log = MethodLog();
    method_log->logfile_path = "status.json";
    method_log->log_frequency = 600.0;
  }
  method_log->log.restart_clock();
  method_log->log.begin_lap();
  Index n_pass_next_sample = sample_period;
  CountType n_step = 0;
  while (!data.completion_check.is_complete(data.samplers, data.sample_weight, data.n_pass,
                                            method_log->log)) {
    OccEvent const &event = propose_event_f(random_number_generator);
    double delta_potential_energy = potential_occ_delta_per_supercell_f(event);
    if (metropolis_acceptance(delta_potential_energy, beta, random_number_generator)) {
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
      for (auto const &pair : data.sampling_functions) {
        auto const &f = pair.second;
        data.samplers.at(f.name)->push_back(f());
      }
      if (write_status_f && method_log->log_frequency.has_value() &&
          method_log->log.lap_time() >= method_log->log_frequency.value())
        write_status_f(data, *method_log);
    }
  }
  if (write_status_f) write_status_f(data, *method_log);
}

/// basic_semigrand_canonical.hh:486-590; tagged so that `run` can sample them
/// on the device instead of calling back
inline StateSamplingFunction make_parametric_composition_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc_calculator) {
  if (mc_calculator == nullptr)
    throw std::runtime_error(
        "Error in parametric_composition sampling function: mc_calculator == nullptr");
  std::vector<Index> shape;
  shape.push_back(mc_calculator->system->param_composition_calculator.n_independent_compositions());
  auto *raw = mc_calculator.get();
  std::weak_ptr<SemiGrandCanonicalCalculator> weak = mc_calculator;  // no ownership cycle
  auto f = [weak]() -> std::vector<double> {
    auto raw = weak.lock();
    if (raw == nullptr)
      throw std::runtime_error(
          "Error in parametric_composition sampling function: mc_calculator == nullptr");
    if (raw->param_composition_calculator->state == nullptr)
      throw std::runtime_error(
          "Error in parametric_composition sampling function: "
          "mc_calculator->param_composition_calculator->state == nullptr");
    return raw->param_composition_calculator->per_unitcell();
  };
  StateSamplingFunction sf("param_composition", "Parametric composition", shape, f);
  sf.builtin = CMG_Q_PARAM_COMPOSITION;
  sf.builtin_owner = raw;
  return sf;
}
inline StateSamplingFunction make_formation_energy_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc_calculator) {
  if (mc_calculator == nullptr)
    throw std::runtime_error(
        "Error in formation_energy sampling function: mc_calculator == nullptr");
  auto *raw = mc_calculator.get();
  std::weak_ptr<SemiGrandCanonicalCalculator> weak = mc_calculator;
  auto f = [weak]() -> std::vector<double> {
    auto raw = weak.lock();
    if (raw == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: mc_calculator == nullptr");
    if (raw->formation_energy_calculator->state == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: "
          "mc_calculator->formation_energy_calculator->state == nullptr");
    return std::vector<double>{raw->formation_energy_calculator->per_unitcell()};
  };
  StateSamplingFunction sf("formation_energy", "Intensive formation energy", {}, f);
  sf.builtin = CMG_Q_FORMATION_ENERGY;
  sf.builtin_owner = raw;
  return sf;
}
inline StateSamplingFunction make_potential_energy_f(
    std::shared_ptr<SemiGrandCanonicalCalculator> mc_calculator) {
  if (mc_calculator == nullptr)
    throw std::runtime_error(
        "Error in formation_energy sampling function: mc_calculator == nullptr");
  auto *raw = mc_calculator.get();
  std::weak_ptr<SemiGrandCanonicalCalculator> weak = mc_calculator;
  auto f = [weak]() -> std::vector<double> {
    auto raw = weak.lock();
    if (raw == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: mc_calculator == nullptr");
    if (raw->potential.state == nullptr)
      throw std::runtime_error(
          "Error in formation_energy sampling function: mc_calculator->potential.state == nullptr");
    return std::vector<double>{raw->potential.per_unitcell()};
  };
  StateSamplingFunction sf("potential_energy", "Intensive potential energy", {}, f);
  sf.builtin = CMG_Q_POTENTIAL_ENERGY;
  sf.builtin_owner = raw;
  return sf;
}

// ---------------------------------------------------------------------------
// Conversions (include/casm/monte/Conversions.hh:43-135; src/casm/monte/Conversions.cc)
//
// The reference builds it from an xtal::BasicStructure; libcasm-xtal is absent
// here, so the primitive cell is given as plain arrays (ConversionsPrim): the
// lattice column matrix, the fractional basis coordinates and the occupant names
// of every sublattice.  Everything the reference computes from those is mirrored:
// l / b / ijk / bijk / unitl / asym conversions for ANY integer transformation
// matrix (snf.hh), Cartesian / fractional site coordinates, and the occ_index <->
// species_index tables (Conversions.cc:147-172, :296-321).  What needs the prim's
// factor group (the default asymmetric unit, Conversions.cc:14-52) is replaced by
// its occupant-order part: sublattices with identical occupant lists share an
// orbit unless b_to_asym / unitl_to_asym is given (the reference's
// make_with_custom_asym / make_with_custom_unitcell constructors).
// Scalar calls are evaluated on the host; the batched forms run on the device.
// ---------------------------------------------------------------------------
struct ConversionsPrim {
  /// lat_column_mat, row-major (a, b, c are the COLUMNS); identity by default
  std::array<double, 9> lat_column_mat{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
  /// fractional coordinates of the basis sites
  std::vector<std::array<double, 3>> basis_frac;
  /// occupant names per sublattice, in occupation-index order (xtal::Prim occ_dof)
  std::vector<std::vector<std::string>> occ_dof;
};

class Conversions {
 public:
  typedef std::array<long, 9> matrix_type;  // row-major 3 x 3

  /// Conversions.cc:63-68 (asymmetric unit from the occupant lists, see above)
  Conversions(ConversionsPrim const &prim, matrix_type const &transformation_matrix_to_super, int device = 0)
      : Conversions(prim, default_species_list(prim), transformation_matrix_to_super,
                    matrix_type{{1, 0, 0, 0, 1, 0, 0, 0, 1}},
                    default_b_to_asym(prim, default_species_list(prim)), device) {}
  /// Conversions.cc:86-92: user specified asymmetric unit with reduced symmetry
  Conversions(ConversionsPrim const &prim, matrix_type const &transformation_matrix_to_super,
              std::vector<Index> const &b_to_asym, int device = 0)
      : Conversions(prim, default_species_list(prim), transformation_matrix_to_super,
                    matrix_type{{1, 0, 0, 0, 1, 0, 0, 0, 1}}, b_to_asym, device) {}
  /// Conversions.cc:118-173: user specified asymmetric unit in a sub-supercell
  Conversions(ConversionsPrim const &prim, std::vector<std::string> const &species_list,
              matrix_type const &transformation_matrix_to_super,
              matrix_type const &unit_transformation_matrix_to_super,
              std::vector<Index> const &unitl_to_asym, int device = 0)
      : m_prim(prim), m_T(transformation_matrix_to_super), m_unit_T(unit_transformation_matrix_to_super),
        m_species(species_list), m_unitl_to_asym(unitl_to_asym), m_device(device) {
    if (prim.occ_dof.empty()) throw std::runtime_error("Conversions: the prim has no basis sites");
    if (m_prim.basis_frac.empty()) m_prim.basis_frac.assign(prim.occ_dof.size(), {{0.0, 0.0, 0.0}});
    if (m_prim.basis_frac.size() != prim.occ_dof.size())
      throw std::runtime_error("Conversions: basis_frac and occ_dof differ in size");
    const int64_t nb = static_cast<int64_t>(prim.occ_dof.size());
    int64_t t[9], u[9];
    for (int i = 0; i < 9; ++i) {
      t[i] = m_T[i];
      u[i] = m_unit_T[i];
    }
    m_l_conv = SiteIndexConverter(Mat3l::from_row_major(t), nb);
    m_unit_conv = SiteIndexConverter(Mat3l::from_row_major(u), nb);
    {  // U must tile into S: S = U * T' with T' integer  <=>  adj(U) * T divisible by det(U)
      Mat3l q = mul(adjugate(m_unit_conv.T), m_l_conv.T);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          if (q.a[r][c] % m_unit_conv.detT != 0)
            throw std::runtime_error("Conversions: the unit supercell does not tile the supercell");
    }
    if (static_cast<int64_t>(m_unitl_to_asym.size()) != m_unit_conv.total_sites())
      throw std::runtime_error("Conversions: unitl_to_asym.size() != number of sites of the unit supercell");
    for (int b = 0; b < nb; ++b) {
      auto const &f = m_prim.basis_frac[b];
      std::array<double, 3> c;
      for (int r = 0; r < 3; ++r)
        c[r] = m_prim.lat_column_mat[3 * r] * f[0] + m_prim.lat_column_mat[3 * r + 1] * f[1] +
               m_prim.lat_column_mat[3 * r + 2] * f[2];
      m_basis_cart.push_back(c);
    }
    // Conversions.cc:133-146
    m_Nasym = *std::max_element(m_unitl_to_asym.begin(), m_unitl_to_asym.end()) + 1;
    m_asym_to_unitl.resize(m_Nasym);
    m_asym_to_b.resize(m_Nasym);
    for (Index unitl = 0; unitl < m_unit_conv.total_sites(); ++unitl) {
      Index asym = m_unitl_to_asym[unitl];
      if (asym < 0) throw std::runtime_error("Conversions: negative asymmetric unit index");
      m_asym_to_unitl[asym].insert(unitl);
      m_asym_to_b[asym].insert(unitl_to_b(unitl));
    }
    // Conversions.cc:148-172: [b][occ] -> species and its inverse (size if not allowed)
    std::vector<std::vector<Index>> conv = make_index_converter(prim, m_species);
    std::vector<std::vector<Index>> conv_inv;
    for (auto const &row : conv) {
      std::vector<Index> occ_indices(m_species.size(), static_cast<Index>(row.size()));
      Index occ_index = 0;
      for (Index species_index : row) occ_indices[species_index] = occ_index++;
      conv_inv.push_back(occ_indices);
    }
    m_occ_to_species.resize(m_Nasym);
    m_species_to_occ.resize(m_Nasym);
    for (Index asym = 0; asym < m_Nasym; ++asym) {
      if (m_asym_to_b[asym].empty()) throw std::runtime_error("Conversions: empty asymmetric unit orbit");
      Index b = *m_asym_to_b[asym].begin();
      m_occ_to_species[asym] = conv[b];
      m_species_to_occ[asym] = conv_inv[b];
    }
  }
  /// round-1 form: n_basis sublattices (occupants "A", "B"), T = diag(n0, n1, n2)
  Conversions(std::vector<long> const &diagonal_T, long n_basis, int device = 0)
      : Conversions(simple_prim(n_basis), diag(diagonal_T), device) {}

  std::array<double, 9> lat_column_mat() const { return m_prim.lat_column_mat; }
  Index l_size() const { return m_l_conv.total_sites(); }
  Index l_to_b(Index l) const { return l_to_bijk(l)[0]; }
  std::vector<long> l_to_ijk(Index l) const {
    auto v = l_to_bijk(l);
    return {v[1], v[2], v[3]};
  }
  std::vector<long> l_to_bijk(Index l) const {
    int64_t o[4];
    m_l_conv.bijk(l, o);
    return {static_cast<long>(o[0]), static_cast<long>(o[1]), static_cast<long>(o[2]), static_cast<long>(o[3])};
  }
  Index l_to_unitl(Index l) const { return bijk_to_unitl(l_to_bijk(l)); }
  Index l_to_asym(Index l) const { return m_unitl_to_asym[l_to_unitl(l)]; }
  std::array<double, 3> l_to_cart(Index l) const {
    auto bijk = l_to_bijk(l);
    std::array<double, 3> c = m_basis_cart[bijk[0]];
    for (int r = 0; r < 3; ++r)
      c[r] += m_prim.lat_column_mat[3 * r] * bijk[1] + m_prim.lat_column_mat[3 * r + 1] * bijk[2] +
              m_prim.lat_column_mat[3 * r + 2] * bijk[3];
    return c;
  }
  std::array<double, 3> l_to_frac(Index l) const {
    auto bijk = l_to_bijk(l);
    std::array<double, 3> f = m_prim.basis_frac[bijk[0]];
    for (int r = 0; r < 3; ++r) f[r] += static_cast<double>(bijk[1 + r]);
    return f;
  }
  std::array<double, 3> l_to_basis_cart(Index l) const { return m_basis_cart[l_to_b(l)]; }
  std::array<double, 3> l_to_basis_frac(Index l) const { return m_prim.basis_frac[l_to_b(l)]; }

  Index bijk_to_l(std::vector<long> const &bijk) const { return convert(m_l_conv, bijk); }
  Index bijk_to_l(long b, long i, long j, long k) const { return bijk_to_l(std::vector<long>{b, i, j, k}); }
  Index bijk_to_unitl(std::vector<long> const &bijk) const { return convert(m_unit_conv, bijk); }
  Index bijk_to_asym(std::vector<long> const &bijk) const { return l_to_asym(bijk_to_l(bijk)); }

  Index unitl_size() const { return m_unit_conv.total_sites(); }
  Index unitl_to_b(Index unitl) const { return unitl_to_bijk(unitl)[0]; }
  std::vector<long> unitl_to_bijk(Index unitl) const {
    int64_t o[4];
    m_unit_conv.bijk(unitl, o);
    return {static_cast<long>(o[0]), static_cast<long>(o[1]), static_cast<long>(o[2]), static_cast<long>(o[3])};
  }
  Index unitl_to_asym(Index unitl) const { return m_unitl_to_asym.at(unitl); }

  Index asym_size() const { return m_Nasym; }
  std::set<Index> const &asym_to_b(Index asym) const { return m_asym_to_b.at(asym); }
  std::set<Index> const &asym_to_unitl(Index asym) const { return m_asym_to_unitl.at(asym); }

  matrix_type const &unit_transformation_matrix_to_super() const { return m_unit_T; }
  matrix_type const &transformation_matrix_to_super() const { return m_T; }
  SiteIndexConverter const &unit_index_converter() const { return m_unit_conv; }
  SiteIndexConverter const &index_converter() const { return m_l_conv; }

  Index occ_size(Index asym) const { return static_cast<Index>(m_occ_to_species.at(asym).size()); }
  Index species_index(Index asym, Index occ_index) const { return m_occ_to_species.at(asym).at(occ_index); }
  /// returns occ_size(asym) if the species is not allowed (Conversions.cc:300-303)
  Index occ_index(Index asym, Index species_index) const { return m_species_to_occ.at(asym).at(species_index); }
  bool species_allowed(Index asym, Index species_index) const {
    return occ_index(asym, species_index) != occ_size(asym);
  }
  Index species_size() const { return static_cast<Index>(m_species.size()); }
  /// index of the name, species_size() if absent (find_index semantics)
  Index species_index(std::string const &species_name) const {
    for (size_t i = 0; i < m_species.size(); ++i)
      if (m_species[i] == species_name) return static_cast<Index>(i);
    return species_size();
  }
  std::vector<std::string> const &species_list() const { return m_species; }
  std::string const &species_name(Index species_index) const { return m_species.at(species_index); }
  /// every species is a single atom in this mirror (no xtal::Molecule)
  Index components_size(Index species_index) const {
    m_species.at(species_index);
    return 1;
  }

  /// batched forms, evaluated on the device
  std::vector<int64_t> l_to_bijk_batch(std::vector<int64_t> const &l) const {
    std::vector<int64_t> out(4 * l.size());
    int64_t t[9];
    for (int i = 0; i < 9; ++i) t[i] = m_T[i];
    cmg_check(cmg_conv_general_l_to_bijk(m_device, t, m_l_conv.n_basis, l.data(),
                                         static_cast<int64_t>(l.size()), out.data()));
    return out;
  }
  std::vector<int64_t> bijk_to_l_batch(std::vector<int64_t> const &bijk) const {
    std::vector<int64_t> out(bijk.size() / 4);
    int64_t t[9];
    for (int i = 0; i < 9; ++i) t[i] = m_T[i];
    cmg_check(cmg_conv_general_bijk_to_l(m_device, t, m_l_conv.n_basis, bijk.data(),
                                         static_cast<int64_t>(out.size()), out.data()));
    return out;
  }

  /// xtal::struc_molecule order: unique occupant names in order of first appearance
  static std::vector<std::string> default_species_list(ConversionsPrim const &prim) {
    std::vector<std::string> out;
    for (auto const &site : prim.occ_dof)
      for (auto const &name : site)
        if (std::find(out.begin(), out.end(), name) == out.end()) out.push_back(name);
    return out;
  }
  /// xtal::make_index_converter: [b][occ] -> species index
  static std::vector<std::vector<Index>> make_index_converter(ConversionsPrim const &prim,
                                                              std::vector<std::string> const &species) {
    std::vector<std::vector<Index>> conv;
    for (auto const &site : prim.occ_dof) {
      std::vector<Index> row;
      for (auto const &name : site) {
        auto it = std::find(species.begin(), species.end(), name);
        if (it == species.end())
          throw std::runtime_error("Conversions: occupant '" + name + "' is not in the species list");
        row.push_back(static_cast<Index>(it - species.begin()));
      }
      conv.push_back(row);
    }
    return conv;
  }
  /// Conversions.cc:14-46 with one symmetry orbit: sublattices are told apart by
  /// their occ -> species lists, orbits numbered in the order of those keys
  static std::vector<Index> default_b_to_asym(ConversionsPrim const &prim,
                                              std::vector<std::string> const &species) {
    std::vector<std::vector<Index>> conv = make_index_converter(prim, species);
    std::map<std::vector<Index>, std::vector<Index>> by_occ;
    for (size_t b = 0; b < conv.size(); ++b) by_occ[conv[b]].push_back(static_cast<Index>(b));
    std::vector<Index> b_to_asym(conv.size());
    Index asym = 0;
    for (auto const &pair : by_occ) {
      for (Index b : pair.second) b_to_asym[b] = asym;
      ++asym;
    }
    return b_to_asym;
  }

 private:
  static ConversionsPrim simple_prim(long n_basis) {
    if (n_basis < 1) throw std::runtime_error("Conversions: need a 3-vector of supercell extents and n_basis >= 1");
    ConversionsPrim p;
    p.occ_dof.assign(static_cast<size_t>(n_basis), {"A", "B"});
    p.basis_frac.assign(static_cast<size_t>(n_basis), {{0.0, 0.0, 0.0}});
    return p;
  }
  static matrix_type diag(std::vector<long> const &d) {
    if (d.size() != 3) throw std::runtime_error("Conversions: need a 3-vector of supercell extents and n_basis >= 1");
    return matrix_type{{d[0], 0, 0, 0, d[1], 0, 0, 0, d[2]}};
  }
  static Index convert(SiteIndexConverter const &f, std::vector<long> const &bijk) {
    if (bijk.size() != 4) throw std::runtime_error("Conversions: bijk needs 4 entries");
    const int64_t in[4] = {bijk[0], bijk[1], bijk[2], bijk[3]};
    return static_cast<Index>(f.linear_site_index(in));
  }
  ConversionsPrim m_prim;
  matrix_type m_T, m_unit_T;
  std::vector<std::string> m_species;
  std::vector<std::array<double, 3>> m_basis_cart;
  SiteIndexConverter m_l_conv, m_unit_conv;
  Index m_Nasym = 0;
  std::vector<Index> m_unitl_to_asym;
  std::vector<std::set<Index>> m_asym_to_unitl, m_asym_to_b;
  std::vector<std::vector<Index>> m_occ_to_species, m_species_to_occ;
  int m_device;
};

}  // namespace casm_monte_b200

#endif
