// monte_module.cpp -- pybind11 face of include/casm_monte_b200/monte.hh.
//
// Mirrors the part of the libcasm.monte Python API that the Ising SGC path
// uses: same class names, constructor arguments, attribute
// and method names as python/src/monte*.cpp of the reference, so the
// reference's own Python tests for this path run with only the import root
// changed.  One extension module; casmcode_monte_b200/monte/** re-exports it in
// the reference's package layout.  numpy arrays are used where the reference
// uses Eigen (pybind11/eigen.h needs Eigen, which is not required here).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/stl_bind.h>

#include "casm_monte_b200/events.hh"
#include "casm_monte_b200/monte.hh"
#include "casm_monte_b200/run_management.hh"

namespace py = pybind11;
using namespace casm_monte_b200;

typedef SemiGrandCanonicalCalculator calculator_type;
typedef SemiGrandCanonicalEventGenerator<default_engine_type> event_generator_type;
typedef std::map<std::string, std::vector<double>> VectorValueMap;
typedef std::map<std::string, double> ScalarValueMap;
typedef std::map<std::string, bool> BooleanValueMap;

// JSON-valued sampling (include/casm/monte/sampling/StateSamplingFunction.hh:79-95,
// Sampler.hh:114-117): JSON values are Python objects here.
struct PyJsonStateSamplingFunction {
  std::string name, description;
  py::object function;
};
struct PyJsonSampler {
  py::list values;
};
typedef std::map<std::string, PyJsonStateSamplingFunction> PyJsonStateSamplingFunctionMap;
typedef std::map<std::string, PyJsonSampler> PyJsonSamplerMap;

// per-run extras that live next to SemiGrandCanonicalData in the binding
struct RunExtras {
  PyJsonStateSamplingFunctionMap json_sampling_functions;
  PyJsonSamplerMap json_samplers;
};
// Keyed by the data object's address.  Deliberately leaked: it holds Python
// objects, which must not be released after the interpreter has shut down.
typedef std::map<SemiGrandCanonicalData const *, std::shared_ptr<RunExtras>> ExtrasRegistry;
static ExtrasRegistry &extras_registry() {
  static ExtrasRegistry *r = new ExtrasRegistry();
  return *r;
}
static void register_extras(SemiGrandCanonicalData const *d, std::shared_ptr<RunExtras> ex) {
  ExtrasRegistry &r = extras_registry();
  if (r.size() > 4096) r.clear();  // bound the bookkeeping of long sessions
  r[d] = std::move(ex);
}

PYBIND11_MAKE_OPAQUE(SamplerMap);
PYBIND11_MAKE_OPAQUE(StateSamplingFunctionMap);
PYBIND11_MAKE_OPAQUE(PyJsonStateSamplingFunctionMap);
PYBIND11_MAKE_OPAQUE(PyJsonSamplerMap);
PYBIND11_MAKE_OPAQUE(RequestedPrecisionMap);
PYBIND11_MAKE_OPAQUE(ResultsAnalysisFunctionMap);
PYBIND11_MAKE_OPAQUE(ScalarValueMap);
PYBIND11_MAKE_OPAQUE(VectorValueMap);
PYBIND11_MAKE_OPAQUE(BooleanValueMap);
PYBIND11_MAKE_OPAQUE(std::vector<long>);
PYBIND11_MAKE_OPAQUE(std::vector<int>);

namespace {

py::array_t<double> as_array(std::vector<double> const &v) {
  py::array_t<double> a(v.size());
  std::copy(v.begin(), v.end(), a.mutable_data());
  return a;
}
std::vector<double> to_dvec(py::handle h) {
  py::array_t<double, py::array::c_style | py::array::forcecast> a =
      py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(h);
  if (!a) throw std::runtime_error("expected a float array");
  std::vector<double> v(a.size());
  std::copy(a.data(), a.data() + a.size(), v.begin());
  return v;
}
std::vector<int> to_ivec(py::handle h) {
  py::array_t<long, py::array::c_style | py::array::forcecast> a =
      py::array_t<long, py::array::c_style | py::array::forcecast>::ensure(h);
  if (!a) throw std::runtime_error("expected an integer array");
  std::vector<int> v(a.size());
  for (py::ssize_t i = 0; i < a.size(); ++i) v[i] = static_cast<int>(a.data()[i]);
  return v;
}

// ---- ValueMap <-> dict (src/casm/monte/io/json/ValueMap_json_io.cc:9-55) ----
ValueMap valuemap_from_dict(py::dict d) {
  ValueMap v;
  for (auto item : d) {
    std::string key = py::str(item.first);
    py::handle val = item.second;
    if (py::isinstance<py::bool_>(val)) {
      v.boolean_values[key] = val.cast<bool>();
    } else if (py::isinstance<py::int_>(val) || py::isinstance<py::float_>(val)) {
      v.scalar_values[key] = val.cast<double>();
    } else {
      py::array_t<double, py::array::c_style | py::array::forcecast> a =
          py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(val);
      if (!a) throw std::runtime_error("Error in ValueMap.from_dict: unsupported value for '" + key + "'");
      if (a.ndim() == 1) {
        v.vector_values[key] = std::vector<double>(a.data(), a.data() + a.size());
      } else if (a.ndim() == 2) {
        MatrixValue m;
        m.rows = a.shape(0);
        m.cols = a.shape(1);
        m.data.resize(a.size());
        for (Index c = 0; c < m.cols; ++c)
          for (Index r = 0; r < m.rows; ++r) m.data[r + m.rows * c] = a.at(r, c);
        v.matrix_values[key] = m;
      } else {
        throw std::runtime_error("Error in ValueMap.from_dict: unsupported array rank");
      }
    }
  }
  return v;
}
py::dict valuemap_to_dict(ValueMap const &v) {
  py::dict d;
  for (auto const &p : v.boolean_values) d[py::str(p.first)] = p.second;
  for (auto const &p : v.scalar_values) d[py::str(p.first)] = p.second;
  for (auto const &p : v.vector_values) d[py::str(p.first)] = py::cast(p.second);
  for (auto const &p : v.matrix_values) {
    py::list rows;
    for (Index r = 0; r < p.second.rows; ++r) {
      py::list row;
      for (Index c = 0; c < p.second.cols; ++c) row.append(p.second.data[r + p.second.rows * c]);
      rows.append(row);
    }
    d[py::str(p.first)] = rows;
  }
  return d;
}

py::dict req_prec_to_dict(RequestedPrecision const &r) {
  py::dict d;
  if (r.abs_convergence_is_required) d["abs_precision"] = r.abs_precision;
  if (r.rel_convergence_is_required) d["rel_precision"] = r.rel_precision;
  return d;
}
py::dict stats_to_dict(BasicStatistics const &s) {
  py::dict d;
  d["mean"] = s.mean;
  d["calculated_precision"] = s.calculated_precision;
  return d;
}
py::dict eq_results_to_dict(EquilibrationCheckResults const &v) {
  py::dict d;
  d["all_equilibrated"] = v.all_equilibrated;
  if (v.all_equilibrated)
    d["N_samples_for_all_to_equilibrate"] = v.N_samples_for_all_to_equilibrate;
  else
    d["N_samples_for_equilibration"] = "did_not_equilibrate";
  py::list l;
  for (auto const &p : v.individual_results) {
    py::dict t;
    t["is_equilibrated"] = p.second.is_equilibrated;
    if (p.second.is_equilibrated)
      t["N_samples_for_equilibration"] = p.second.N_samples_for_equilibration;
    else
      t["N_samples_for_equilibration"] = "did_not_equilibrate";
    t["sampler_name"] = p.first.sampler_name;
    t["component_name"] = p.first.component_name;
    t["component_index"] = p.first.component_index;
    l.append(t);
  }
  d["individual_results"] = l;
  return d;
}
py::dict conv_results_to_dict(ConvergenceCheckResults const &v) {
  py::dict d;
  d["all_converged"] = v.all_converged;
  d["N_samples_for_statistics"] = v.N_samples_for_statistics;
  py::list l;
  for (auto const &p : v.individual_results) {
    py::dict t;
    t["is_converged"] = p.second.is_converged;
    t["requested_precision"] = req_prec_to_dict(p.second.requested_precision);
    t["stats"] = stats_to_dict(p.second.stats);
    t["sampler_name"] = p.first.sampler_name;
    t["component_name"] = p.first.component_name;
    t["component_index"] = p.first.component_index;
    l.append(t);
  }
  d["individual_results"] = l;
  return d;
}
// include/casm/monte/checks/io/json/CompletionCheck_json_io.hh:421-438
py::dict cc_results_to_dict(CompletionCheckResults const &v) {
  py::dict d;
  d["has_all_minimums_met"] = v.has_all_minimums_met;
  d["has_any_maximum_met"] = v.has_any_maximum_met;
  d["count"] = v.count.has_value() ? py::cast(*v.count) : py::none();
  d["time"] = v.time.has_value() ? py::cast(*v.time) : py::none();
  d["clocktime"] = v.clocktime;
  d["n_samples"] = v.n_samples;
  d["is_complete"] = v.is_complete;
  if (v.n_samples_at_convergence_check.has_value()) {
    d["n_samples_at_convergence_check"] = *v.n_samples_at_convergence_check;
    d["equilibration_check_results"] = eq_results_to_dict(v.equilibration_check_results);
    d["convergence_check_results"] = conv_results_to_dict(v.convergence_check_results);
  }
  return d;
}

template <typename T>
py::list to_pylist(std::vector<T> const &v) {
  py::list l;
  for (auto const &x : v) l.append(x);
  return l;
}
template <typename T>
py::array_t<long> to_long_array(std::vector<T> const &v) {
  py::array_t<long> a(v.size());
  for (size_t i = 0; i < v.size(); ++i) a.mutable_data()[i] = static_cast<long>(v[i]);
  return a;
}
// CutoffCheckParams <-> dict (src/casm/monte/checks/io/json/CutoffCheck_json_io.cc:11-92)
py::dict cutoff_to_dict(CutoffCheckParams const &p) {
  py::dict d;
  auto put = [&](const char *key, auto const &mn, auto const &mx) {
    if (!mn.has_value() && !mx.has_value()) return;
    py::dict t;
    if (mn.has_value()) t["min"] = *mn;
    if (mx.has_value()) t["max"] = *mx;
    d[key] = t;
  };
  put("count", p.min_count, p.max_count);
  put("time", p.min_time, p.max_time);
  put("sample", p.min_sample, p.max_sample);
  put("clocktime", p.min_clocktime, p.max_clocktime);
  return d;
}
CutoffCheckParams cutoff_from_dict(py::dict d) {
  CutoffCheckParams p;
  auto get = [&](const char *key, auto &mn, auto &mx) {
    if (!d.contains(key) || d[key].is_none()) return;
    py::dict t = d[key].cast<py::dict>();
    typedef typename std::remove_reference<decltype(mn)>::type::value_type T;
    if (t.contains("min") && !t["min"].is_none()) mn = t["min"].cast<T>();
    if (t.contains("max") && !t["max"].is_none()) mx = t["max"].cast<T>();
  };
  get("count", p.min_count, p.max_count);
  get("time", p.min_time, p.max_time);
  get("sample", p.min_sample, p.max_sample);
  get("clocktime", p.min_clocktime, p.max_clocktime);
  return p;
}
RequestedPrecision req_prec_from_dict(py::dict d) {
  // src/casm/monte/sampling/io/json/Sampler_json_io.cc:33-78 ("precision" is the deprecated key)
  RequestedPrecision r;
  for (const char *k : {"abs_precision", "precision"})
    if (d.contains(k)) {
      r.abs_convergence_is_required = true;
      r.abs_precision = d[k].cast<double>();
    }
  if (d.contains("rel_precision")) {
    r.rel_convergence_is_required = true;
    r.rel_precision = d["rel_precision"].cast<double>();
  }
  return r;
}
// include/casm/monte/checks/io/json/CompletionCheck_json_io.hh:36-344 (parse): same keys,
// defaults and error messages; all errors are collected and reported together
CompletionCheckParams completion_check_params_from_dict(py::dict data,
                                                        StateSamplingFunctionMap const &sampling_functions) {
  std::vector<std::string> errors;
  CompletionCheckParams p;
  double confidence = 0.95;
  Index method = 1, n_resamples = 10000;
  if (data.contains("confidence")) confidence = data["confidence"].cast<double>();
  if (data.contains("weighted_observations_method")) method = data["weighted_observations_method"].cast<Index>();
  if (data.contains("n_resamples")) n_resamples = data["n_resamples"].cast<Index>();
  p.equilibration_check_f = default_equilibration_check;
  p.calc_statistics_f = BasicStatisticsCalculator(confidence, method, n_resamples);
  if (data.contains("cutoff") && !data["cutoff"].is_none()) p.cutoff_params = cutoff_from_dict(data["cutoff"].cast<py::dict>());
  if (data.contains("convergence")) {
    if (!py::isinstance<py::list>(data["convergence"]) && !py::isinstance<py::tuple>(data["convergence"])) {
      errors.push_back("Error: \"convergence\" must be an array");
    } else {
      for (py::handle h : data["convergence"]) {
        py::dict c = h.cast<py::dict>();
        if (!c.contains("quantity")) {
          errors.push_back("Error: missing required option \"quantity\"");
          continue;
        }
        std::string quantity = c["quantity"].cast<std::string>();
        auto fit = sampling_functions.find(quantity);
        if (fit == sampling_functions.end()) {
          errors.push_back("Error: \"" + quantity + "\" is not a sampling option.");
          continue;
        }
        StateSamplingFunction const &f = fit->second;
        RequestedPrecision precision = req_prec_from_dict(c);
        const bool has_index = c.contains("component_index"), has_name = c.contains("component_name");
        if (has_index)
          for (Index index : c["component_index"].cast<std::vector<Index>>()) {
            const Index size = static_cast<Index>(f.component_names.size());
            if (index < 0 || index >= size) {
              errors.push_back("Error: For \"" + f.name + "\", component index " + std::to_string(index) +
                               " is out of range. Valid range is [0," + std::to_string(size) + ").");
              continue;
            }
            p.requested_precision.emplace(SamplerComponent(f.name, index, f.component_names[index]), precision);
          }
        if (has_name)
          for (std::string const &name : c["component_name"].cast<std::vector<std::string>>()) {
            auto it = std::find(f.component_names.begin(), f.component_names.end(), name);
            if (it == f.component_names.end()) {
              errors.push_back("Error: For \"" + f.name + "\", component name " + name + " is not valid.");
              continue;
            }
            p.requested_precision.emplace(
                SamplerComponent(f.name, static_cast<Index>(it - f.component_names.begin()), name), precision);
          }
        if (!has_index && !has_name)
          for (Index index = 0; index < static_cast<Index>(f.component_names.size()); ++index)
            p.requested_precision.emplace(SamplerComponent(f.name, index, f.component_names[index]), precision);
      }
    }
  }
  std::string spacing = "linear";
  if (data.contains("spacing")) spacing = data["spacing"].cast<std::string>();
  if (spacing == "linear") {
    p.log_spacing = false;
    p.check_begin = 100;
    p.check_period = 100;
  } else if (spacing == "log") {
    p.log_spacing = true;
    p.check_begin = 0;
    p.check_base = 10.0;
    p.check_shift = 2.0;
    p.check_period_max = 10000;
  } else {
    errors.push_back("Error: \"spacing\" must be one of \"linear\", \"log\".");
  }
  if (data.contains("begin")) p.check_begin = data["begin"].cast<CountType>();
  if (data.contains("period")) p.check_period = data["period"].cast<CountType>();
  if (p.check_period <= 1) errors.push_back("Error: \"period\" must > 0.");
  if (data.contains("base")) p.check_base = data["base"].cast<double>();
  if (data.contains("shift")) p.check_shift = data["shift"].cast<double>();
  if (data.contains("period_max")) p.check_period_max = data["period_max"].cast<CountType>();
  if (p.check_base <= 1.0) errors.push_back("Error: \"base\" must > 1.0");
  if (!errors.empty()) {
    std::string msg = "Error in libcasm.monte.sampling.CompletionCheckParams.from_dict";
    for (auto const &e : errors) msg += "\n  " + e;
    throw std::runtime_error(msg);
  }
  return p;
}
// CompletionCheck_json_io.hh:347-395 (to_json)
py::dict completion_check_params_to_dict(CompletionCheckParams const &p) {
  py::dict d;
  d["cutoff"] = cutoff_to_dict(p.cutoff_params);
  py::list conv;
  for (auto const &pair : p.requested_precision) {
    py::dict t;
    t["quantity"] = pair.first.sampler_name;
    t["component_index"] = std::vector<Index>{pair.first.component_index};
    t["component_name"] = std::vector<std::string>{pair.first.component_name};
    for (auto item : req_prec_to_dict(pair.second)) t[item.first] = item.second;
    conv.append(t);
  }
  d["convergence"] = conv;
  if (!p.log_spacing) {
    d["spacing"] = "linear";
    d["begin"] = p.check_begin;
    d["period"] = p.check_period;
  } else {
    d["spacing"] = "log";
    d["begin"] = p.check_begin;
    d["base"] = p.check_base;
    d["shift"] = p.check_shift;
    d["period_max"] = p.check_period_max;
  }
  return d;
}

py::dict config_to_dict(IsingConfiguration const &c) {
  py::dict d;
  d["shape"] = to_pylist(c.shape);
  d["occupation"] = to_pylist(c.occupation());
  return d;
}
IsingConfiguration config_from_dict(py::dict d) {
  if (!d.contains("shape"))
    throw std::runtime_error("Error reading IsingConfiguration from JSON: no 'shape'");
  if (!d.contains("occupation"))
    throw std::runtime_error("Error reading IsingConfiguration from JSON: no 'occupation'");
  IsingConfiguration c(to_ivec(d["shape"]));
  c.set_occupation(to_ivec(d["occupation"]));
  return c;
}

struct PyEngine {
  std::shared_ptr<default_engine_type> e;
};

}  // namespace

PYBIND11_MODULE(_monte_b200, m) {
  m.doc() = "B200-native Ising SGC Metropolis path behind the libcasm.monte API subset";
  m.attr("KB") = KB;

  // ------------------------------------------------------------------ monte
  py::bind_map<ScalarValueMap>(m, "ScalarValueMap");
  py::bind_map<BooleanValueMap>(m, "BooleanValueMap");
  // vector values: numpy views so that values.vector_values["x"][0] = 2.0 sticks
  py::class_<VectorValueMap>(m, "VectorValueMap")
      .def(py::init<>())
      .def("__len__", [](VectorValueMap const &v) { return v.size(); })
      .def("__contains__", [](VectorValueMap const &v, std::string const &k) { return v.count(k) > 0; })
      .def("__iter__", [](VectorValueMap const &v) { return py::make_key_iterator(v.begin(), v.end()); },
           py::keep_alive<0, 1>())
      .def("keys", [](VectorValueMap const &v) {
        py::list l;
        for (auto const &p : v) l.append(p.first);
        return l;
      })
      .def("items", [](py::object self) {
        auto &v = self.cast<VectorValueMap &>();
        py::list l;
        for (auto &p : v)
          l.append(py::make_tuple(p.first, py::array_t<double>({p.second.size()}, {sizeof(double)},
                                                               p.second.data(), self)));
        return l;
      })
      .def("__getitem__", [](py::object self, std::string const &k) {
        auto &v = self.cast<VectorValueMap &>();
        auto it = v.find(k);
        if (it == v.end()) throw py::key_error(k);
        return py::array_t<double>({it->second.size()}, {sizeof(double)}, it->second.data(), self);
      })
      .def("__setitem__", [](VectorValueMap &v, std::string const &k, py::object val) { v[k] = to_dvec(val); })
      .def("__delitem__", [](VectorValueMap &v, std::string const &k) {
        if (!v.erase(k)) throw py::key_error(k);
      });

  py::class_<ValueMap>(m, "ValueMap")
      .def(py::init([](std::optional<py::dict> data) {
             return data.has_value() ? valuemap_from_dict(*data) : ValueMap();
           }),
           py::arg("data") = py::none())
      .def_readwrite("boolean_values", &ValueMap::boolean_values)
      .def_readwrite("scalar_values", &ValueMap::scalar_values)
      .def_readwrite("vector_values", &ValueMap::vector_values)
      .def("is_mismatched", [](ValueMap const &a, ValueMap const &b) { return is_mismatched(a, b); })
      .def("make_incremented_values",
           [](ValueMap const &a, ValueMap const &inc, double n) { return make_incremented_values(a, inc, n); },
           py::arg("increment"), py::arg("n_increment"))
      .def_static("from_dict", &valuemap_from_dict, py::arg("data"))
      .def("to_dict", &valuemap_to_dict)
      .def("__copy__", [](ValueMap const &v) { return ValueMap(v); })
      .def("__deepcopy__", [](ValueMap const &v, py::dict) { return ValueMap(v); });

  py::class_<MethodLog>(m, "MethodLog")
      .def(py::init([](std::optional<std::string> logfile_path, std::optional<double> log_frequency) {
             MethodLog l;
             if (logfile_path.has_value()) {
               l.logfile_path = *logfile_path;
               l.reset();
             }
             l.log_frequency = log_frequency;
             return l;
           }),
           py::arg("logfile_path") = py::none(), py::arg("log_frequency") = py::none())
      .def("logfile_path", [](MethodLog const &l) { return l.logfile_path; })
      .def("log_frequency", [](MethodLog const &l) { return l.log_frequency; })
      .def("reset", &MethodLog::reset)
      .def("reset_to_stdout", &MethodLog::reset_to_stdout)
      .def("restart_clock", [](MethodLog &l) { l.log.restart_clock(); })
      .def("time_s", [](MethodLog const &l) { return l.log.time_s(); })
      .def("begin_lap", [](MethodLog &l) { l.log.begin_lap(); })
      .def("lap_time", [](MethodLog const &l) { return l.log.lap_time(); })
      .def("print", [](MethodLog &l, std::string const &what) { (*l.log.out) << what; l.log.out->flush(); })
      .def("section", [](MethodLog &l, std::string const &what, bool show_clock) {
        (*l.log.out) << "-- " << what << " -- ";
        if (show_clock) (*l.log.out) << "Time: " << l.log.time_s() << " (s)";
        (*l.log.out) << std::endl;
      }, py::arg("what"), py::arg("show_clock") = false);

  // python/src/monte.cpp:469-550
  py::class_<PyEngine>(m, "RandomNumberEngine")
      .def(py::init([]() {
        PyEngine e;
        e.e = std::make_shared<default_engine_type>();
        std::random_device device;
        e.e->seed(device());
        return e;
      }))
      .def("seed", [](PyEngine &e, uint64_t value) { e.e->seed(value); }, py::arg("value"))
      .def("seed_seq", [](PyEngine &e, std::vector<uint32_t> values) {
        std::seed_seq ss(values.begin(), values.end());
        e.e->seed(ss);
      }, py::arg("values"))
      .def("dump", [](PyEngine const &e) {
        std::stringstream ss;
        ss << *e.e;
        return ss.str();
      })
      .def("load", [](PyEngine &e, std::string state) {
        std::stringstream ss(state);
        ss >> *e.e;
      }, py::arg("state"));

  py::class_<RandomNumberGenerator<>>(m, "RandomNumberGenerator")
      .def(py::init([](std::optional<PyEngine> engine) {
             return RandomNumberGenerator<>(engine.has_value() ? engine->e : nullptr);
           }),
           py::arg("engine") = py::none())
      .def("random_int", [](RandomNumberGenerator<> &g, uint64_t maximum_value) {
        return g.random_int<uint64_t>(maximum_value);
      }, py::arg("maximum_value"))
      .def("random_real", [](RandomNumberGenerator<> &g, double maximum_value) {
        return g.random_real<double>(maximum_value);
      }, py::arg("maximum_value"))
      .def("engine", [](RandomNumberGenerator<> &g) {
        PyEngine e;
        e.e = g.engine;
        return e;
      });

  // ----------------------------------------------------------------- events
  py::bind_vector<std::vector<long>>(m, "LongVector");
  py::bind_vector<std::vector<int>>(m, "IntVector");
  // python lists / tuples / arrays are accepted wherever these vectors are expected
  py::implicitly_convertible<py::list, std::vector<long>>();
  py::implicitly_convertible<py::tuple, std::vector<long>>();
  py::implicitly_convertible<py::array, std::vector<long>>();
  py::implicitly_convertible<py::list, std::vector<int>>();
  py::implicitly_convertible<py::tuple, std::vector<int>>();
  py::implicitly_convertible<py::array, std::vector<int>>();
  py::class_<OccTransform>(m, "OccTransform")
      .def(py::init<>())
      .def_readwrite("linear_site_index", &OccTransform::l)
      .def_readwrite("mol_id", &OccTransform::mol_id)
      .def_readwrite("asym", &OccTransform::asym)
      .def_readwrite("from_species", &OccTransform::from_species)
      .def_readwrite("to_species", &OccTransform::to_species);
  py::class_<OccEvent>(m, "OccEvent")
      .def(py::init<>())
      .def_readwrite("linear_site_index", &OccEvent::linear_site_index)
      .def_readwrite("new_occ", &OccEvent::new_occ)
      .def_readwrite("occ_transform", &OccEvent::occ_transform);

  // ---- general multi-species proposal machinery (include/casm_monte_b200/events.hh;
  // python/src/monte_events.cpp:432-1100 of the reference) ----
  py::class_<OccCandidate>(m, "OccCandidate")
      .def(py::init<Index, Index>(), py::arg("asymmetric_unit_index"), py::arg("species_index"))
      .def_readwrite("asymmetric_unit_index", &OccCandidate::asym)
      .def_readwrite("species_index", &OccCandidate::species_index)
      .def("is_valid", [](OccCandidate const &c, Conversions const &convert) { return is_valid(convert, c); })
      .def("__lt__", [](OccCandidate const &a, OccCandidate const &b) { return a < b; })
      .def("__eq__", [](OccCandidate const &a, OccCandidate const &b) { return a == b; })
      .def("to_tuple", [](OccCandidate const &c) { return py::make_tuple(c.asym, c.species_index); });
  py::class_<OccSwap>(m, "OccSwap")
      .def(py::init<OccCandidate const &, OccCandidate const &>(), py::arg("first"), py::arg("second"))
      .def_readwrite("first", &OccSwap::cand_a)
      .def_readwrite("second", &OccSwap::cand_b)
      .def("reverse", &OccSwap::reverse)
      .def("sort", [](OccSwap &s) { s.sort(); })
      .def("sorted", &OccSwap::sorted)
      .def("is_valid", [](OccSwap const &s, Conversions const &convert) { return is_valid(convert, s); })
      .def("__lt__", [](OccSwap const &a, OccSwap const &b) { return a < b; })
      .def("__eq__", [](OccSwap const &a, OccSwap const &b) { return a == b; })
      .def("to_tuple", [](OccSwap const &s) {
        return py::make_tuple(s.cand_a.asym, s.cand_a.species_index, s.cand_b.asym, s.cand_b.species_index);
      });
  py::class_<OccCandidateList>(m, "OccCandidateList")
      .def(py::init<Conversions const &>(), py::arg("convert"))
      .def(py::init<std::vector<OccCandidate>, Conversions const &>(), py::arg("candidates"), py::arg("convert"))
      .def("index", [](OccCandidateList const &l, OccCandidate const &c) { return l.index(c); })
      .def("matching_index", [](OccCandidateList const &l, Index asym, Index species_index) { return l.index(asym, species_index); })
      .def("__getitem__", [](OccCandidateList const &l, Index i) { return l[i]; })
      .def("__len__", &OccCandidateList::size)
      .def("__iter__", [](OccCandidateList const &l) { return py::make_iterator(l.begin(), l.end()); },
           py::keep_alive<0, 1>());
  m.def("is_allowed_canonical_swap", [](Conversions const &c, OccCandidate a, OccCandidate b) { return allowed_canonical_swap(c, a, b); });
  m.def("make_canonical_swaps", &make_canonical_swaps, py::arg("convert"), py::arg("occ_candidate_list"));
  m.def("is_allowed_semigrand_canonical_swap", [](Conversions const &c, OccCandidate a, OccCandidate b) { return allowed_semigrand_canonical_swap(c, a, b); });
  m.def("make_semigrand_canonical_swaps", &make_semigrand_canonical_swaps, py::arg("convert"), py::arg("occ_candidate_list"));
  m.def("get_n_allowed_per_unitcell", &get_n_allowed_per_unitcell, py::arg("convert"), py::arg("semigrand_canonical_swaps"));
  py::class_<Mol>(m, "Mol")
      .def(py::init<>())
      .def_readwrite("id", &Mol::id)
      .def_readwrite("linear_site_index", &Mol::l)
      .def_readwrite("asymmetric_unit_index", &Mol::asym)
      .def_readwrite("species_index", &Mol::species_index)
      .def_readwrite("mol_location_index", &Mol::loc);
  py::class_<OccLocation>(m, "OccLocation")
      .def(py::init<Conversions const &, OccCandidateList const &, bool, bool, bool>(), py::arg("convert"),
           py::arg("candidate_list"), py::arg("update_atoms") = false, py::arg("track_unique_atoms") = false,
           py::arg("save_atom_info") = false, py::keep_alive<1, 2>(), py::keep_alive<1, 3>())
      .def("initialize", [](OccLocation &o, py::object occupation) { o.initialize(to_ivec(occupation)); }, py::arg("occupation"))
      .def("apply",
           [](OccLocation &o, OccEvent const &e, py::array_t<int32_t, py::array::c_style> occupation) {
             // the caller's array is updated in place, as the reference's Eigen::Ref argument is
             std::vector<int> occ(occupation.data(), occupation.data() + occupation.size());
             o.apply(e, occ);
             std::copy(occ.begin(), occ.end(), occupation.mutable_data());
           },
           py::arg("e"), py::arg("occupation"))
      .def("choose_mol", [](OccLocation const &o, OccCandidate const &c, RandomNumberGenerator<> &rng) { return o.choose_mol(c, rng); },
           py::arg("cand"), py::arg("random_number_generator"))
      .def("choose_mol_by_candidate_index", [](OccLocation const &o, Index i, RandomNumberGenerator<> &rng) { return o.choose_mol(i, rng); },
           py::arg("cand_index"), py::arg("random_number_generator"))
      .def("mol_size", &OccLocation::mol_size)
      .def("mol", [](OccLocation const &o, Index id) { return o.mol(id); })
      .def("cand_size", [](OccLocation const &o, OccCandidate const &c) { return o.cand_size(c); })
      .def("cand_size_by_candidate_index", [](OccLocation const &o, Index i) { return o.cand_size(i); })
      .def("mol_id", [](OccLocation const &o, OccCandidate const &c, Index loc) { return o.mol_id(c, loc); })
      .def("mol_id_by_candidate_index", [](OccLocation const &o, Index i, Index loc) { return o.mol_id(i, loc); })
      .def("linear_site_index_to_mol_id", &OccLocation::l_to_mol_id);
  m.def("choose_canonical_swap",
        [](OccLocation const &o, std::vector<OccSwap> const &swaps, RandomNumberGenerator<> &rng) { return choose_canonical_swap(o, swaps, rng); },
        py::arg("occ_location"), py::arg("canonical_swaps"), py::arg("random_number_generator"));
  m.def("propose_canonical_event_from_swap",
        [](OccEvent &e, OccLocation const &o, OccSwap const &swap, RandomNumberGenerator<> &rng) { propose_canonical_event_from_swap(e, o, swap, rng); },
        py::arg("occ_event"), py::arg("occ_location"), py::arg("swap"), py::arg("random_number_generator"));
  m.def("propose_canonical_event",
        [](OccEvent &e, OccLocation const &o, std::vector<OccSwap> const &swaps, RandomNumberGenerator<> &rng) { propose_canonical_event(e, o, swaps, rng); },
        py::arg("occ_event"), py::arg("occ_location"), py::arg("canonical_swaps"), py::arg("random_number_generator"));
  m.def("choose_semigrand_canonical_swap",
        [](OccLocation const &o, std::vector<OccSwap> const &swaps, RandomNumberGenerator<> &rng) { return choose_semigrand_canonical_swap(o, swaps, rng); },
        py::arg("occ_location"), py::arg("semigrand_canonical_swaps"), py::arg("random_number_generator"));
  m.def("propose_semigrand_canonical_event_from_swap",
        [](OccEvent &e, OccLocation const &o, OccSwap const &swap, RandomNumberGenerator<> &rng) { propose_semigrand_canonical_event_from_swap(e, o, swap, rng); },
        py::arg("occ_event"), py::arg("occ_location"), py::arg("swap"), py::arg("random_number_generator"));
  m.def("propose_semigrand_canonical_event",
        [](OccEvent &e, OccLocation const &o, std::vector<OccSwap> const &swaps, RandomNumberGenerator<> &rng) { propose_semigrand_canonical_event(e, o, swaps, rng); },
        py::arg("occ_event"), py::arg("occ_location"), py::arg("semigrand_canonical_swaps"), py::arg("random_number_generator"));

  // events.Conversions (python/src/monte_events.cpp:85-430).  libcasm.xtal is absent, so
  // the prim is given as arrays: occ_dof (names per sublattice), the 3 x 3
  // transformation matrix, and optionally the lattice column-vector matrix and the
  // fractional basis coordinates as COLUMNS (shape 3 x n_basis, like xtal.Prim).
  auto to_matrix = [](py::object o) {
    auto a = py::array_t<int64_t, py::array::c_style | py::array::forcecast>::ensure(o);
    if (!a || a.ndim() != 2 || a.shape(0) != 3 || a.shape(1) != 3)
      throw std::runtime_error("Conversions: transformation matrix must be 3 x 3 integers");
    Conversions::matrix_type t;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) t[3 * r + c] = static_cast<long>(a.at(r, c));
    return t;
  };
  auto to_prim = [](std::vector<std::vector<std::string>> occ_dof, py::object lattice, py::object frac) {
    ConversionsPrim p;
    p.occ_dof = std::move(occ_dof);
    if (!lattice.is_none()) {
      auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(lattice);
      if (!a || a.ndim() != 2 || a.shape(0) != 3 || a.shape(1) != 3)
        throw std::runtime_error("Conversions: lattice column vector matrix must be 3 x 3");
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) p.lat_column_mat[3 * r + c] = a.at(r, c);
    }
    if (!frac.is_none()) {
      auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(frac);
      if (!a || a.ndim() != 2 || a.shape(0) != 3 || static_cast<size_t>(a.shape(1)) != p.occ_dof.size())
        throw std::runtime_error("Conversions: coordinate_frac must have shape (3, n_basis)");
      for (py::ssize_t b = 0; b < a.shape(1); ++b) p.basis_frac.push_back({{a.at(0, b), a.at(1, b), a.at(2, b)}});
    }
    return p;
  };
  auto bijk_arg = [](std::vector<long> const &bijk) {
    if (bijk.size() != 4) throw std::runtime_error("bijk must have 4 entries");
    return bijk;
  };
  py::class_<Conversions>(m, "Conversions")
      .def(py::init<std::vector<long>, long, int>(), py::arg("supercell_extents"),
           py::arg("n_basis") = 1, py::arg("device") = 0)
      .def(py::init([to_matrix, to_prim](std::vector<std::vector<std::string>> occ_dof, py::object T,
                                         py::object lattice, py::object frac, int device) {
             return Conversions(to_prim(std::move(occ_dof), lattice, frac), to_matrix(T), device);
           }),
           py::arg("occ_dof"), py::arg("transformation_matrix_to_super"),
           py::arg("lattice_column_vector_matrix") = py::none(), py::arg("coordinate_frac") = py::none(),
           py::arg("device") = 0)
      .def_static(
          "make_with_custom_asym",
          [to_matrix, to_prim](std::vector<std::vector<std::string>> occ_dof, py::object T, std::vector<Index> b_to_asym,
                               py::object lattice, py::object frac, int device) {
            return Conversions(to_prim(std::move(occ_dof), lattice, frac), to_matrix(T), b_to_asym, device);
          },
          py::arg("occ_dof"), py::arg("transformation_matrix_to_super"), py::arg("b_to_asym"),
          py::arg("lattice_column_vector_matrix") = py::none(), py::arg("coordinate_frac") = py::none(),
          py::arg("device") = 0)
      .def_static(
          "make_with_custom_unitcell",
          [to_matrix, to_prim](std::vector<std::vector<std::string>> occ_dof, std::vector<std::string> species_list,
                               py::object T, py::object unit_T, std::vector<Index> unitl_to_asym, py::object lattice,
                               py::object frac, int device) {
            return Conversions(to_prim(std::move(occ_dof), lattice, frac), species_list, to_matrix(T),
                               to_matrix(unit_T), unitl_to_asym, device);
          },
          py::arg("occ_dof"), py::arg("species_list"), py::arg("transformation_matrix_to_super"),
          py::arg("unit_transformation_matrix_to_super"), py::arg("unitl_to_asym"),
          py::arg("lattice_column_vector_matrix") = py::none(), py::arg("coordinate_frac") = py::none(),
          py::arg("device") = 0)
      .def("lat_column_mat",
           [](Conversions const &c) {
             py::array_t<double> a({3, 3});
             auto m = c.lat_column_mat();
             std::copy(m.begin(), m.end(), a.mutable_data());
             return a;
           })
      .def("l_size", &Conversions::l_size)
      .def("l_to_b", &Conversions::l_to_b)
      .def("l_to_ijk", &Conversions::l_to_ijk)
      .def("l_to_bijk", [](Conversions const &c, long l) { return c.l_to_bijk(l); })
      .def("l_to_unitl", &Conversions::l_to_unitl)
      .def("l_to_asym", &Conversions::l_to_asym)
      .def("l_to_cart", &Conversions::l_to_cart)
      .def("l_to_frac", &Conversions::l_to_frac)
      .def("l_to_basis_cart", &Conversions::l_to_basis_cart)
      .def("l_to_basis_frac", &Conversions::l_to_basis_frac)
      .def("bijk_to_l", [bijk_arg](Conversions const &c, std::vector<long> bijk) { return c.bijk_to_l(bijk_arg(bijk)); })
      .def("bijk_to_unitl", [bijk_arg](Conversions const &c, std::vector<long> bijk) { return c.bijk_to_unitl(bijk_arg(bijk)); })
      .def("bijk_to_asym", [bijk_arg](Conversions const &c, std::vector<long> bijk) { return c.bijk_to_asym(bijk_arg(bijk)); })
      .def("unitl_size", &Conversions::unitl_size)
      .def("unitl_to_b", &Conversions::unitl_to_b)
      .def("unitl_to_bijk", &Conversions::unitl_to_bijk)
      .def("unitl_to_asym", &Conversions::unitl_to_asym)
      .def("asym_size", &Conversions::asym_size)
      .def("asym_to_b", [](Conversions const &c, Index asym) { return c.asym_to_b(asym); })
      .def("asym_to_unitl", [](Conversions const &c, Index asym) { return c.asym_to_unitl(asym); })
      .def("transformation_matrix_to_super",
           [](Conversions const &c) {
             py::array_t<int64_t> a({3, 3});
             for (int i = 0; i < 9; ++i) a.mutable_data()[i] = c.transformation_matrix_to_super()[i];
             return a;
           })
      .def("unit_transformation_matrix_to_super",
           [](Conversions const &c) {
             py::array_t<int64_t> a({3, 3});
             for (int i = 0; i < 9; ++i) a.mutable_data()[i] = c.unit_transformation_matrix_to_super()[i];
             return a;
           })
      .def("occ_size", &Conversions::occ_size)
      .def("occ_to_species_index", [](Conversions const &c, Index asym, Index occ) { return c.species_index(asym, occ); })
      .def("species_to_occ_index", [](Conversions const &c, Index asym, Index sp) { return c.occ_index(asym, sp); })
      .def("species_allowed", &Conversions::species_allowed)
      .def("species_size", &Conversions::species_size)
      .def("species_name_to_index", [](Conversions const &c, std::string name) { return c.species_index(name); })
      .def("species_index_to_name", [](Conversions const &c, Index sp) { return c.species_name(sp); })
      .def("species_list", [](Conversions const &c) { return c.species_list(); })
      .def("species_index_to_atoms_size", &Conversions::components_size)
      .def("l_to_bijk_batch", [](Conversions const &c, std::vector<int64_t> l) {
        auto out = c.l_to_bijk_batch(l);
        py::array_t<int64_t> a({static_cast<py::ssize_t>(l.size()), static_cast<py::ssize_t>(4)});
        std::copy(out.begin(), out.end(), a.mutable_data());
        return a;
      })
      .def("bijk_to_l_batch", [](Conversions const &c,
                                 py::array_t<int64_t, py::array::c_style | py::array::forcecast> bijk) {
        std::vector<int64_t> in(bijk.data(), bijk.data() + bijk.size());
        return c.bijk_to_l_batch(in);
      });

  // --------------------------------------------------------------- ising_cpp
  py::class_<IsingConfiguration>(m, "IsingConfiguration")
      .def(py::init([](py::object shape, int fill_value) {
             return IsingConfiguration(to_ivec(shape), fill_value);
           }),
           py::arg("shape") = std::vector<int>{0, 0}, py::arg("fill_value") = 1)
      .def_property("shape", [](IsingConfiguration const &c) { return to_long_array(c.shape); },
                    [](IsingConfiguration &c, py::object s) { c.shape = to_ivec(s); })
      .def_readonly("n_sites", &IsingConfiguration::n_sites)
      .def_readonly("n_variable_sites", &IsingConfiguration::n_variable_sites)
      .def_readonly("n_unitcells", &IsingConfiguration::n_unitcells)
      .def("occupation", [](IsingConfiguration const &c) {
        auto const &o = c.occupation();
        py::array_t<int32_t> a(o.size());
        std::copy(o.begin(), o.end(), a.mutable_data());
        return a;
      })
      .def("set_occupation", [](IsingConfiguration &c, py::object occ) { c.set_occupation(to_ivec(occ)); },
           py::arg("occupation"))
      .def("occ", &IsingConfiguration::occ, py::arg("linear_site_index"))
      .def("set_occ", &IsingConfiguration::set_occ, py::arg("linear_site_index"), py::arg("new_occ"))
      .def("within", &IsingConfiguration::within, py::arg("index"), py::arg("dim"))
      .def("from_linear_site_index", [](IsingConfiguration const &c, Index l) {
        return to_long_array(c.from_linear_site_index(l));
      }, py::arg("linear_site_index"))
      .def("to_linear_site_index", [](IsingConfiguration const &c, py::object mi) {
        return c.to_linear_site_index(to_ivec(mi));
      }, py::arg("multi_index"))
      .def("to_dict", &config_to_dict)
      .def_static("from_dict", &config_from_dict, py::arg("data"))
      .def("__copy__", [](IsingConfiguration const &c) { return IsingConfiguration(c); })
      .def("__deepcopy__", [](IsingConfiguration const &c, py::dict) { return IsingConfiguration(c); });

  py::class_<IsingState>(m, "IsingState")
      .def(py::init([](IsingConfiguration const &configuration, ValueMap const &conditions,
                       std::optional<ValueMap> properties) {
             return IsingState(configuration, conditions, properties.value_or(ValueMap()));
           }),
           py::arg("configuration"), py::arg("conditions"), py::arg("properties") = py::none())
      .def_readwrite("configuration", &IsingState::configuration)
      .def_readwrite("conditions", &IsingState::conditions)
      .def_readwrite("properties", &IsingState::properties)
      .def("to_dict", [](IsingState const &s) {
        py::dict d;
        d["configuration"] = config_to_dict(s.configuration);
        d["conditions"] = valuemap_to_dict(s.conditions);
        d["properties"] = valuemap_to_dict(s.properties);
        return d;
      })
      .def_static("from_dict", [](py::dict d) {
        return IsingState(config_from_dict(d["configuration"].cast<py::dict>()),
                          valuemap_from_dict(d["conditions"].cast<py::dict>()),
                          d.contains("properties") ? valuemap_from_dict(d["properties"].cast<py::dict>())
                                                   : ValueMap());
      }, py::arg("data"));

  py::class_<IsingFormationEnergy>(m, "IsingFormationEnergy")
      .def(py::init([](double J, int lattice_type, bool use_nlist, IsingState const *state) {
             IsingFormationEnergy f(J, lattice_type, use_nlist);
             return f;
           }),
           py::arg("J") = 1.0, py::arg("lattice_type") = 1, py::arg("use_nlist") = true,
           py::arg("state") = nullptr)
      .def_readwrite("J", &IsingFormationEnergy::J)
      .def_readwrite("lattice_type", &IsingFormationEnergy::lattice_type)
      .def("set_state", &IsingFormationEnergy::set_state, py::arg("state"), py::keep_alive<1, 2>())
      .def("per_supercell", &IsingFormationEnergy::per_supercell)
      .def("per_unitcell", &IsingFormationEnergy::per_unitcell)
      .def("occ_delta_per_supercell", [](IsingFormationEnergy const &f, std::vector<long> l, std::vector<int> o) {
        return f.occ_delta_per_supercell(l, o);
      }, py::arg("linear_site_index"), py::arg("new_occ"));

  py::class_<IsingParamComposition>(m, "IsingParamComposition")
      .def(py::init([](IsingState const *state) { return IsingParamComposition(); }), py::arg("state") = nullptr)
      .def("set_state", &IsingParamComposition::set_state, py::arg("state"), py::keep_alive<1, 2>())
      .def("n_independent_compositions", &IsingParamComposition::n_independent_compositions)
      .def("per_supercell", [](IsingParamComposition const &c) { return as_array(c.per_supercell()); })
      .def("per_unitcell", [](IsingParamComposition const &c) { return as_array(c.per_unitcell()); })
      .def("occ_delta_per_supercell", [](IsingParamComposition const &c, std::vector<long> l, std::vector<int> o) {
        return as_array(c.occ_delta_per_supercell(l, o));
      }, py::arg("linear_site_index"), py::arg("new_occ"));

  py::class_<IsingSystem, std::shared_ptr<IsingSystem>>(m, "IsingSystem")
      .def(py::init<IsingFormationEnergy, IsingParamComposition>(), py::arg("formation_energy_calculator"),
           py::arg("param_composition_calculator"))
      .def_readwrite("formation_energy_calculator", &IsingSystem::formation_energy_calculator)
      .def_readwrite("param_composition_calculator", &IsingSystem::param_composition_calculator);

  // ----------------------------------------------------------------- sampling
  py::class_<Sampler, std::shared_ptr<Sampler>>(m, "Sampler")
      .def(py::init([](std::vector<Index> shape, std::optional<std::vector<std::string>> component_names,
                       CountType capacity_increment) {
             if (component_names.has_value())
               return std::make_shared<Sampler>(shape, *component_names, capacity_increment);
             return std::make_shared<Sampler>(shape, capacity_increment);
           }),
           py::arg("shape"), py::arg("component_names") = py::none(), py::arg("capacity_increment") = 1000)
      .def("append", [](Sampler &s, py::object v) { s.push_back(to_dvec(v)); }, py::arg("vector"))
      .def("set_values", [](Sampler &s, py::array_t<double, py::array::c_style | py::array::forcecast> a) {
        if (a.ndim() != 2) throw std::runtime_error("set_values expects a 2d array");
        std::vector<std::vector<double>> rows(a.shape(0), std::vector<double>(a.shape(1)));
        for (py::ssize_t r = 0; r < a.shape(0); ++r)
          for (py::ssize_t c = 0; c < a.shape(1); ++c) rows[r][c] = a.at(r, c);
        s.set_values(rows);
      })
      .def("clear", &Sampler::clear)
      .def("set_sample_capacity", &Sampler::set_sample_capacity)
      .def("set_capacity_increment", &Sampler::set_capacity_increment)
      .def("component_names", &Sampler::component_names)
      .def("shape", &Sampler::shape)
      .def("n_components", &Sampler::n_components)
      .def("n_samples", &Sampler::n_samples)
      .def("sample_capacity", &Sampler::sample_capacity)
      .def("values", [](Sampler const &s) {
        py::array_t<double> a({static_cast<py::ssize_t>(s.n_samples()), static_cast<py::ssize_t>(s.n_components())});
        auto w = a.mutable_unchecked<2>();
        for (Index c = 0; c < s.n_components(); ++c)
          for (CountType r = 0; r < s.n_samples(); ++r) w(r, c) = s.component_data(c)[r];
        return a;
      })
      .def("component", [](Sampler const &s, Index i) { return as_array(s.component(i)); })
      .def("sample", [](Sampler const &s, CountType i) { return as_array(s.sample(i)); });
  py::bind_map<SamplerMap>(m, "SamplerMap");
  m.def("get_n_samples", &get_n_samples, py::arg("samplers"));
  m.def("scalar_as_vector", [](double v) { return as_array(std::vector<double>{v}); });
  m.def("vector_as_vector", [](py::object v) { return as_array(to_dvec(v)); });
  m.def("matrix_as_vector", [](py::array_t<double, py::array::f_style | py::array::forcecast> a) {
    return as_array(std::vector<double>(a.data(), a.data() + a.size()));  // column-major unrolling
  });
  m.def("default_component_names", &default_component_names, py::arg("shape"));
  m.def("colmajor_component_names", &colmajor_component_names, py::arg("n_rows"), py::arg("n_cols"));

  py::class_<SamplerComponent>(m, "SamplerComponent")
      .def(py::init<std::string, Index, std::string>(), py::arg("sampler_name"), py::arg("component_index"),
           py::arg("component_name"))
      .def_readwrite("sampler_name", &SamplerComponent::sampler_name)
      .def_readwrite("component_index", &SamplerComponent::component_index)
      .def_readwrite("component_name", &SamplerComponent::component_name)
      .def("__lt__", [](SamplerComponent const &a, SamplerComponent const &b) { return a < b; })
      .def("__eq__", [](SamplerComponent const &a, SamplerComponent const &b) { return !(a < b) && !(b < a); })
      .def("__hash__", [](SamplerComponent const &a) {
        return py::hash(py::make_tuple(a.sampler_name, a.component_index));
      });

  py::class_<StateSamplingFunction>(m, "StateSamplingFunction")
      .def(py::init([](std::string name, std::string description, std::vector<Index> shape, py::object function,
                       std::optional<std::vector<std::string>> component_names) {
             auto f = [function]() -> std::vector<double> {
               py::gil_scoped_acquire gil;
               return to_dvec(function());
             };
             return StateSamplingFunction(name, description, shape, f, component_names);
           }),
           py::arg("name"), py::arg("description"), py::arg("shape"), py::arg("function"),
           py::arg("component_names") = py::none())
      .def_readwrite("name", &StateSamplingFunction::name)
      .def_readwrite("description", &StateSamplingFunction::description)
      .def_readwrite("shape", &StateSamplingFunction::shape)
      .def_readwrite("component_names", &StateSamplingFunction::component_names)
      .def("__call__", [](StateSamplingFunction const &f) { return as_array(f()); });
  py::bind_map<StateSamplingFunctionMap>(m, "StateSamplingFunctionMap");

  py::class_<PyJsonStateSamplingFunction>(m, "jsonStateSamplingFunction")
      .def(py::init([](std::string name, std::string description, py::object function) {
             return PyJsonStateSamplingFunction{name, description, function};
           }),
           py::arg("name"), py::arg("description"), py::arg("function"))
      .def_readwrite("name", &PyJsonStateSamplingFunction::name)
      .def_readwrite("description", &PyJsonStateSamplingFunction::description)
      .def_readwrite("function", &PyJsonStateSamplingFunction::function)
      .def("__call__", [](PyJsonStateSamplingFunction const &f) { return f.function(); });
  py::bind_map<PyJsonStateSamplingFunctionMap>(m, "jsonStateSamplingFunctionMap");
  py::class_<PyJsonSampler>(m, "jsonSampler")
      .def(py::init<>())
      .def_readwrite("values", &PyJsonSampler::values)
      .def("to_list", [](PyJsonSampler const &s) { return py::list(s.values); });
  py::bind_map<PyJsonSamplerMap>(m, "jsonSamplerMap");

  py::class_<RequestedPrecision>(m, "RequestedPrecision")
      .def(py::init([](std::optional<double> abs, std::optional<double> rel) {
             RequestedPrecision r;
             if (abs.has_value()) {
               r.abs_convergence_is_required = true;
               r.abs_precision = *abs;
             }
             if (rel.has_value()) {
               r.rel_convergence_is_required = true;
               r.rel_precision = *rel;
             }
             return r;
           }),
           py::arg("abs") = py::none(), py::arg("rel") = py::none())
      .def_readwrite("abs_convergence_is_required", &RequestedPrecision::abs_convergence_is_required)
      .def_readwrite("abs_precision", &RequestedPrecision::abs_precision)
      .def_readwrite("rel_convergence_is_required", &RequestedPrecision::rel_convergence_is_required)
      .def_readwrite("rel_precision", &RequestedPrecision::rel_precision)
      .def("to_dict", &req_prec_to_dict)
      .def_static("from_dict", &req_prec_from_dict, py::arg("data"));
  py::bind_map<RequestedPrecisionMap>(m, "RequestedPrecisionMap");

  py::class_<BasicStatistics>(m, "BasicStatistics")
      .def(py::init<>())
      .def_readwrite("mean", &BasicStatistics::mean)
      .def_readwrite("calculated_precision", &BasicStatistics::calculated_precision)
      .def("relative_precision", [](BasicStatistics const &s) { return get_calculated_relative_precision(s); })
      .def("to_dict", &stats_to_dict);
  py::class_<BasicStatisticsCalculator>(m, "BasicStatisticsCalculator")
      .def(py::init<double, Index, Index>(), py::arg("confidence") = 0.95,
           py::arg("weighted_observations_method") = 1, py::arg("n_resamples") = 10000)
      .def_readwrite("confidence", &BasicStatisticsCalculator::confidence)
      .def_readwrite("weighted_observations_method", &BasicStatisticsCalculator::method)
      .def_readwrite("n_resamples", &BasicStatisticsCalculator::n_resamples)
      .def("__call__", [](BasicStatisticsCalculator const &c, py::object obs, py::object w) {
        return c(to_dvec(obs), w.is_none() ? std::vector<double>() : to_dvec(w));
      }, py::arg("observations"), py::arg("sample_weight") = py::none())
      .def("calculate", [](BasicStatisticsCalculator const &c, py::object obs, py::object w) {
        return c(to_dvec(obs), w.is_none() ? std::vector<double>() : to_dvec(w));
      }, py::arg("observations"), py::arg("sample_weight") = py::none())
      .def("to_dict", [](BasicStatisticsCalculator const &c) {
        py::dict d;
        d["confidence"] = c.confidence;
        d["weighted_observations_method"] = c.method;
        d["n_resamples"] = c.n_resamples;
        return d;
      })
      // python/src/monte_sampling.cpp:2363: keys as to_dict, defaults as the constructor
      .def_static("from_dict", [](py::dict d) {
        BasicStatisticsCalculator c;
        if (d.contains("confidence")) c.confidence = d["confidence"].cast<double>();
        if (d.contains("weighted_observations_method")) c.method = d["weighted_observations_method"].cast<Index>();
        if (d.contains("n_resamples")) c.n_resamples = d["n_resamples"].cast<Index>();
        return c;
      }, py::arg("data"));

  // C++-only in the reference (BasicStatistics.hh:57-59); exposed for the parity tests
  m.def("resample", [](py::object obs, py::object w, double weight_sum, Index n) {
    return resample(to_dvec(obs), to_dvec(w), weight_sum, n);
  }, py::arg("observations"), py::arg("sample_weight"), py::arg("sample_weight_sum"),
     py::arg("n_equally_spaced"));

  py::class_<IndividualEquilibrationCheckResult>(m, "IndividualEquilibrationResult")
      .def(py::init<>())
      .def_readwrite("is_equilibrated", &IndividualEquilibrationCheckResult::is_equilibrated)
      .def_readwrite("N_samples_for_equilibration", &IndividualEquilibrationCheckResult::N_samples_for_equilibration);
  py::class_<EquilibrationCheckResults>(m, "EquilibrationCheckResults")
      .def(py::init<>())
      .def_readwrite("all_equilibrated", &EquilibrationCheckResults::all_equilibrated)
      .def_readwrite("N_samples_for_all_to_equilibrate", &EquilibrationCheckResults::N_samples_for_all_to_equilibrate)
      .def_readwrite("individual_results", &EquilibrationCheckResults::individual_results)
      .def("to_dict", &eq_results_to_dict);
  m.def("default_equilibration_check", [](py::object obs, py::object w, RequestedPrecision rp) {
    return default_equilibration_check(to_dvec(obs), w.is_none() ? std::vector<double>() : to_dvec(w), rp);
  }, py::arg("observations"), py::arg("sample_weight"), py::arg("requested_precision"));

  py::class_<IndividualConvergenceCheckResult>(m, "IndividualConvergenceResult")
      .def(py::init<>())
      .def_readwrite("is_converged", &IndividualConvergenceCheckResult::is_converged)
      .def_readwrite("requested_precision", &IndividualConvergenceCheckResult::requested_precision)
      .def_readwrite("stats", &IndividualConvergenceCheckResult::stats);
  py::class_<ConvergenceCheckResults>(m, "ConvergenceCheckResults")
      .def(py::init<>())
      .def_readwrite("all_converged", &ConvergenceCheckResults::all_converged)
      .def_readwrite("N_samples_for_statistics", &ConvergenceCheckResults::N_samples_for_statistics)
      .def_readwrite("individual_results", &ConvergenceCheckResults::individual_results)
      .def("to_dict", &conv_results_to_dict);

  py::class_<CutoffCheckParams>(m, "CutoffCheckParams")
      // python/src/monte_sampling.cpp make_cutoff_check_params
      .def(py::init([](std::optional<CountType> min_count, std::optional<CountType> max_count,
                       std::optional<TimeType> min_time, std::optional<TimeType> max_time,
                       std::optional<CountType> min_sample, std::optional<CountType> max_sample,
                       std::optional<TimeType> min_clocktime, std::optional<TimeType> max_clocktime) {
             CutoffCheckParams p;
             p.min_count = min_count;
             p.max_count = max_count;
             p.min_time = min_time;
             p.max_time = max_time;
             p.min_sample = min_sample;
             p.max_sample = max_sample;
             p.min_clocktime = min_clocktime;
             p.max_clocktime = max_clocktime;
             return p;
           }),
           py::arg("min_count") = std::nullopt, py::arg("max_count") = std::nullopt,
           py::arg("min_time") = std::nullopt, py::arg("max_time") = std::nullopt,
           py::arg("min_sample") = std::nullopt, py::arg("max_sample") = std::nullopt,
           py::arg("min_clocktime") = std::nullopt, py::arg("max_clocktime") = std::nullopt)
      .def("to_dict", &cutoff_to_dict)
      .def_static("from_dict", &cutoff_from_dict, py::arg("data"))
      .def_readwrite("min_count", &CutoffCheckParams::min_count)
      .def_readwrite("min_time", &CutoffCheckParams::min_time)
      .def_readwrite("min_sample", &CutoffCheckParams::min_sample)
      .def_readwrite("min_clocktime", &CutoffCheckParams::min_clocktime)
      .def_readwrite("max_count", &CutoffCheckParams::max_count)
      .def_readwrite("max_time", &CutoffCheckParams::max_time)
      .def_readwrite("max_sample", &CutoffCheckParams::max_sample)
      .def_readwrite("max_clocktime", &CutoffCheckParams::max_clocktime);
  m.def("all_minimums_met", &all_minimums_met);
  m.def("any_maximum_met", &any_maximum_met);

  py::class_<CompletionCheckParams>(m, "CompletionCheckParams")
      // python/src/monte_sampling.cpp:222-283 make_completion_check_params
      .def(py::init([](std::optional<RequestedPrecisionMap> requested_precision,
                       std::optional<CutoffCheckParams> cutoff_params, py::object calc_statistics_f,
                       py::object equilibration_check_f, bool log_spacing, std::optional<CountType> check_begin,
                       std::optional<CountType> check_period, std::optional<double> check_base,
                       std::optional<double> check_shift, std::optional<CountType> check_period_max) {
             CompletionCheckParams result;  // default check functions
             if (!log_spacing) {
               if (!check_period.has_value()) check_period = 100;
               result.check_begin = *check_period;
               result.check_period = *check_period;
             } else {
               result.check_begin = 0;
               result.check_base = 10.0;
               result.check_shift = 2.0;
               result.check_period_max = 10000;
             }
             if (cutoff_params.has_value()) result.cutoff_params = *cutoff_params;
             if (requested_precision.has_value()) result.requested_precision = *requested_precision;
             result.log_spacing = log_spacing;
             if (check_begin.has_value()) result.check_begin = *check_begin;
             if (check_period.has_value()) result.check_period = *check_period;
             if (check_base.has_value()) result.check_base = *check_base;
             if (check_shift.has_value()) result.check_shift = *check_shift;
             if (check_period_max.has_value()) result.check_period_max = *check_period_max;
             py::object self = py::cast(result);
             if (!calc_statistics_f.is_none()) self.attr("calc_statistics_f") = calc_statistics_f;
             if (!equilibration_check_f.is_none()) self.attr("equilibration_check_f") = equilibration_check_f;
             return self.cast<CompletionCheckParams>();
           }),
           py::arg("requested_precision") = std::nullopt, py::arg("cutoff_params") = std::nullopt,
           py::arg("calc_statistics_f") = py::none(), py::arg("equilibration_check_f") = py::none(),
           py::arg("log_spacing") = false, py::arg("check_begin") = std::nullopt,
           py::arg("check_period") = std::nullopt, py::arg("check_base") = std::nullopt,
           py::arg("check_shift") = std::nullopt, py::arg("check_period_max") = std::nullopt)
      .def("to_dict", &completion_check_params_to_dict)
      .def_static("from_dict", &completion_check_params_from_dict, py::arg("data"), py::arg("sampling_functions"))
      .def_readwrite("cutoff_params", &CompletionCheckParams::cutoff_params)
      .def_readwrite("requested_precision", &CompletionCheckParams::requested_precision)
      .def_readwrite("log_spacing", &CompletionCheckParams::log_spacing)
      .def_readwrite("check_begin", &CompletionCheckParams::check_begin)
      .def_readwrite("check_period", &CompletionCheckParams::check_period)
      .def_readwrite("check_base", &CompletionCheckParams::check_base)
      .def_readwrite("check_shift", &CompletionCheckParams::check_shift)
      .def_readwrite("check_period_max", &CompletionCheckParams::check_period_max)
      .def_property("calc_statistics_f", [](CompletionCheckParams const &) { return py::none(); },
                    [](CompletionCheckParams &p, py::object f) {
                      if (py::isinstance<BasicStatisticsCalculator>(f)) {
                        p.calc_statistics_f = f.cast<BasicStatisticsCalculator>();
                      } else {
                        p.calc_statistics_f = [f](std::vector<double> const &o, std::vector<double> const &w) {
                          py::gil_scoped_acquire gil;
                          return f(as_array(o), as_array(w)).cast<BasicStatistics>();
                        };
                      }
                    })
      .def_property("equilibration_check_f", [](CompletionCheckParams const &) { return py::none(); },
                    [](CompletionCheckParams &p, py::object f) {
                      p.equilibration_check_f = [f](std::vector<double> const &o, std::vector<double> const &w,
                                                    RequestedPrecision rp) {
                        py::gil_scoped_acquire gil;
                        return f(as_array(o), as_array(w), rp).cast<IndividualEquilibrationCheckResult>();
                      };
                    });

  py::class_<CompletionCheckResults>(m, "CompletionCheckResults")
      .def(py::init<>())
      .def_readwrite("params", &CompletionCheckResults::params)
      .def_readwrite("count", &CompletionCheckResults::count)
      .def_readwrite("time", &CompletionCheckResults::time)
      .def_readwrite("clocktime", &CompletionCheckResults::clocktime)
      .def_readwrite("n_samples", &CompletionCheckResults::n_samples)
      .def_readwrite("has_all_minimums_met", &CompletionCheckResults::has_all_minimums_met)
      .def_readwrite("has_any_maximum_met", &CompletionCheckResults::has_any_maximum_met)
      .def_readwrite("n_samples_at_convergence_check", &CompletionCheckResults::n_samples_at_convergence_check)
      .def_readwrite("equilibration_check_results", &CompletionCheckResults::equilibration_check_results)
      .def_readwrite("convergence_check_results", &CompletionCheckResults::convergence_check_results)
      .def_readwrite("is_complete", &CompletionCheckResults::is_complete)
      .def("partial_reset", [](CompletionCheckResults &r) { r.partial_reset(); })
      .def("full_reset", [](CompletionCheckResults &r) { r.full_reset(); })
      .def("to_dict", &cc_results_to_dict);

  py::class_<CompletionCheck>(m, "CompletionCheck")
      .def(py::init<CompletionCheckParams>(), py::arg("params"))
      .def("reset", &CompletionCheck::reset)
      .def("params", &CompletionCheck::params, py::return_value_policy::reference_internal)
      .def("results", &CompletionCheck::results, py::return_value_policy::reference_internal)
      .def("n_checks", &CompletionCheck::n_checks)
      .def("count_check", [](CompletionCheck &c, SamplerMap const &samplers, Sampler const &w, CountType count,
                             MethodLog &log) { return c.is_complete(samplers, w, count, log.log); },
           py::arg("samplers"), py::arg("sample_weight"), py::arg("count"), py::arg("method_log"))
      .def("check", [](CompletionCheck &c, SamplerMap const &samplers, Sampler const &w, MethodLog &log) {
        return c.is_complete(samplers, w, log.log);
      }, py::arg("samplers"), py::arg("sample_weight"), py::arg("method_log"))
      .def("time_check", [](CompletionCheck &c, SamplerMap const &samplers, Sampler const &w, double time,
                            MethodLog &log) { return c.is_complete_time(samplers, w, time, log.log); },
           py::arg("samplers"), py::arg("sample_weight"), py::arg("time"), py::arg("method_log"))
      .def("count_and_time_check", [](CompletionCheck &c, SamplerMap const &samplers, Sampler const &w,
                                      CountType count, double time, MethodLog &log) {
        return c.is_complete(samplers, w, count, time, log.log);
      }, py::arg("samplers"), py::arg("sample_weight"), py::arg("count"), py::arg("time"), py::arg("method_log"))
      .def("__call__", [](CompletionCheck &c, SamplerMap const &samplers, Sampler const &w,
                          std::optional<CountType> count, std::optional<double> time, MethodLog &log) {
        if (count && time) return c.is_complete(samplers, w, *count, *time, log.log);
        if (count) return c.is_complete(samplers, w, *count, log.log);
        if (time) return c.is_complete_time(samplers, w, *time, log.log);
        return c.is_complete(samplers, w, log.log);
      }, py::arg("samplers"), py::arg("sample_weight"), py::arg("count") = py::none(),
           py::arg("time") = py::none(), py::arg("method_log"));

  // ------------------------------------------------------------------ methods
  m.def("metropolis_acceptance", [](double dE, double beta, RandomNumberGenerator<> &rng) {
    return metropolis_acceptance(dE, beta, rng);
  }, py::arg("delta_potential_energy"), py::arg("beta"), py::arg("random_number_generator"));

  // python/src/monte_methods.cpp:196-263: the generic loop with Python callbacks
  m.def("basic_occupation_metropolis",
        [](std::shared_ptr<SemiGrandCanonicalData> data, double temperature, py::object dpotential_f,
           py::object propose_event_f, py::object apply_event_f, int sample_period,
           std::optional<MethodLog> method_log, std::optional<PyEngine> random_engine,
           py::object write_status_f) {
          auto holder = std::make_shared<OccEvent>();
          std::function<double(OccEvent const &)> dpot = [dpotential_f](OccEvent const &e) {
            return dpotential_f(py::cast(&e, py::return_value_policy::reference)).cast<double>();
          };
          std::function<OccEvent const &(RandomNumberGenerator<> &)> propose =
              [propose_event_f, holder](RandomNumberGenerator<> &rng) -> OccEvent const & {
            *holder = propose_event_f(py::cast(&rng, py::return_value_policy::reference)).cast<OccEvent>();
            return *holder;
          };
          std::function<void(OccEvent const &)> apply = [apply_event_f](OccEvent const &e) {
            apply_event_f(py::cast(&e, py::return_value_policy::reference));
          };
          std::function<void(BasicOccupationMetropolisData const &, MethodLog &)> wsf;
          if (write_status_f.is_none()) {
            wsf = [](BasicOccupationMetropolisData const &d, MethodLog &log) {
              default_write_run_status(d, log, std::cout);
              default_finish_write_status(d, log);
            };
          } else {
            wsf = [write_status_f, data](BasicOccupationMetropolisData const &, MethodLog &log) {
              write_status_f(data, py::cast(&log, py::return_value_policy::reference));
            };
          }
          basic_occupation_metropolis<default_engine_type>(
              *data, temperature, dpot, propose, apply, sample_period, method_log,
              random_engine.has_value() ? random_engine->e : nullptr, wsf);
        },
        py::arg("data"), py::arg("temperature"), py::arg("potential_occ_delta_per_supercell_f"),
        py::arg("propose_event_f"), py::arg("apply_event_f"), py::arg("sample_period") = 1,
        py::arg("method_log") = py::none(), py::arg("random_engine") = py::none(),
        py::arg("write_status_f") = py::none());

  // ------------------------------------------------- ising_cpp.semigrand_canonical
  py::class_<SemiGrandCanonicalConditions, std::shared_ptr<SemiGrandCanonicalConditions>>(
      m, "SemiGrandCanonicalConditions")
      .def(py::init([](double temperature, py::object exchange_potential) {
             return SemiGrandCanonicalConditions(temperature, to_dvec(exchange_potential));
           }),
           py::arg("temperature"), py::arg("exchange_potential"))
      .def_readwrite("temperature", &SemiGrandCanonicalConditions::temperature)
      .def_property("exchange_potential",
                    [](SemiGrandCanonicalConditions const &c) { return as_array(c.exchange_potential); },
                    [](SemiGrandCanonicalConditions &c, py::object v) { c.exchange_potential = to_dvec(v); })
      .def("to_values", &SemiGrandCanonicalConditions::to_values)
      .def_static("from_values", &SemiGrandCanonicalConditions::from_values, py::arg("values"))
      .def("to_dict", [](SemiGrandCanonicalConditions const &c) { return valuemap_to_dict(c.to_values()); })
      .def_static("from_dict", [](py::dict d) {
        return SemiGrandCanonicalConditions::from_values(valuemap_from_dict(d));
      }, py::arg("data"));

  py::class_<SemiGrandCanonicalPotential>(m, "SemiGrandCanonicalPotential")
      .def(py::init<std::shared_ptr<IsingSystem>>(), py::arg("system"))
      .def("set_state", &SemiGrandCanonicalPotential::set_state, py::arg("state"), py::arg("conditions"),
           py::keep_alive<1, 2>())
      .def("per_supercell", &SemiGrandCanonicalPotential::per_supercell)
      .def("per_unitcell", &SemiGrandCanonicalPotential::per_unitcell)
      .def("occ_delta_per_supercell", [](SemiGrandCanonicalPotential const &p, std::vector<long> l, std::vector<int> o) {
        return p.occ_delta_per_supercell(l, o);
      }, py::arg("linear_site_index"), py::arg("new_occ"))
      .def("occ_event_delta_per_supercell", [](SemiGrandCanonicalPotential const &p, OccEvent const &e) {
        return p.occ_delta_per_supercell(e);
      }, py::arg("occ_event"));

  py::class_<SemiGrandCanonicalData, std::shared_ptr<SemiGrandCanonicalData>>(m, "SemiGrandCanonicalData")
      .def(py::init([](StateSamplingFunctionMap const &sf, PyJsonStateSamplingFunctionMap const &jsf,
                       CountType n_steps_per_pass, CompletionCheckParams const &p) {
             auto d = std::make_shared<SemiGrandCanonicalData>(sf, n_steps_per_pass, p);
             auto ex = std::make_shared<RunExtras>();
             ex->json_sampling_functions = jsf;
             for (auto const &kv : jsf) ex->json_samplers.emplace(kv.first, PyJsonSampler());
             register_extras(d.get(), ex);
             return d;
           }),
           py::arg("sampling_functions"), py::arg("json_sampling_functions"), py::arg("n_steps_per_pass"),
           py::arg("completion_check_params"))
      .def_readwrite("sampling_functions", &SemiGrandCanonicalData::sampling_functions)
      .def_readwrite("samplers", &SemiGrandCanonicalData::samplers)
      .def_property_readonly("json_sampling_functions", [](SemiGrandCanonicalData const &d) {
        return extras_registry().at(&d)->json_sampling_functions;
      })
      .def_property_readonly("json_samplers", [](SemiGrandCanonicalData const &d) -> PyJsonSamplerMap & {
        return extras_registry().at(&d)->json_samplers;
      }, py::return_value_policy::reference)
      .def_readwrite("sample_weight", &SemiGrandCanonicalData::sample_weight)
      .def_readwrite("n_pass", &SemiGrandCanonicalData::n_pass)
      .def_readwrite("n_steps_per_pass", &SemiGrandCanonicalData::n_steps_per_pass)
      .def_readwrite("n_accept", &SemiGrandCanonicalData::n_accept)
      .def_readwrite("n_reject", &SemiGrandCanonicalData::n_reject)
      .def_readonly("completion_check", &SemiGrandCanonicalData::completion_check)
      .def("acceptance_rate", &SemiGrandCanonicalData::acceptance_rate)
      .def("rejection_rate", &SemiGrandCanonicalData::rejection_rate)
      .def("reset", &SemiGrandCanonicalData::reset)
      .def("to_dict", [](SemiGrandCanonicalData const &d) {
        // basic_occupation_metropolis.hh:168-180
        py::dict out;
        out["completion_check_results"] = cc_results_to_dict(d.completion_check.results());
        out["n_pass"] = d.n_pass;
        out["n_steps_per_pass"] = d.n_steps_per_pass;
        out["n_accept"] = static_cast<long>(d.n_accept);
        out["n_reject"] = static_cast<long>(d.n_reject);
        out["acceptance_rate"] = d.acceptance_rate();
        out["rejection_rate"] = d.rejection_rate();
        return out;
      });

  py::class_<event_generator_type>(m, "SemiGrandCanonicalEventGenerator")
      .def(py::init<>())
      .def("set_state", &event_generator_type::set_state, py::arg("state"), py::keep_alive<1, 2>())
      .def("propose", &event_generator_type::propose, py::arg("random_number_generator"),
           py::return_value_policy::reference_internal)
      .def("apply", &event_generator_type::apply, py::arg("occ_event"));

  py::class_<calculator_type, std::shared_ptr<calculator_type>>(m, "SemiGrandCanonicalCalculator")
      .def(py::init<std::shared_ptr<IsingSystem>>(), py::arg("system"))
      .def_readonly("system", &calculator_type::system)
      .def_property_readonly("state", [](calculator_type const &c) { return c.state; },
                             py::return_value_policy::reference_internal)
      .def_readonly("conditions", &calculator_type::conditions)
      .def_readonly("potential", &calculator_type::potential)
      .def_property_readonly("formation_energy_calculator",
                             [](calculator_type const &c) { return c.formation_energy_calculator; },
                             py::return_value_policy::reference_internal)
      .def_property_readonly("param_composition_calculator",
                             [](calculator_type const &c) { return c.param_composition_calculator; },
                             py::return_value_policy::reference_internal)
      .def_readonly("data", &calculator_type::data)
      .def_readwrite("update_mode", &calculator_type::update_mode)
      .def_readwrite("overlap_checks", &calculator_type::overlap_checks)
      .def_readonly("last_kernel", &calculator_type::last_kernel)
      .def("default_sampling_functions", [](std::shared_ptr<calculator_type> mc) {
        StateSamplingFunctionMap fns;
        for (auto const &f : {make_parametric_composition_f(mc), make_formation_energy_f(mc),
                              make_potential_energy_f(mc)})
          fns.emplace(f.name, f);
        return fns;
      })
      .def("default_json_sampling_functions", [](py::object self) {
        PyJsonStateSamplingFunctionMap fns;
        py::object weak = py::module_::import("weakref").attr("ref")(self);
        py::object f = py::cpp_function([weak]() -> py::object {
          py::object mc = weak();
          if (mc.is_none()) throw std::runtime_error("Error in configuration sampling function: mc_calculator == nullptr");
          auto &c = mc.cast<calculator_type &>();
          if (c.state == nullptr)
            throw std::runtime_error("Error in configuration sampling function: mc_calculator->state == nullptr");
          return config_to_dict(c.state->configuration);
        });
        fns.emplace("configuration", PyJsonStateSamplingFunction{"configuration", "Configuration values", f});
        return fns;
      })
      .def("run", [](std::shared_ptr<calculator_type> mc, IsingState &state,
                     StateSamplingFunctionMap const &sampling_functions,
                     PyJsonStateSamplingFunctionMap const &json_sampling_functions,
                     CompletionCheckParams const &completion_check_params,
                     event_generator_type const &event_generator, int sample_period,
                     std::optional<MethodLog> method_log, std::optional<PyEngine> random_engine,
                     py::object write_status_f, std::optional<std::string> update_mode) {
        if (update_mode.has_value()) mc->update_mode = *update_mode;
        auto extras = std::make_shared<RunExtras>();
        extras->json_sampling_functions = json_sampling_functions;
        for (auto const &kv : json_sampling_functions) extras->json_samplers.emplace(kv.first, PyJsonSampler());
        calculator_type::write_status_type wsf;
        if (write_status_f.is_none()) {
          wsf = default_write_status;
        } else {
          wsf = [write_status_f, mc](calculator_type const &, MethodLog &log) {
            write_status_f(mc, py::cast(&log, py::return_value_policy::reference));
          };
        }
        std::function<void()> json_hook;
        if (!json_sampling_functions.empty()) {
          json_hook = [extras]() {
            for (auto &kv : extras->json_sampling_functions)
              extras->json_samplers.at(kv.first).values.append(kv.second.function());
          };
        }
        // `data` is created inside run; register the extras as soon as it exists
        // by wrapping the status writer and the hook
        auto reg = [mc, extras]() {
          if (mc->data) register_extras(mc->data.get(), extras);
        };
        calculator_type::write_status_type wsf2 = [wsf, reg](calculator_type const &c, MethodLog &log) {
          reg();
          wsf(c, log);
        };
        std::function<void()> hook2;
        if (json_hook) hook2 = [json_hook, reg]() { reg(); json_hook(); };
        mc->run(state, sampling_functions, completion_check_params, event_generator, sample_period, method_log,
                random_engine.has_value() ? random_engine->e : nullptr, wsf2, hook2);
        reg();
      },
           py::arg("state"), py::arg("sampling_functions"), py::arg("json_sampling_functions"),
           py::arg("completion_check_params"), py::arg("event_generator"), py::arg("sample_period") = 1,
           py::arg("method_log") = py::none(), py::arg("random_engine") = py::none(),
           py::arg("write_status_f") = py::none(), py::arg("update_mode") = py::none());

  m.def("default_write_status", &default_write_status, py::arg("mc_calculator"), py::arg("method_log"));

  // ------------------------------------------------- sampling schedules
  // python/src/monte_sampling.cpp:414-742
  py::enum_<SAMPLE_MODE>(m, "SAMPLE_MODE")
      .value("BY_PASS", SAMPLE_MODE::BY_PASS)
      .value("BY_STEP", SAMPLE_MODE::BY_STEP)
      .value("BY_TIME", SAMPLE_MODE::BY_TIME)
      .export_values();
  py::enum_<SAMPLE_METHOD>(m, "SAMPLE_METHOD")
      .value("LINEAR", SAMPLE_METHOD::LINEAR)
      .value("LOG", SAMPLE_METHOD::LOG)
      .value("CUSTOM", SAMPLE_METHOD::CUSTOM)
      .export_values();
  py::class_<SamplingParams>(m, "SamplingParams")
      .def(py::init([](std::vector<std::string> sampler_names, std::vector<std::string> json_sampler_names,
                       SAMPLE_MODE sample_mode, SAMPLE_METHOD sample_method, double period,
                       std::optional<double> begin, double base, double shift, py::object custom_sample_at,
                       bool stochastic_sample_period, bool do_sample_trajectory, bool do_sample_time) {
             // make_sampling_params, python/src/monte_sampling.cpp:49-88
             if (!begin.has_value()) begin = (sample_method == SAMPLE_METHOD::LINEAR) ? period : 0.0;
             if (sample_method == SAMPLE_METHOD::CUSTOM && custom_sample_at.is_none())
               throw std::runtime_error(
                   "Error in make_sampling_params: sample_method==SAMPLE_METHOD::CUSTOM and "
                   "!custom_sample_at.has_value()");
             SamplingParams s;
             s.sampler_names = sampler_names;
             s.json_sampler_names = json_sampler_names;
             s.sample_mode = sample_mode;
             s.sample_method = sample_method;
             s.period = period;
             s.begin = begin.value();
             s.base = base;
             s.shift = shift;
             if (!custom_sample_at.is_none())
               s.custom_sample_at = [custom_sample_at](CountType n) {
                 py::gil_scoped_acquire gil;
                 return custom_sample_at(n).cast<double>();
               };
             s.stochastic_sample_period = stochastic_sample_period;
             s.do_sample_trajectory = do_sample_trajectory;
             s.do_sample_time = do_sample_time;
             return s;
           }),
           py::arg("sampler_names") = std::vector<std::string>(),
           py::arg("json_sampler_names") = std::vector<std::string>(),
           py::arg("sample_mode") = SAMPLE_MODE::BY_PASS, py::arg("sample_method") = SAMPLE_METHOD::LINEAR,
           py::arg("period") = 1.0, py::arg("begin") = py::none(), py::arg("base") = std::pow(10.0, 1.0 / 10.0),
           py::arg("shift") = 0.0, py::arg("custom_sample_at") = py::none(),
           py::arg("stochastic_sample_period") = false, py::arg("do_sample_trajectory") = false,
           py::arg("do_sample_time") = false)
      .def_readwrite("sample_mode", &SamplingParams::sample_mode)
      .def_readwrite("sample_method", &SamplingParams::sample_method)
      .def_readwrite("begin", &SamplingParams::begin)
      .def_readwrite("period", &SamplingParams::period)
      .def_readwrite("base", &SamplingParams::base)
      .def_readwrite("shift", &SamplingParams::shift)
      .def_readwrite("sampler_names", &SamplingParams::sampler_names)
      .def_readwrite("json_sampler_names", &SamplingParams::json_sampler_names)
      .def_readwrite("stochastic_sample_period", &SamplingParams::stochastic_sample_period)
      .def_readwrite("do_sample_trajectory", &SamplingParams::do_sample_trajectory)
      .def_readwrite("do_sample_time", &SamplingParams::do_sample_time)
      .def("append_to_sampler_names", [](SamplingParams &s, std::string name) { s.sampler_names.push_back(name); },
           py::arg("name"))
      .def("remove_from_sampler_names",
           [](SamplingParams &s, std::string name) {
             auto it = std::find(s.sampler_names.begin(), s.sampler_names.end(), name);
             if (it != s.sampler_names.end()) s.sampler_names.erase(it);
           },
           py::arg("name"))
      .def("extend_sampler_names",
           [](SamplingParams &s, std::vector<std::string> names) {
             s.sampler_names.insert(s.sampler_names.end(), names.begin(), names.end());
           },
           py::arg("names"))
      .def("append_to_json_sampler_names",
           [](SamplingParams &s, std::string name) { s.json_sampler_names.push_back(name); }, py::arg("name"))
      .def("remove_from_json_sampler_names",
           [](SamplingParams &s, std::string name) {
             auto it = std::find(s.json_sampler_names.begin(), s.json_sampler_names.end(), name);
             if (it != s.json_sampler_names.end()) s.json_sampler_names.erase(it);
           },
           py::arg("name"))
      .def("extend_json_sampler_names",
           [](SamplingParams &s, std::vector<std::string> names) {
             s.json_sampler_names.insert(s.json_sampler_names.end(), names.begin(), names.end());
           },
           py::arg("names"))
      .def("sample_at", [](SamplingParams const &s, CountType i) { return sample_at(i, s); },
           py::arg("sample_index"));
  m.def("sample_at", [](CountType i, SamplingParams const &s) { return sample_at(i, s); },
        py::arg("sample_index"), py::arg("sampling_params"));
  m.def("stochastic_count_step",
        [](double rate, RandomNumberGenerator<> &rng) { return stochastic_count_step(rate, rng); },
        py::arg("sample_rate"), py::arg("random_number_generator"));
  m.def("stochastic_time_step",
        [](double rate, RandomNumberGenerator<> &rng) { return stochastic_time_step(rate, rng); },
        py::arg("sample_rate"), py::arg("random_number_generator"));

  // ------------------------------------------------- run management
  typedef RunManager<default_engine_type> run_manager_type;
  typedef SamplingFixture<default_engine_type> fixture_type;
  py::class_<MonteCounter>(m, "MonteCounter")
      .def(py::init<>())
      .def_readonly("sample_mode", &MonteCounter::sample_mode)
      .def_readonly("steps_per_pass", &MonteCounter::steps_per_pass)
      .def_readonly("step", &MonteCounter::step)
      .def_readonly("pass_", &MonteCounter::pass)
      .def_readonly("count", &MonteCounter::count)
      .def_readonly("time", &MonteCounter::time)
      .def_readonly("n_accept", &MonteCounter::n_accept)
      .def_readonly("n_reject", &MonteCounter::n_reject)
      .def("reset", &MonteCounter::reset, py::arg("sample_mode"), py::arg("steps_per_pass"))
      .def("increment_step", &MonteCounter::increment_step)
      .def("increment_n_accept", &MonteCounter::increment_n_accept)
      .def("increment_n_reject", &MonteCounter::increment_n_reject)
      .def("advance_passes", &MonteCounter::advance_passes, py::arg("n_passes"), py::arg("d_accept"),
           py::arg("d_reject"));

  py::class_<Results>(m, "Results")
      .def_readonly("sampler_names", &Results::sampler_names)
      .def_readonly("json_sampler_names", &Results::json_sampler_names)
      .def_readonly("sampling_functions", &Results::sampling_functions)
      .def_readonly("analysis_functions", &Results::analysis_functions)
      .def_readonly("samplers", &Results::samplers)
      .def_property_readonly("json_samplers",
                             [](Results const &r) {
                               py::object loads = py::module_::import("json").attr("loads");
                               py::dict d;
                               for (auto const &kv : r.json_samplers) {
                                 py::list l;
                                 for (auto const &text : kv.second->values) l.append(loads(text));
                                 d[py::str(kv.first)] = l;
                               }
                               return d;
                             })
      .def_property_readonly("analysis",
                             [](Results const &r) {
                               py::dict d;
                               for (auto const &kv : r.analysis) d[py::str(kv.first)] = as_array(kv.second);
                               return d;
                             })
      .def_readonly("sample_count", &Results::sample_count)
      .def_readonly("sample_time", &Results::sample_time)
      .def_readonly("sample_weight", &Results::sample_weight)
      .def_readonly("sample_clocktime", &Results::sample_clocktime)
      .def_property_readonly("sample_trajectory",
                             [](Results const &r) {
                               py::list l;
                               for (auto const &occ : r.sample_trajectory) {
                                 py::array_t<int> a(occ.size());
                                 std::copy(occ.begin(), occ.end(), a.mutable_data());
                                 l.append(a);
                               }
                               return l;
                             })
      .def_readonly("completion_check_results", &Results::completion_check_results)
      .def_readonly("n_accept", &Results::n_accept)
      .def_readonly("n_reject", &Results::n_reject)
      .def_readonly("elapsed_clocktime", &Results::elapsed_clocktime)
      .def_readonly("initial_memory_used_MiB", &Results::initial_memory_used_MiB)
      .def_readonly("final_memory_used_MiB", &Results::final_memory_used_MiB)
      .def("is_auto_converge_mode", [](Results const &r) { return is_auto_converge_mode(r); })
      .def("N_samples", [](Results const &r) { return N_samples(r); })
      .def("N_samples_for_statistics", [](Results const &r) { return N_samples_for_statistics(r); })
      .def("N_samples_for_all_to_equilibrate", [](Results const &r) { return N_samples_for_all_to_equilibrate(r); })
      .def("all_equilibrated", [](Results const &r) { return all_equilibrated(r); })
      .def("all_converged", [](Results const &r) { return all_converged(r); })
      .def("acceptance_rate", [](Results const &r) { return acceptance_rate(r); })
      // Results.hh:217-330: per-component statistics as written to summary.json
      .def("quantity_stats", [](Results const &r, std::string const &name) {
        QuantityStats q(name, *r.samplers.at(name), r);
        py::dict d;
        d["shape"] = q.shape;
        d["is_scalar"] = q.is_scalar;
        d["component_names"] = q.component_names;
        py::list conv, stats;
        for (auto const &c : q.is_converged) conv.append(c.has_value() ? py::object(py::bool_(*c)) : py::object(py::none()));
        for (auto const &st : q.component_stats) {
          if (st.has_value()) {
            py::dict sd;
            sd["mean"] = st->mean;
            sd["calculated_precision"] = st->calculated_precision;
            stats.append(sd);
          } else {
            stats.append(py::none());
          }
        }
        d["is_converged"] = conv;
        d["component_stats"] = stats;
        return d;
      }, py::arg("quantity_name"));

  py::class_<ResultsAnalysisFunction>(m, "ResultsAnalysisFunction")
      .def(py::init([](std::string name, std::string description, std::vector<Index> shape, py::object function,
                       std::optional<std::vector<std::string>> component_names) {
             auto f = [function](Results const &results) -> std::vector<double> {
               py::gil_scoped_acquire gil;
               return to_dvec(function(py::cast(&results, py::return_value_policy::reference)));
             };
             return ResultsAnalysisFunction(name, description, shape, f, component_names);
           }),
           py::arg("name"), py::arg("description"), py::arg("shape"), py::arg("function"),
           py::arg("component_names") = py::none())
      .def_readwrite("name", &ResultsAnalysisFunction::name)
      .def_readwrite("description", &ResultsAnalysisFunction::description)
      .def_readwrite("shape", &ResultsAnalysisFunction::shape)
      .def_readwrite("component_names", &ResultsAnalysisFunction::component_names)
      .def("__call__", [](ResultsAnalysisFunction const &f, Results const &r) { return as_array(f(r)); });
  py::bind_map<ResultsAnalysisFunctionMap>(m, "ResultsAnalysisFunctionMap");
  m.def("make_heat_capacity_f", &make_heat_capacity_f, py::arg("mc_calculator"));
  m.def("make_susceptibility_f", &make_susceptibility_f, py::arg("mc_calculator"));

  py::class_<SamplingFixtureParams>(m, "SamplingFixtureParams")
      .def(py::init([](std::string label, StateSamplingFunctionMap const &sampling_functions,
                       PyJsonStateSamplingFunctionMap const &json_sampling_functions,
                       ResultsAnalysisFunctionMap const &analysis_functions, SamplingParams const &sampling_params,
                       CompletionCheckParams const &completion_check_params, std::vector<std::string> analysis_names,
                       py::object results_io, std::optional<MethodLog> method_log) {
             // JSON-valued functions return Python objects; the fixture stores JSON text
             casm_monte_b200::jsonStateSamplingFunctionMap jfs;
             for (auto const &kv : json_sampling_functions) {
               py::object pyf = kv.second.function;
               jfs.emplace(kv.first, casm_monte_b200::jsonStateSamplingFunction{
                                         kv.second.name, kv.second.description, [pyf]() -> std::string {
                                           py::gil_scoped_acquire gil;
                                           return py::module_::import("json").attr("dumps")(pyf()).cast<std::string>();
                                         }});
             }
             ResultsIOFunction io;
             if (!results_io.is_none())
               io = [results_io](Results const &results, ValueMap const &conditions, Index run_index) {
                 py::gil_scoped_acquire gil;
                 results_io.attr("write")(py::cast(&results, py::return_value_policy::reference),
                                          py::cast(conditions), run_index);
               };
             return SamplingFixtureParams(label, sampling_functions, jfs, analysis_functions, sampling_params,
                                          completion_check_params, analysis_names, io,
                                          method_log.has_value() ? *method_log : MethodLog());
           }),
           py::arg("label"), py::arg("sampling_functions"), py::arg("json_sampling_functions"),
           py::arg("analysis_functions"), py::arg("sampling_params"), py::arg("completion_check_params"),
           py::arg("analysis_names") = std::vector<std::string>(), py::arg("results_io") = py::none(),
           py::arg("method_log") = py::none())
      .def_readonly("label", &SamplingFixtureParams::label)
      .def_readonly("sampling_functions", &SamplingFixtureParams::sampling_functions)
      .def_readonly("analysis_functions", &SamplingFixtureParams::analysis_functions)
      .def_readonly("sampling_params", &SamplingFixtureParams::sampling_params)
      .def_readonly("completion_check_params", &SamplingFixtureParams::completion_check_params)
      .def_readonly("analysis_names", &SamplingFixtureParams::analysis_names)
      .def_readonly("method_log", &SamplingFixtureParams::method_log);

  py::class_<fixture_type, std::shared_ptr<fixture_type>>(m, "SamplingFixture")
      .def(py::init([](SamplingFixtureParams const &params, PyEngine const &engine) {
             return std::make_shared<fixture_type>(params, engine.e);
           }),
           py::arg("params"), py::arg("engine"))
      .def("label", &fixture_type::label)
      .def("params", &fixture_type::params, py::return_value_policy::reference_internal)
      .def("counter", &fixture_type::counter, py::return_value_policy::reference_internal)
      .def("results", &fixture_type::results, py::return_value_policy::reference_internal)
      .def("completion_check_results", &fixture_type::completion_check_results,
           py::return_value_policy::reference_internal)
      .def("next_sample_count", &fixture_type::next_sample_count)
      .def("initialize", &fixture_type::initialize, py::arg("steps_per_pass"))
      .def("is_complete", &fixture_type::is_complete)
      .def("increment_step", &fixture_type::increment_step)
      .def("increment_n_accept", &fixture_type::increment_n_accept)
      .def("increment_n_reject", &fixture_type::increment_n_reject)
      .def("advance_passes", &fixture_type::advance_passes, py::arg("n_passes"), py::arg("d_accept"),
           py::arg("d_reject"))
      .def("sample_data", &fixture_type::sample_data, py::arg("state"))
      .def("sample_data_by_count_if_due", &fixture_type::sample_data_by_count_if_due, py::arg("state"))
      .def("steps_to_next_event", &fixture_type::steps_to_next_event)
      .def("write_status", &fixture_type::write_status, py::arg("run_index"))
      .def("finalize", &fixture_type::finalize, py::arg("state"), py::arg("run_index"));

  py::class_<run_manager_type, std::shared_ptr<run_manager_type>>(m, "RunManager")
      .def(py::init([](PyEngine const &engine, std::vector<SamplingFixtureParams> const &params, bool global_cutoff) {
             return std::make_shared<run_manager_type>(engine.e, params, global_cutoff);
           }),
           py::arg("engine"), py::arg("sampling_fixture_params"), py::arg("global_cutoff") = true)
      .def_readwrite("run_index", &run_manager_type::run_index)
      .def_readwrite("global_cutoff", &run_manager_type::global_cutoff)
      .def_property_readonly("engine", [](run_manager_type const &r) { return PyEngine{r.engine}; })
      .def_readonly("sampling_fixtures", &run_manager_type::sampling_fixtures)
      .def("initialize", &run_manager_type::initialize, py::arg("steps_per_pass"))
      .def("is_complete", &run_manager_type::is_complete)
      .def("write_status_if_due", &run_manager_type::write_status_if_due)
      .def("increment_step", &run_manager_type::increment_step)
      .def("increment_n_accept", &run_manager_type::increment_n_accept)
      .def("increment_n_reject", &run_manager_type::increment_n_reject)
      .def("advance_passes", &run_manager_type::advance_passes, py::arg("n_passes"), py::arg("d_accept"),
           py::arg("d_reject"))
      .def("sample_data_by_count_if_due", &run_manager_type::sample_data_by_count_if_due, py::arg("state"))
      .def("steps_to_next_event", &run_manager_type::steps_to_next_event)
      .def("finalize", &run_manager_type::finalize, py::arg("final_state"));

  m.def("occupation_metropolis",
        [](std::shared_ptr<calculator_type> mc_calculator, IsingState &state,
           std::shared_ptr<run_manager_type> run_manager, std::string update_mode) {
          if (!mc_calculator || !run_manager)
            throw std::runtime_error("Error in occupation_metropolis: mc_calculator / run_manager is None");
          occupation_metropolis(*mc_calculator, state, *run_manager, update_mode);
        },
        py::arg("mc_calculator"), py::arg("state"), py::arg("run_manager"), py::arg("update_mode") = "auto");
}
