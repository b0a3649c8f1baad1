"""JSON results writer with the file layout of the reference's jsonResultsIO
(include/casm/monte/run_management/io/json/jsonResultsIO_impl.hh:247-365):

  <output_dir>/summary.json             one entry appended per completed run
  <output_dir>/run.<index>/observations.json   if write_observations
  <output_dir>/run.<index>/trajectory.json     if write_trajectory

summary.json holds, under "conditions", "statistics", "completion_check_results"
and "analysis", arrays with one element per run (:33-246).
"""
import json
import os


def _ensure(obj, key, kind):
    if key not in obj:
        obj[key] = kind()
    elif not isinstance(obj[key], kind):
        what = "an object" if kind is dict else "an array"
        raise RuntimeError(f'JSON Error: "{key}" is expected to be {what}.')
    return obj[key]


class jsonResultsIO:
    def __init__(self, output_dir, write_trajectory=False, write_observations=False):
        self.output_dir = str(output_dir)
        self.write_trajectory = bool(write_trajectory)
        self.write_observations = bool(write_observations)

    # jsonResultsIO_impl.hh:262-272
    def write(self, results, conditions, run_index):
        self.write_summary(results, conditions)
        if self.write_trajectory:
            self._write_trajectory(results, run_index)
        if self.write_observations:
            self._write_observations(results, run_index)

    # :274-283
    def to_dict(self):
        return {
            "method": "json",
            "kwargs": {
                "output_dir": self.output_dir,
                "write_trajectory": self.write_trajectory,
                "write_observations": self.write_observations,
            },
        }

    def run_dir(self, run_index):
        path = os.path.join(self.output_dir, f"run.{run_index}")
        os.makedirs(path, exist_ok=True)
        return path

    # :354-365
    def read_summary(self):
        path = os.path.join(self.output_dir, "summary.json")
        if not os.path.exists(path):
            return {"conditions": {}, "statistics": {}, "completion_check_results": {}}
        with open(path) as f:
            return json.load(f)

    # :33-66, scalar / vector conditions (component names default to "0", "1", ...)
    @staticmethod
    def _append_condition(name, value, shape, component_names, out):
        j = _ensure(out, name, dict)
        j["shape"] = list(shape)
        if len(shape) == 0:
            _ensure(j, "value", list).append(float(value[0]))
        else:
            j["component_names"] = list(component_names)
            for i, cname in enumerate(component_names):
                _ensure(j, cname, list).append(float(value[i]))

    # :285-330
    def write_summary(self, results, conditions):
        js = self.read_summary()
        for key in ("conditions", "statistics", "completion_check_results", "analysis"):
            _ensure(js, key, dict)
        cond = conditions.to_dict() if hasattr(conditions, "to_dict") else dict(conditions)
        for name, value in cond.items():
            if isinstance(value, bool):
                continue
            if isinstance(value, (int, float)):
                self._append_condition(name, [value], [], ["0"], js["conditions"])
            else:
                flat = [x for row in value for x in row] if value and isinstance(value[0], (list, tuple)) else list(value)
                if value and isinstance(value[0], (list, tuple)):
                    shape = [len(value), len(value[0])]
                    names = [f"{r},{c}" for c in range(shape[1]) for r in range(shape[0])]
                    flat = [value[r][c] for c in range(shape[1]) for r in range(shape[0])]  # column-major unroll
                else:
                    shape = [len(flat)]
                    names = [str(i) for i in range(len(flat))]
                fn = results.sampling_functions
                if name in fn and list(fn[name].shape) == shape:
                    names = list(fn[name].component_names)
                self._append_condition(name, flat, shape, names, js["conditions"])

        # statistics, :102-144 + BasicStatistics.cc:190-208
        for name in results.samplers:
            q = results.quantity_stats(name)
            qj = _ensure(js["statistics"], name, dict)
            qj["shape"] = list(q["shape"])

            def append(tj, i):
                st = q["component_stats"][i]
                _ensure(tj, "mean", list).append(None if st is None else st["mean"])
                _ensure(tj, "calculated_precision", list).append(None if st is None else st["calculated_precision"])
                if q["is_converged"][i] is not None:
                    _ensure(tj, "is_converged", list).append(bool(q["is_converged"][i]))

            if q["is_scalar"]:
                append(_ensure(qj, "value", dict), 0)
            else:
                qj["component_names"] = list(q["component_names"])
                for i, cname in enumerate(q["component_names"]):
                    append(_ensure(qj, cname, dict), i)

        # completion check results, :146-181
        cj = js["completion_check_results"]
        if results.is_auto_converge_mode():
            _ensure(cj, "all_equilibrated", list).append(bool(results.all_equilibrated()))
            _ensure(cj, "all_converged", list).append(bool(results.all_converged()))
            _ensure(cj, "N_samples_for_all_to_equilibrate", list).append(
                int(results.N_samples_for_all_to_equilibrate()) if results.all_equilibrated() else "did_not_equilibrate"
            )
        _ensure(cj, "N_samples", list).append(int(results.N_samples()))
        _ensure(cj, "N_samples_for_statistics", list).append(int(results.N_samples_for_statistics()))
        _ensure(cj, "acceptance_rate", list).append(float(results.acceptance_rate()))
        _ensure(cj, "initial_memory_used_MiB", list).append(results.initial_memory_used_MiB)
        _ensure(cj, "final_memory_used_MiB", list).append(results.final_memory_used_MiB)
        _ensure(cj, "elapsed_clocktime", list).append(results.elapsed_clocktime)
        _ensure(cj, "count", list).append(int(results.sample_count[-1]))
        if len(results.sample_time):
            _ensure(cj, "time", list).append(float(results.sample_time[-1]))

        # analysis, :183-244
        not_equilibrated = results.is_auto_converge_mode() and not results.all_equilibrated()
        for name, value in results.analysis.items():
            if name not in results.analysis_functions:
                raise RuntimeError(
                    f"Error in append_results_analysis_to_json: No matching analysis function found for '{name}'."
                )
            f = results.analysis_functions[name]
            vj = _ensure(js["analysis"], name, dict)
            vj["shape"] = list(f.shape)
            if len(f.shape) == 0:
                _ensure(vj, "value", list).append("did_not_equilibrate" if not_equilibrated else float(value[0]))
            else:
                vj["component_names"] = list(f.component_names)
                for i, cname in enumerate(f.component_names):
                    _ensure(vj, cname, list).append("did_not_equilibrate" if not_equilibrated else float(value[i]))

        os.makedirs(self.output_dir, exist_ok=True)
        tmp = os.path.join(self.output_dir, "summary.json.tmp")
        with open(tmp, "w") as f:
            json.dump(js, f)
        os.replace(tmp, os.path.join(self.output_dir, "summary.json"))  # SafeOfstream: write then move

    # :332-336: the sampled configurations (occupation vectors here)
    def _write_trajectory(self, results, run_index):
        data = [[int(x) for x in occ] for occ in results.sample_trajectory]
        with open(os.path.join(self.run_dir(run_index), "trajectory.json"), "w") as f:
            json.dump(data, f)

    # :338-352... observations.json
    def _write_observations(self, results, run_index):
        js = {}
        if len(results.sample_count):
            js["count"] = [int(x) for x in results.sample_count]
        if len(results.sample_time):
            js["time"] = [float(x) for x in results.sample_time]
        if results.sample_weight.n_samples():
            js["weight"] = [float(x) for x in results.sample_weight.component(0)]
        if len(results.sample_clocktime):
            js["clocktime"] = [float(x) for x in results.sample_clocktime]
        for name, sampler in results.samplers.items():
            entry = {"shape": list(sampler.shape())}
            if len(sampler.shape()) == 0:
                entry["value"] = [float(x) for x in sampler.component(0)]
            else:
                entry["component_names"] = list(sampler.component_names())
                entry["value"] = [[float(x) for x in row] for row in sampler.values()]
            js[name] = entry
        for name, values in results.json_samplers.items():
            js[name] = {"value": list(values)}
        with open(os.path.join(self.run_dir(run_index), "observations.json"), "w") as f:
            json.dump(js, f)
