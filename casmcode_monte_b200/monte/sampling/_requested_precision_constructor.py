"""``converge(...).set_precision(...)`` helper, same call surface as
python/libcasm/monte/sampling/_requested_precision_constructor.py:6-131."""
from typing import Optional

from .._ext import ext as _ext


class RequestedPrecisionConstructor:
    """Fills ``completion_check_params.requested_precision`` one quantity at a time."""

    def __init__(self, sampling_functions, completion_check_params):
        self.sampling_functions = sampling_functions
        self.completion_check_params = completion_check_params

    def set_precision(
        self,
        quantity: str,
        abs: Optional[float] = None,
        rel: Optional[float] = None,
        component_name: Optional[list] = None,
        component_index: Optional[list] = None,
    ):
        if quantity not in self.sampling_functions:
            raise Exception(f"{quantity} is not in sampling_functions")
        if rel is None and abs is None:
            raise Exception("No abs or rel precision specified")
        names = list(self.sampling_functions[quantity].component_names)
        wanted = set(component_index or [])
        for n in component_name or []:
            if n not in names:
                raise Exception(f"{n} is not a component of {quantity}")
            wanted.add(names.index(n))
        if not wanted:
            wanted = set(range(len(names)))
        for i in sorted(wanted):
            key = _ext.SamplerComponent(sampler_name=quantity, component_index=i, component_name=names[i])
            self.completion_check_params.requested_precision[key] = _ext.RequestedPrecision(abs=abs, rel=rel)
        return self


def converge(sampling_functions, completion_check_params):
    return RequestedPrecisionConstructor(sampling_functions, completion_check_params)
