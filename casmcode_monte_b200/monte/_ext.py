"""Loads the pybind11 host module (built in-tree next to the package)."""
try:
    from .. import _monte_b200 as ext
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "casmcode_monte_b200._monte_b200 is not built: run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (g++ + nvcc). There is no CPU fallback."
    ) from e
