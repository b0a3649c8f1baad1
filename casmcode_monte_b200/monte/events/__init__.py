"""Mirror of the ``libcasm.monte.events`` names on this path."""
from .._ext import ext as _ext

_names = (
    "Conversions IntVector LongVector OccEvent OccTransform OccCandidate OccSwap OccCandidateList Mol OccLocation "
    "is_allowed_canonical_swap make_canonical_swaps is_allowed_semigrand_canonical_swap make_semigrand_canonical_swaps "
    "get_n_allowed_per_unitcell choose_canonical_swap propose_canonical_event_from_swap propose_canonical_event "
    "choose_semigrand_canonical_swap propose_semigrand_canonical_event_from_swap propose_semigrand_canonical_event"
).split()
for _n in _names:
    globals()[_n] = getattr(_ext, _n)

__all__ = list(_names)
