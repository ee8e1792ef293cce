"""Helpers with the reference's `rlinf.utils.*` names that sit on the hot path."""
from .distributed import masked_normalization, masked_stats, normalize_from_stats  # noqa: F401
