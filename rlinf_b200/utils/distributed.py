"""Masked fp64 normalisations of rlinf/utils/distributed.py, CUDA-backed.

`masked_normalization` (:866-939), `masked_stats` (:942-954), `normalize_from_stats` (:957-965).  The reference issues
three all-reduces (factor, sum, sum of squares); here the three fp64 statistics come out of ONE reduction kernel into
one 24-byte buffer that is all-reduced ONCE, and one elementwise kernel applies the normalisation.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import ops


def masked_stats(x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{count, sum, sum of squares} of the valid values, float64[3] on the device (distributed.py:942-954)."""
    return ops.masked_moments(x, mask)


def normalize_from_stats(x: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    """(x - mean) * rsqrt(max(var, 0) + 1e-5) from (all-reduced) {count, sum, sumsq} (distributed.py:957-965)."""
    return ops.masked_normalize(x, stats, None, mode=1).view(x.shape)


@torch.no_grad()
def masked_normalization(x: torch.Tensor, mask: Optional[torch.Tensor] = None, dim=None, inplace: bool = False,
                         unbiased: bool = False, eps: float = 1e-5, high_precision: bool = True,
                         all_reduce: bool = True, group=None) -> torch.Tensor:
    """Advantage normalisation over the whole (data-parallel) batch (distributed.py:866-939).  Semantics kept: the
    input is multiplied by the mask BEFORE the statistics and the masked entries come out as (0 - mean)/(std + eps).
    `dim` other than None / all dimensions is not needed by any caller of the reference and is not implemented."""
    if dim is not None and tuple(sorted(d % x.dim() for d in (dim if isinstance(dim, (tuple, list)) else (dim,)))) != tuple(
            range(x.dim())):
        raise NotImplementedError("masked_normalization: only dim=None (all dimensions) is implemented")
    if not high_precision:
        raise NotImplementedError("masked_normalization: statistics are always accumulated in float64")
    if mask is not None and mask.shape != x.shape:
        raise AssertionError((tuple(mask.shape), tuple(x.shape), dim))
    stats = ops.masked_moments(x, mask)
    if mask is None:
        stats[0] = float(x.numel())
    if all_reduce and dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    out = ops.masked_normalize(x, stats, mask, mode=0, eps=eps, unbiased=unbiased).view(x.shape)
    if inplace and x.is_cuda and x.dtype == torch.float32:
        x.copy_(out)
        return x
    return out
