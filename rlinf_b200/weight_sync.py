"""Actor -> rollout parameter hand-off.

Two transports over the SAME flat fp32 parameter buffer:

* `broadcast_flat` - what this framework's own ranks use: ONE NCCL broadcast of `[flat_params | version]`
  (north star: "an NCCL broadcast of updated params back to rollout workers").
* `BucketWeightSync` - the wire format of the reference's `BucketWeightSyncer`
  (rlinf/hybrid_engines/weight_syncer/bucket_syncer.py:113-373), so a stock RLinf rollout worker can receive from this
  actor (or this rollout replica from a stock actor): a stream of `dict[str, Tensor]` buckets keyed by the
  reference's parameter names, cut by the reference's rule (`iter_named_tensor_buckets` :33-110: append tensors until
  the running byte count reaches `bucket_size`, a tensor is never split), the first bucket additionally carrying
  `total_buckets` and `syncer_version` as int32 scalars.  Payload tensors are VIEWS of the flat buffer (no staging copy)
  unless a transport dtype is requested.  `apply` consumes such buckets - from either implementation - with
  `load_state_dict(strict=False)` semantics (unknown keys ignored) and returns the version.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional

import torch
import torch.distributed as dist

TOTAL_BUCKETS_KEY = "total_buckets"      # bucket_syncer.py:142
SYNCER_VERSION_KEY = "syncer_version"    # bucket_syncer.py:143


def plan_buckets(named_sizes: Iterable[tuple[str, int]], bucket_size: int) -> list[list[str]]:
    """The reference's cut rule on (name, nbytes) pairs: host logic, no tensors involved."""
    plan, cur, held = [], [], 0
    for name, nbytes in named_sizes:
        if name in (TOTAL_BUCKETS_KEY, SYNCER_VERSION_KEY):
            raise ValueError(f"Bucket payload key conflicts with metadata key: {name}")
        cur.append(name)
        held += nbytes
        if held >= bucket_size:
            plan.append(cur)
            cur, held = [], 0
    if held > 0:
        plan.append(cur)
    if not plan:
        raise ValueError("No parameters to sync")
    return plan


class BucketWeightSync:
    def __init__(self, policy, bucket_size: int = 128 * 1024 * 1024, bucket_dtype: Optional[torch.dtype] = None,
                 param_names_need_sync: Optional[list[str]] = None):
        self.policy = policy
        self.bucket_size = int(bucket_size)
        self.bucket_dtype = bucket_dtype
        names = [n for n, _ in policy.named_parameters()]
        self.param_names_need_sync = list(param_names_need_sync) if param_names_need_sync is not None else names
        if not self.param_names_need_sync:
            raise ValueError("param_names_need_sync must not be empty")

    def _transport_dtype(self, dtype: torch.dtype) -> torch.dtype:
        return self.bucket_dtype if (self.bucket_dtype is not None and dtype.is_floating_point) else dtype

    def iter_buckets(self, version: int) -> Iterator[dict]:
        views = dict(self.policy.named_parameters())
        items = [(k, views[k]) for k in self.param_names_need_sync if k in views and "_extra_state" not in k]
        esize = torch.empty((), dtype=self._transport_dtype(torch.float32)).element_size()
        plan = plan_buckets(((k, v.numel() * esize) for k, v in items), self.bucket_size)
        dev = self.policy.flat_params.device
        bucket = {TOTAL_BUCKETS_KEY: torch.tensor(len(plan), dtype=torch.int32, device=dev),
                  SYNCER_VERSION_KEY: torch.as_tensor(version, dtype=torch.int32, device=dev)}
        for names in plan:
            for k in names:
                t = views[k]
                td = self._transport_dtype(t.dtype)
                bucket[k] = t if td == t.dtype else t.to(td)
            yield bucket
            bucket = {}

    def sync(self, send: Callable[[dict], None], version: int) -> None:
        for bucket in self.iter_buckets(version):
            send(bucket)

    def apply(self, recv: Callable[[], dict]) -> int:
        """Receive `total_buckets` buckets and copy every known tensor into the flat buffer; returns the version."""
        views = dict(self.policy.named_parameters())
        bucket = dict(recv())
        total = int(bucket.pop(TOTAL_BUCKETS_KEY).item())
        version = int(bucket.pop(SYNCER_VERSION_KEY).item())
        for i in range(total):
            if i > 0:
                bucket = recv()
            for k, v in bucket.items():
                dst = views.get(k)
                if dst is None:
                    continue  # strict=False
                if dst.shape != v.shape:
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(dst.shape)}")
                dst.copy_(v)  # casts a bf16 transport back to fp32, H2D if the bucket was staged on the host
        self.policy.mark_params_changed()
        return version


def broadcast_flat(policy, version: int, src: int = 0, group=None) -> int:
    """One NCCL broadcast of [flat parameters | version]; returns the version every rank now holds."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(version)
    buf = policy.sync_buffer()
    buf[-1] = float(version)  # exact for |version| < 2^24
    dist.broadcast(buf, src=src, group=group)
    policy.mark_params_changed()
    return int(buf[-1].item())
