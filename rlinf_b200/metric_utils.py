"""Rollout metrics and loss mask with the reference's names (rlinf/utils/metric_utils.py)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def compute_loss_mask(dones):
    """rlinf/utils/metric_utils.py:516-537."""
    return ops.loss_mask(dones)


def compute_rollout_metrics(data_buffer: dict, world_size: int = 1, group=None) -> dict:
    """rlinf/utils/metric_utils.py:422-506: mean rewards; mean/min/max of advantages and returns over the valid
    (loss-masked) entries, reduced over ranks (SUM of sum/count, MAX of -min/max). One kernel per tensor, one
    packed all-reduce pair, ONE device->host copy (the reference does 2 all-reduces + 4 .item() per tensor)."""
    loss_mask = data_buffer.get("loss_mask", None)
    names = [k for k in ("rewards", "advantages", "returns") if data_buffer.get(k, None) is not None]
    if not names:
        return {}
    stats = []
    for k in names:
        x = data_buffer[k]
        div = 1
        m = loss_mask
        if m is not None:
            if m.shape != x.shape:
                if m.numel() * x.shape[-1] == x.numel():
                    div = x.shape[-1]
                else:
                    m = torch.broadcast_to(m, x.shape).contiguous()
            m = m.contiguous()
        stats.append(ops.masked_stats(x.contiguous(), m, div))
    st = torch.stack(stats)  # [k, 4] = count, sum, min, max
    if world_size > 1:
        sc = st[:, :2].contiguous()
        mm = torch.stack([-st[:, 2], st[:, 3]], dim=1).contiguous()
        dist.all_reduce(sc, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
        st = torch.cat([sc, -mm[:, :1], mm[:, 1:]], dim=1)
    host = st.tolist()
    out = {}
    for k, (cnt, total, mn, mx) in zip(names, host):
        mean = total / cnt if cnt > 0 else float("nan")
        if cnt <= 0:
            mn = mx = float("nan")
        if k == "rewards":
            out["rewards"] = mean
        else:
            out[f"{k}_mean"], out[f"{k}_max"], out[f"{k}_min"] = mean, mx, mn
    return out
