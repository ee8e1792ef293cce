"""Learning-rate schedules of the actor (host logic: one multiplier per `run_training` call).

Mirror of `FSDPModelManager.build_lr_scheduler` (rlinf/hybrid_engines/fsdp/fsdp_model_manager.py:465-499) and
`get_lr_scheduler` (rlinf/hybrid_engines/fsdp/utils.py:522-604): a `LambdaLR` whose multiplier is applied to every
param group's base lr, stepped ONCE per `run_training` (workers/actor/embodied_fsdp_actor_worker.py:571).
`LambdaLR.__init__` already performs one step, so the multiplier in force during the k-th `run_training` call
(k = 0, 1, ...) is `lr_lambda(k)`.  The multiplier is a Python double, as in torch; the optimiser kernel reads
`base_lr * multiplier` from a device-resident table (rb200_adamw_step_dev), so CUDA-graph replays follow it.
"""
from __future__ import annotations

import math


class LRSchedule:
    def __init__(self, optim_cfg=None, base_lr: float = 1.0):
        g = (optim_cfg or {}).get
        self.kind = g("lr_scheduler", "constant")
        self.total_steps = g("total_training_steps", 0)
        warm = int(g("lr_warmup_steps", -1))
        if warm < 0:
            warm = int(g("lr_warmup_steps_ratio", 0.0) * self.total_steps)
        self.num_warmup_steps = warm
        self.num_cycles = g("num_cycles", 0.5)
        min_lr, min_lr_rate = g("min_lr", 0.0), g("min_lr_rate", None)
        if min_lr_rate is not None:  # utils.py:532-534: min_lr_rate wins
            min_lr = None
        self.min_lr, self.min_lr_rate = min_lr, min_lr_rate
        self.base_lr = float(base_lr)
        self.last_epoch = 0
        if self.kind not in ("constant", "cosine", "openpi_cosine", "ref_warmup_cosine", "torch_constant",
                             "torch_cosine"):
            raise NotImplementedError(f"Scheduler type {self.kind} is not supported")
        if self.kind == "cosine" and self.min_lr is None and self.min_lr_rate is None:
            raise ValueError("One of min_lr or min_lr_rate should be set through the `lr_scheduler_kwargs`")

    def multiplier(self, step: int | None = None) -> float:
        s = self.last_epoch if step is None else int(step)
        w, n = self.num_warmup_steps, self.total_steps
        if self.kind == "constant":
            return float(s) / float(max(1.0, w)) if s < w else 1.0
        if self.kind == "cosine":  # transformers.get_cosine_with_min_lr_schedule_with_warmup
            rate = self.min_lr_rate if self.min_lr_rate is not None else self.min_lr / self.base_lr
            if s < w:
                return float(s) / float(max(1, w))
            progress = float(s - w) / float(max(1, n - w))
            factor = 0.5 * (1.0 + math.cos(math.pi * float(self.num_cycles) * 2.0 * progress))
            return max(0, factor * (1 - rate) + rate)
        if self.kind in ("openpi_cosine", "ref_warmup_cosine"):
            if self.min_lr_rate is not None:
                min_mult = self.min_lr_rate
            elif self.min_lr and self.base_lr > 0:
                min_mult = self.min_lr / self.base_lr
            else:
                min_mult = 0.0
            if s < w:
                init = 1.0 / (w + 1)
                return init + (1.0 - init) * s / max(1, w)
            progress = min(1.0, (s - w) / max(1, n - w))
            return min_mult + (1.0 - min_mult) * 0.5 * (1.0 + math.cos(math.pi * progress))
        if self.kind == "torch_constant":
            return 1.0
        # torch_cosine: CosineAnnealingLR(T_max=total, eta_min=1e-6), closed form
        eta_min = 1e-6 / self.base_lr if self.base_lr > 0 else 0.0
        return eta_min + (1.0 - eta_min) * 0.5 * (1.0 + math.cos(math.pi * s / max(1, n)))

    def step(self) -> float:
        self.last_epoch += 1
        return self.multiplier()
