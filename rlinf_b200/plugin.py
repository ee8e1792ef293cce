"""The binding a maintainer adds to RLinf: re-register the CUDA-backed callables in the reference's own registries.

    import rlinf.algorithms                      # fills ADV_REGISTRY / LOSS_REGISTRY / LOSS_SCALE_REGISTRY first
    import rlinf_b200.plugin as b200
    b200.install()                               # every entry of the three registries now launches librlinf_b200.so

`rlinf.algorithms.registry.calculate_adv_and_returns` / `policy_loss` (registry.py:77-124), their 7 worker call sites
 and rlinf/config.py stay untouched: the reference keeps doing its own pre/post-processing and hands
the registered callable the kwargs it always did (step-major [T,B] tensors for advantages, the kwargs of
preprocess_loss_inputs for losses); the callables copy host tensors to the current CUDA device (there is no CPU
path) and return CUDA tensors.
"""
from __future__ import annotations

from typing import Optional

from . import algorithms as _alg


def install(adv_registry: Optional[dict] = None, loss_registry: Optional[dict] = None,
            loss_scale_registry: Optional[dict] = None, names=None) -> dict:
    """Overwrite the given registries (default: the reference's, imported from rlinf.algorithms.registry) with the
    entries of rlinf_b200.algorithms. `names` restricts the set. Returns {registry name: [replaced keys]}."""
    if adv_registry is None or loss_registry is None:
        from rlinf.algorithms import registry as ref  # the reference must be importable for the default

        adv_registry = ref.ADV_REGISTRY if adv_registry is None else adv_registry
        loss_registry = ref.LOSS_REGISTRY if loss_registry is None else loss_registry
        if loss_scale_registry is None:
            loss_scale_registry = getattr(ref, "LOSS_SCALE_REGISTRY", None)
    done = {"adv": [], "loss": [], "loss_scale": []}
    for key, src, dst in (("adv", _alg.ADV_REGISTRY, adv_registry), ("loss", _alg.LOSS_REGISTRY, loss_registry),
                          ("loss_scale", _alg.LOSS_SCALE_REGISTRY, loss_scale_registry)):
        if dst is None:
            continue
        for name, fn in src.items():
            if names is not None and name not in names:
                continue
            dst[name] = fn
            done[key].append(name)
    return done
