"""Import the UNMODIFIED reference (`/root/reference`, RLinf v0.4.0) in this container.

Only `tests/golden/make_golden.py` (run by hand, here) and the optional
`tests/test_oracle_vs_reference.py` use this.  `/root/reference` does not exist
on the GPU box, so nothing in the `-m gpu` tests, `smoke()` or `bench.py` may
import this module.  Nothing is copied out of the reference: it is executed in
place and only its numerical outputs are stored (as `.npz` fixtures).

The reference cannot be imported as-is because `rlinf/__init__.py` registers
OmegaConf resolvers (omegaconf is not installed) and `rlinf/utils/utils.py`
imports `rlinf.scheduler.Worker` (-> Ray, not installed).  Two tiny stand-in
modules placed in `sys.modules` before the import are enough for the
algorithm, data-schema and MLP-policy modules to load.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RLINF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rlinf", "algorithms"))


def _stub_modules() -> None:
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:  # noqa: D401 - stand-in
            @staticmethod
            def register_new_resolver(name, fn=None, replace=True, **kw):
                return None

            @staticmethod
            def register_resolver(name, fn=None, replace=True, **kw):
                return None

            @staticmethod
            def select(cfg, key, default=None):
                return default

        oc.OmegaConf = OmegaConf
        oc.DictConfig = dict
        oc.ListConfig = list
        oc.open_dict = None
        sys.modules["omegaconf"] = oc
        dc = types.ModuleType("omegaconf.dictconfig")
        dc.DictConfig = dict
        sys.modules["omegaconf.dictconfig"] = dc
        lc = types.ModuleType("omegaconf.listconfig")
        lc.ListConfig = list
        sys.modules["omegaconf.listconfig"] = lc

    if "rlinf.scheduler" not in sys.modules:
        sched = types.ModuleType("rlinf.scheduler")

        class _Platform:
            @staticmethod
            def current_device():
                return "cpu"

            @staticmethod
            def synchronize():
                return None

            @staticmethod
            def empty_cache():
                return None

            @staticmethod
            def ipc_collect():
                return None

            @staticmethod
            def is_available():
                return False

        class Worker:
            torch_platform = _Platform()
            torch_device_type = "cpu"

            @staticmethod
            def timer(name):
                def deco(fn):
                    return fn

                return deco

        sched.Worker = Worker
        sched.Channel = object
        sched.Cluster = object
        sched.__path__ = []  # behave as a package so `rlinf.scheduler.worker.worker` can be stubbed too
        sys.modules["rlinf.scheduler"] = sched
        wpkg = types.ModuleType("rlinf.scheduler.worker")
        wpkg.__path__ = []
        wmod = types.ModuleType("rlinf.scheduler.worker.worker")
        wmod.Worker = Worker
        wpkg.worker = wmod
        sys.modules["rlinf.scheduler.worker"] = wpkg
        sys.modules["rlinf.scheduler.worker.worker"] = wmod


def _namespace(name: str, rel: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
    sys.modules[name] = mod


_LOADED = None


def load_reference():
    """Return a namespace with the reference modules used on the hot path."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub_modules()
    # `rlinf.algorithms.__init__` imports rlinf.agents.tool_call.parsers (sglang etc.)
    # -> bypass that package __init__ too; we import the submodules directly.
    import rlinf  # noqa: F401  (its __init__ only registers resolvers)

    _namespace("rlinf.algorithms", "rlinf/algorithms")
    _namespace("rlinf.data", "rlinf/data")
    _namespace("rlinf.data.schema", "rlinf/data/schema")
    _namespace("rlinf.models", "rlinf/models")
    _namespace("rlinf.models.embodiment", "rlinf/models/embodiment")
    _namespace("rlinf.models.embodiment.modules", "rlinf/models/embodiment/modules")
    _namespace("rlinf.models.embodiment.mlp_policy", "rlinf/models/embodiment/mlp_policy")

    import importlib

    ns = types.SimpleNamespace()
    ns.registry = importlib.import_module("rlinf.algorithms.registry")
    ns.advantages = importlib.import_module("rlinf.algorithms.advantages")
    ns.losses = importlib.import_module("rlinf.algorithms.losses")
    ns.alg_utils = importlib.import_module("rlinf.algorithms.utils")
    ns.utils = importlib.import_module("rlinf.utils.utils")
    ns.metric_utils = importlib.import_module("rlinf.utils.metric_utils")
    ns.nested = importlib.import_module("rlinf.utils.nested_dict_process")
    try:
        ns.mlp_policy = importlib.import_module(
            "rlinf.models.embodiment.mlp_policy.mlp_policy"
        )
    except Exception as e:  # pragma: no cover - optional
        ns.mlp_policy = None
        ns.mlp_policy_error = repr(e)
    try:
        ns.embodied_types = importlib.import_module("rlinf.data.schema.embodied_types")
        ns.traj_builder = importlib.import_module(
            "rlinf.data.schema.embodied_trajectory_builder"
        )
    except Exception as e:  # pragma: no cover - optional
        ns.embodied_types = None
        ns.embodied_types_error = repr(e)
    _LOADED = ns
    return ns


if __name__ == "__main__":
    ref = load_reference()
    print("ADV:", sorted(ref.registry.ADV_REGISTRY))
    print("LOSS:", sorted(ref.registry.LOSS_REGISTRY))
    print("mlp_policy:", ref.mlp_policy is not None, getattr(ref, "mlp_policy_error", ""))
    print("types:", ref.embodied_types is not None, getattr(ref, "embodied_types_error", ""))
