"""CPU-side checks of the C-ABI library: it loads and exports every symbol the header declares.
No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rlinf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rlinf_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from rlinf_b200 import build

        build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rlinf_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in rlinf_b200/_lib.py"
    assert lib.rb200_abi_version() == 1
    assert lib.rb200_strerror(-2).decode().startswith("non-positive")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from rlinf_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.Rb200Error, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_no_cpu_fallback_for_cpu_tensors():
    import torch

    from rlinf_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Rb200Error):
        _lib.default_device()
    with pytest.raises(_lib.Rb200Error):
        _lib.ptr(torch.zeros(4))


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "rlinf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_registry_surface():
    import rlinf_b200.algorithms as A

    assert {"gae", "grpo"} <= set(A.ADV_REGISTRY)
    assert {"actor_critic", "actor"} <= set(A.LOSS_REGISTRY)
    with pytest.raises(ValueError, match="not registered"):
        A.get_adv_and_returns("nope")
    with pytest.raises(ValueError, match="not registered"):
        A.get_policy_loss("nope")
