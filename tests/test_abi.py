"""The C-ABI library builds, loads and exports every symbol include/gem_b200.h declares
(no compute: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from globalegomocap_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_functions():
    src = open(os.path.join(REPO, "include", "gem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gem_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from globalegomocap_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert getattr(raw, n) is not None


def test_version_and_error_channel(lib):
    assert lib.gem_version() >= 100
    ctx = ctypes.c_void_p()
    # invalid geometry is rejected before any CUDA call
    rc = lib.gem_ctx_create(ctypes.byref(ctx), 0, 4, 2047, 10, 15, 64, 64, 24)
    assert rc == -1 and b"latent_dim" in lib.gem_last_error()


def test_product_has_no_cpu_fallback():
    import torch
    from globalegomocap_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.GemError):
        engine.Engine(4)


def test_product_never_imports_the_oracle():
    """Neither the oracle nor the reference harness (baseline/) is reachable from the product package or the CLI."""
    assert not re.search(r"^\s*(from|import)\s+(oracle|baseline)\b", open(os.path.join(REPO, "optimize_whole_sequence.py")).read(),
                         flags=re.M)
    pkg = os.path.join(REPO, "globalegomocap_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|baseline)\b", text, flags=re.M), f
