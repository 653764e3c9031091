"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/c3b200.h declares, and its pure-host entry points behave (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from c3_b200 import build, _lib
    build.build_library()          # no-op when up to date; nvcc cross-compiles without a GPU
    return _lib.load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "c3b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(c3b_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    from c3_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 14
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/c3b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "c3b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert "#include <torch" not in hdr and "ATen" not in hdr


def test_version_and_workspace_sizes(lib):
    assert lib.c3b_version() >= 100
    # headline config: generators + row sums + segment scratch, well below 1 GB
    n = lib.c3b_pwc_workspace_bytes(4096, 2, 1000, 9, 0, 0)
    assert 0 < n < (1 << 30)
    assert lib.c3b_pwc_workspace_bytes(0, 2, 1000, 9, 0, 0) == 0
    # Lindblad D=81 needs the per-CTA global workspace
    assert lib.c3b_pwc_workspace_bytes(1024, 2, 1000, 9, 1, 0) > n
    assert lib.c3b_product_workspace_bytes(16, 100, 9) > 0


def test_kernel_selection(lib):
    assert lib.c3b_pwc_path(2, 9, 0) == 1      # register-resident rows kernel
    assert lib.c3b_pwc_path(2, 3, 0) == 1
    assert lib.c3b_pwc_path(2, 9, 1) == 2      # per-sample models -> CTA kernel (shared memory)
    assert lib.c3b_pwc_path(3, 27, 0) == 2
    assert lib.c3b_pwc_path(2, 81, 0) == 3     # global workspace


def test_argument_errors_have_c3_prefix(lib):
    rc = lib.c3b_pwc_closed(None, None, None, 1.0, 0, 0, 0, 0, 0, None, None, None, 0, None)
    assert rc == -1
    assert lib.c3b_last_error().decode().startswith("C3:ERROR:")
    rc = lib.c3b_set_tuning(b"no_such_key", 1)
    assert rc == -1
    assert b"unknown tuning key" in lib.c3b_last_error()
    assert lib.c3b_set_tuning(b"target_units", 32768) == 0


def test_product_path_fails_loudly_without_library(tmp_path, monkeypatch):
    """The product must not silently fall back when the CUDA library is missing."""
    from c3_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from c3_b200 import engine
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.pwc_closed(np.eye(3), np.zeros((1, 3, 3)), np.zeros((1, 1, 4)), 1.0)


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "c3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
