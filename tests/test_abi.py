"""The C-ABI library loads and exports every symbol include/b2e.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from embiggen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    header = open(os.path.join(ROOT, "include", "b2e.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(b2e_[a-z_0-9]+)\s*\(", header)))


def test_header_and_binding_agree():
    assert declared_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    assert _lib.is_built(), "libb2e.so missing: run `python -m embiggen_b200.build`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_config_struct_matches_header_layout():
    # 27 four-byte fields, no padding
    assert ctypes.sizeof(_lib.B2EConfig) == 108
    assert ctypes.sizeof(_lib.B2ECounters) == 48


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.b2e_abi_version() == 1
    handle = ctypes.c_void_p()
    config = _lib.B2EConfig(struct_size=1)
    assert lib.b2e_create(ctypes.byref(config), ctypes.byref(handle)) == _lib.B2E_ERR_INVALID
    assert b"struct_size" in lib.b2e_last_error()
    with pytest.raises(ValueError):
        _lib.check(_lib.B2E_ERR_INVALID)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from embiggen_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        Engine("SkipGram")


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under embiggen_b200/ may import, load or link it."""
    package = os.path.join(ROOT, "embiggen_b200")
    for folder, _, files in os.walk(package):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(folder, name)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), name
                assert "liboracle" not in text and "oracle.h" not in text, name
