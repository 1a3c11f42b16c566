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
    # 28 four-byte fields, no padding
    assert ctypes.sizeof(_lib.B2EConfig) == 112
    assert ctypes.sizeof(_lib.B2ECounters) == 64


def header_struct(name):
    """[(C type, field name, array length or None)] of `typedef struct { ... } name;` in the header."""
    header = open(os.path.join(ROOT, "include", "b2e.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    body = re.search(r"typedef struct \{([^}]*)\}\s*" + name + r"\s*;", header).group(1)
    fields = []
    for declaration in body.split(";"):
        match = re.match(r"\s*(\w+)\s+(\w+)(?:\[(\d+)\])?\s*$", declaration)
        if match:
            fields.append((match.group(1), match.group(2), int(match.group(3)) if match.group(3) else None))
    return fields


@pytest.mark.parametrize("name,mirror", [("b2e_config", _lib.B2EConfig), ("b2e_counters", _lib.B2ECounters),
                                         ("b2e_perceptron_config", _lib.B2EPerceptronConfig)])
def test_ctypes_mirrors_follow_the_header_field_by_field(name, mirror):
    ctypes_of = {"uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64,
                 "float": ctypes.c_float, "double": ctypes.c_double}
    declared = header_struct(name)
    assert [field for _, field, _ in declared] == [field for field, _ in mirror._fields_]
    for (c_type, field, length), (_, python_type) in zip(declared, mirror._fields_):
        expected = ctypes_of[c_type] * length if length else ctypes_of[c_type]
        assert python_type is expected or (length and python_type._type_ is ctypes_of[c_type]
                                           and python_type._length_ == length), field


def test_integration_stub_lists_the_config_fields_in_header_order():
    """INTEGRATION.md shows the ctypes stub a maintainer would write: keep it in step."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = text[text.index("class b2e_config(ctypes.Structure)"):text.index("def fit_transform(graph")]
    names = re.findall(r'"(\w+)"', stub)
    assert names == [field for _, field, _ in header_struct("b2e_config")]


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.b2e_abi_version() == _lib.ABI_VERSION == 3
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
    lib = _lib.load()
    assert lib.b2e_select_device(0) == _lib.B2E_ERR_CUDA and b"no CUDA device" in lib.b2e_last_error()
    # the feature-less perceptron names its device through that call: it fails here, loudly
    from embiggen_b200.edge_prediction import PerceptronEdgePredictionB200
    from embiggen_b200.graph import erdos_renyi
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        PerceptronEdgePredictionB200(edge_features="Degree", number_of_epochs=1).fit(erdos_renyi(50, 100, seed=1))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under embiggen_b200/ may import, load or link it."""
    package = os.path.join(ROOT, "embiggen_b200")
    for folder, _, files in os.walk(package):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(folder, name)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), name
                assert "liboracle" not in text and "oracle.h" not in text, name
