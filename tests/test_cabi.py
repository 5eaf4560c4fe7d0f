"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports every
symbol include/tan_b200.h declares (and nothing declared is missing from the ctypes binding), and
reports errors through codes -- no compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tan_b200.h")


@pytest.fixture(scope="module")
def lib():
    from temporalalignnet_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TAN_API\s+[\w\s\*]+?\b(tan_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared_symbols()
    for n in ("tan_linear_bf16", "tan_layernorm", "tan_attention_bf16", "tan_sim_nce_fwd", "tan_nce_from_logits",
              "tan_nce_reduce", "tan_cast_f32_to_bf16", "tan_device_check", "tan_last_error_string"):
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    from temporalalignnet_b200 import _lib
    for n in _declared_symbols():
        assert hasattr(lib, n), f"{n} declared in include/tan_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_abi_version_and_error_string(lib):
    assert lib.tan_abi_version() == 5
    assert isinstance(lib.tan_last_error_string(), bytes)


def test_struct_layouts_match_header():
    from temporalalignnet_b200._lib import LnArgs, SimGeom
    assert ctypes.sizeof(SimGeom) == 40 and SimGeom.col_off.offset == 32
    # 64-bit: pointers 8-byte aligned; the header's field order packs to this size
    assert ctypes.sizeof(LnArgs) == 168 and LnArgs.out_f32.offset == 64 and LnArgs.strideA.offset == 88
    assert LnArgs.raw_strideA.offset == 152 and LnArgs.raw_strideB.offset == 160


def test_no_gpu_means_error_code_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = lib.tan_device_check()
    assert rc in (-4, -2)
    assert lib.tan_last_error_string() != b""
    # a compute entry point refuses too (device check comes first; null pointers are never touched)
    assert lib.tan_cast_f32_to_bf16(None, None, 8, None) < 0


def test_product_has_no_cpu_path():
    import torch
    from temporalalignnet_b200 import TanError, TemporalAligner, TemporalEncoder
    m = TemporalAligner(1, 1, random_pos_start=0)
    with pytest.raises(TanError):
        m(torch.zeros(1, 8, 1024), torch.zeros(1, 2, 512))
    with pytest.raises(TanError):
        TemporalEncoder(128, 1, 2)(torch.zeros(4, 1, 128))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "temporalalignnet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
