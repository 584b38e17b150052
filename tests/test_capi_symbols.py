"""CPU: the C-ABI library loads and exports every symbol include/margin_head.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "margin_head.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mh_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from face_recognition_models_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def test_header_lists_functions():
    names = header_functions()
    assert "mh_tc_forward" in names and "mh_prologue_w" in names and len(names) >= 18


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in margin_head.h but not exported"


def test_binding_covers_header(lib):
    from face_recognition_models_b200 import _lib
    assert set(_lib.EXPORTED) == set(header_functions())


def test_config_struct_layout():
    from face_recognition_models_b200 import _lib
    assert ctypes.sizeof(_lib.MhConfig) == 16 * 4          # 4 int32 + 12 float, no padding
    assert _lib.MhConfig.s.offset == 16 and _lib.MhConfig.sphere_lambda.offset == 56


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.mh_version()
    assert isinstance(lib.mh_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Shape checks run on the host before any CUDA call, so they are testable here."""
    from face_recognition_models_b200 import _lib
    st = lib.mh_prologue_w(None, 0, 10, 512, None, 256, None, None, None)
    assert st == -1 and b"null pointer" in lib.mh_last_error()
    st = lib.mh_tc_backward_dw(ctypes.c_void_p(16), 100, 256, ctypes.c_void_p(16), ctypes.c_void_p(16), None)
    assert st == -1 and b"padded" in lib.mh_last_error()
    assert lib.mh_fwd_num_tiles(2000128) == 2 * 7813


def test_no_cpu_fallback():
    import torch
    import face_recognition_models_b200 as pkg
    head = pkg.ArcFace(512, 32)
    x = torch.randn(4, 512)
    y = torch.randint(0, 32, (4,))
    with pytest.raises(pkg._lib.MarginHeadError):
        head.fused_loss(x, y)
