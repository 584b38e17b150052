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


def test_step_workspace_struct_mirrors_the_header():
    """MhStepWs (ctypes) must list the fields of mh_step_ws in the header's order and widths (all 8 bytes except the two
    int32 enums): the descriptor crosses the FFI by pointer."""
    from face_recognition_models_b200 import _lib
    src = open(os.path.join(ROOT, "include", "margin_head.h")).read()
    body = re.search(r"typedef struct mh_step_ws \{(.*?)\} mh_step_ws;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.rsplit(" ", 1)[0], decl.split(",")
        first = names[0].rsplit(" ", 1)
        base = first[0].strip()
        for n in [first[1]] + names[1:]:
            fields.append((n.strip().lstrip("*"), 4 if base == "int32_t" and "*" not in n and not base.endswith("*") else 8))
    assert [f[0] for f in fields] == [f[0] for f in _lib.MhStepWs._fields_]
    assert [f[1] for f in fields] == [ctypes.sizeof(f[1]) for f in _lib.MhStepWs._fields_]
    assert fields[-1][0] == "graph_cache"


def test_stash_eligibility_predicates(lib):
    """Host-only: the proven stash (mh_tc_stash_ok) and the guarded stash (mh_tc_stash_guarded_ok) partition the heads."""
    from face_recognition_models_b200 import _lib
    def cfg(fam, s=64.0, **kw):
        c = _lib.MhConfig()
        c.family = _lib.FAMILY[fam]
        c.s = s
        c.m = 0.5
        c.mv_weight = 1.12
        c.sphere_m = 2
        for k, v in kw.items():
            setattr(c, k, v)
        return c
    for fam, s, proven in [("arcface", 64.0, 1), ("cosface", 64.0, 1), ("mv_am", 32.0, 1), ("curricularface", 64.0, 0),
                           ("sphereface", 64.0, 0), ("arcface", 128.0, 0)]:
        c = cfg(fam, s)
        assert lib.mh_tc_stash_ok(ctypes.byref(c), 1000) == proven, fam
        assert lib.mh_tc_stash_guarded_ok(ctypes.byref(c), 1000) == 1 - proven, fam
        assert lib.mh_tc_stash_guarded_ok(ctypes.byref(c), 1) == 0


def test_graph_cache_entry_points_fail_cleanly_without_a_gpu(lib):
    """mh_step_cache_create needs a CUDA device; without one it must report an error (no crash, no handle), and destroying
    a NULL handle is a no-op."""
    import torch
    h = ctypes.c_void_p(0)
    st = lib.mh_step_cache_create(ctypes.byref(h))
    if torch.cuda.is_available():
        assert st == 0 and h.value
        counts = (ctypes.c_int64 * 3)()
        assert lib.mh_step_cache_stats(h, counts) == 0 and list(counts) == [0, 0, 0]
        assert lib.mh_step_cache_destroy(h) == 0
    else:
        assert st != 0 and not h.value and b"mh_step_cache_create" in lib.mh_last_error()
    assert lib.mh_step_cache_destroy(None) == 0


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
    assert lib.mh_fwd_num_tiles(2000128) == 148              # one record per CTA pair and column half (148 SMs)


def test_no_cpu_fallback():
    import torch
    import face_recognition_models_b200 as pkg
    head = pkg.ArcFace(512, 32)
    x = torch.randn(4, 512)
    y = torch.randint(0, 32, (4,))
    with pytest.raises(pkg._lib.MarginHeadError):
        head.fused_loss(x, y)


@pytest.mark.parametrize("units,m_tiles,n_tiles", [(74, 4, 7813), (74, 1, 5), (74, 32, 977), (74, 74, 3), (74, 3, 42),
                                                    (74, 2, 1), (74, 37, 400), (66, 5, 1000)])
def test_a_stationary_schedule_covers_every_tile_once(lib, units, m_tiles, n_tiles):
    """The static schedule of the A-stationary kernels (tc_head.cu StatIter): every (m, n) tile exactly once,
    work balanced across pairs, and at most a handful of resident-tile reloads per pair."""
    import numpy as np
    total = m_tiles * n_tiles
    out = np.full((total + 8, 3), -1, dtype=np.int32)
    cnt = lib.mh_tc_schedule_tiles(units, m_tiles, n_tiles, out.ctypes.data_as(ctypes.c_void_p), total + 8)
    assert cnt == total
    out = out[:total]
    seen = np.zeros((m_tiles, n_tiles), dtype=np.int64)
    np.add.at(seen, (out[:, 1], out[:, 2]), 1)
    assert (seen == 1).all()
    per_pair = np.bincount(out[:, 0], minlength=units)
    if total >= 20 * units:
        assert per_pair.max() <= 1.03 * total / units + 2      # balanced to a few percent
    for p in range(units):
        ms = out[out[:, 0] == p, 1]
        reloads = int((np.diff(ms) != 0).sum()) + (1 if len(ms) else 0)
        assert reloads <= m_tiles + 1
        if p < (units // m_tiles) * m_tiles:
            assert reloads <= 1                                  # bound pairs never reload x^
    assert lib.mh_tc_schedule_tiles(74, 75, 10, None, 0) < 0
