"""CPU: host-side logic of the drop-in modules (no CUDA needed)."""
import inspect
import math
import os
import sys

import pytest
import torch

import face_recognition_models_b200 as pkg
from face_recognition_models_b200.functional import sphere_lambda

REF = "/root/reference"


def test_state_dict_names_and_layouts_match_reference():
    assert list(pkg.ArcFace(512, 10).state_dict()) == ["weight"]
    assert list(pkg.SphereFace(512, 10).state_dict()) == ["weight"]
    assert list(pkg.MV_Softmax(512, 10).state_dict()) == ["weight"]
    assert list(pkg.CosFace(512, 10).state_dict()) == ["kernel"]
    assert list(pkg.CurricularFace(512, 10).state_dict()) == ["kernel", "t"]
    assert list(pkg.AdaFace(512, 10).state_dict()) == ["kernel", "t", "batch_mean", "batch_std"]
    assert list(pkg.ElasticArcFace(512, 10).state_dict()) == ["kernel"]
    assert list(pkg.MagFace(512, 10).state_dict()) == ["kernel"]
    assert tuple(pkg.ArcFace(512, 10).weight.shape) == (10, 512)
    assert tuple(pkg.MagFace(512, 10).kernel.shape) == (512, 10)
    a = pkg.AdaFace(512, 10)
    assert float(a.batch_mean) == 20.0 and float(a.batch_std) == 100.0


def test_initialisers_follow_reference():
    torch.manual_seed(0)
    k = pkg.CosFace(512, 64).kernel                    # uniform(-1,1).renorm_(2,1,1e-5).mul_(1e5): unit-norm columns
    assert torch.allclose(k.norm(dim=0), torch.ones(64), atol=1e-3)
    c = pkg.CurricularFace(512, 4096).kernel           # normal std 0.01
    assert abs(float(c.std()) - 0.01) < 1e-3
    w = pkg.ArcFace(512, 100).weight                   # xavier_uniform: bound sqrt(6/(fan_in+fan_out))
    assert float(w.abs().max()) <= math.sqrt(6.0 / (512 + 100)) + 1e-6


def test_sphereface_annealing_matches_formula():
    h = pkg.SphereFace(512, 10, m=2)
    for it in (1, 10, 1000, 100000):
        assert sphere_lambda(it) == max(5.0, 1000.0 * (1 + 0.12 * it) ** (-1))      # criterion.py:60 as written
    h._pre_forward(None)
    assert h.iter == 1 and abs(h.lamb - 1000.0 / 1.12) < 1e-9


def test_mv_softmax_margin_type_flip_is_honoured():
    h = pkg.MV_Softmax(512, 10, margin_type="am")
    assert h._engine.cfg.family == pkg._lib.FAMILY["mv_am"]
    h.margin_type = "arc"                              # evaluate_models.py:50,53 does this after construction
    h._pre_forward(None)
    assert h._engine.cfg.family == pkg._lib.FAMILY["mv_arc"]


def test_device_id_model_parallel_is_replaced():
    with pytest.raises(NotImplementedError):
        pkg.ArcFace(512, 10, device_id=[0, 1])


def test_elastic_rejects_ignore_labels():
    h = pkg.ElasticCosFace(512, 10)
    with pytest.raises(ValueError):
        h._sample_margins(torch.zeros(2, 512), torch.tensor([1, -1]))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_constructor_signatures_equal_reference():
    """Same parameter names, order and defaults as the reference classes (drop-in boundary)."""
    sys.path.insert(0, REF)
    import contextlib
    import io
    import warnings
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import criterion as C
    for name in ("SphereFace", "CosFace", "ArcFace", "MV_Softmax", "CurricularFace", "AdaFace", "ElasticCosFace",
                 "ElasticArcFace", "MagFace"):
        ref = inspect.signature(getattr(C, name).__init__)
        got = inspect.signature(getattr(pkg, name).__init__)
        rp = [(p.name, p.default) for p in ref.parameters.values()]
        gp = [(p.name, p.default) for p in got.parameters.values()]
        assert rp == gp, (name, rp, gp)


def test_head_sgd_host_side():
    """HeadSGD is a torch.optim.Optimizer over the heads' class-centre parameters; it refuses foreign modules and CPU
    parameters (the update is CUDA-only: no fallback), and its param_groups drive LR schedulers as usual."""
    import torch
    import face_recognition_models_b200 as pkg
    from face_recognition_models_b200 import _lib as L
    a, b = pkg.ArcFace(512, 32), pkg.CosFace(512, 48)
    opt = pkg.HeadSGD([a, b], lr=0.1, momentum=0.9, weight_decay=5e-4)
    assert isinstance(opt, torch.optim.Optimizer) and opt._step_supports_amp_scaling
    assert [p.shape for p in opt.param_groups[0]["params"]] == [a.weight.shape, b.kernel.shape]
    assert a.head_parameter() is a.weight and b.head_parameter() is b.kernel
    assert a.head_engine() is a._engine
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    opt.step()                                   # no grads yet: nothing to do, must not touch the library
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 0.05) < 1e-12
    a.weight.grad = torch.zeros_like(a.weight)
    with pytest.raises(L.MarginHeadError):
        opt.step()
    with pytest.raises(TypeError):
        pkg.HeadSGD([torch.nn.Linear(4, 4)], lr=0.1)
    with pytest.raises(ValueError):
        pkg.HeadSGD([], lr=0.1)
    with pytest.raises(ValueError):
        pkg.HeadSGD([a], lr=-1.0)
    # the shadow bookkeeping is plain host state
    eng = a.head_engine()
    assert eng._shadow is None
    eng._shadow = (1, 2, 3)
    eng.invalidate_shadow()
    assert eng._shadow is None


def test_sharded_head_rejects_shards_that_are_too_small():
    """chunks of ceil(C/R) can leave trailing ranks with no classes; every rank must refuse, not one."""
    from face_recognition_models_b200.sharded import shard_range
    # world 8, C = 9: chunks of 2 -> ranks 5..7 own 1, 0, 0 classes
    sizes = [e - b for b, e in (shard_range(9, 8, r) for r in range(8))]
    assert min(sizes) == 0 and sum(sizes) == 9
    import face_recognition_models_b200.sharded as sh

    class _Comm:                     # a ShardComm stand-in for rank 0 of 8 (no process group needed)
        group, world, rank, _backend = None, 8, 0, "none"
    real = sh.ShardComm
    sh.ShardComm = lambda group=None: _Comm()
    try:
        with pytest.raises(ValueError, match="at least 2"):
            pkg.ShardedMarginHead("arcface", 9, s=64.0, m=0.5, easy_margin=False)
        head = pkg.ShardedMarginHead("arcface", 64, s=64.0, m=0.5, easy_margin=False)       # 8 classes per rank: fine
        assert head.engine.shard.c_total == 64 and head.engine.shard.c_offset == 0
    finally:
        sh.ShardComm = real


def test_mh_lib_env_selects_the_library(tmp_path, monkeypatch):
    """MH_LIB must be honoured by _lib.load() (variant builds for A/B runs), and a bad path must fail loudly."""
    from face_recognition_models_b200 import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setenv("MH_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(L.MarginHeadError, match="nope.so"):
        L.load()
    monkeypatch.setenv("MH_LIB", L.LIB_PATH)
    assert L.load() is not None


def test_partial_fc_sampling_keeps_positives_and_remaps_labels():
    """class sampling on the device - every positive of the shard is kept, the rest are random negatives,
    the index is sorted, labels are remapped into the sampled id space, an out-of-range label maps outside it."""
    torch.manual_seed(0)
    head = pkg.ShardedMarginHead("arcface", 1000, s=64.0, m=0.5, easy_margin=False, sample_rate=0.1)   # no process group: world 1
    assert head.num_sample == 100 and head.sub_engine.C == 100
    y = torch.randint(0, 1000, (64,))
    index, y_sub = head.sample_classes(y)
    assert index.shape == (100,) and bool((index[1:] > index[:-1]).all())
    assert set(y.tolist()) <= set(index.tolist())                      # all positives sampled
    assert torch.equal(index[y_sub], y)                                # remapped labels point at the same classes
    index2, _ = head.sample_classes(y)
    assert not torch.equal(index, index2)                              # the negatives are re-drawn every step
    yb = y.clone()
    yb[5] = 1000
    _, y_sub_b = head.sample_classes(yb)
    assert int(y_sub_b[5]) == 100 and torch.equal(index2[:0], index2[:0])   # id outside [0, world * k): NaN downstream
    with pytest.raises(ValueError):
        pkg.ShardedMarginHead("arcface", 1000, s=64.0, m=0.5, easy_margin=False, sample_rate=0.0)
    # more distinct positives than num_sample: the dropped targets are poisoned, not silently lost
    small = pkg.ShardedMarginHead("arcface", 1000, s=64.0, m=0.5, easy_margin=False, sample_rate=0.01)   # k = 10
    y_many = torch.arange(0, 640, 10)                                  # 64 distinct classes
    idx, ys = small.sample_classes(y_many)
    kept = torch.isin(y_many, idx)
    assert int(kept.sum()) == 10 and bool((ys[~kept] == 10).all()) and torch.equal(idx[ys[kept]], y_many[kept])


def test_deepcopy_of_a_head_drops_workspaces_and_graph_cache():
    """copy.deepcopy(head) (EMA copies): the copy must not alias the original's device workspaces, raw-pointer step
    descriptor or CUDA-graph cache handle (double destroy); hyper-parameters, modes and parameters are copied."""
    import copy
    import ctypes
    import torch
    import face_recognition_models_b200 as pkg
    head = pkg.CurricularFace(512, 100, m=0.4, s=48.0)
    head.backward_mode = "recompute"
    eng = head._engine
    eng._ws["probe"] = torch.zeros(3)                    # stand-ins for state a forward would have left behind
    eng._graph_cache = ctypes.c_void_p(0)                # NULL handle: destroy is a no-op, but the copy must not share it
    eng._step_key = ("k",)
    eng._shadow = (1, 2, 3)
    dup = copy.deepcopy(head)
    d = dup._engine
    assert d is not eng and d._ws == {} and d._graph_cache is None and d._step_key is None and d._shadow is None
    assert d.family == eng.family and d.C == eng.C and d.backward_mode == "recompute"
    assert abs(d.cfg.s - 48.0) < 1e-6 and abs(d.cfg.m - 0.4) < 1e-6 and d.cfg is not eng.cfg
    assert torch.equal(dup.kernel, head.kernel) and dup.kernel.data_ptr() != head.kernel.data_ptr()
    assert "probe" in eng._ws                            # the original keeps its own


def test_backward_mode_selection_is_host_logic(monkeypatch):
    """Which backward a head takes (0 recompute / 1 proven stash / 2 guarded stash) is decided on the host from the
    hyper-parameters, the padded shape and backward_mode: no device needed."""
    import face_recognition_models_b200 as pkg
    from face_recognition_models_b200.functional import GUARDED_MIN_BC
    monkeypatch.delenv("MH_STASH_GUARDED", raising=False)
    big = (1024, 2_000_128)                                        # cfg4's padded shape
    small = (512, 10_752)                                          # cfg2's
    assert small[0] * small[1] < GUARDED_MIN_BC <= big[0] * big[1]
    arc = pkg.ArcFace(512, 1000)._engine
    assert arc._stash_kind(*small) == 1 and arc._stash_kind(*big) == 1
    arc.backward_mode = "recompute"
    assert arc._stash_kind(*big) == 0
    for head in (pkg.CurricularFace(512, 1000), pkg.SphereFace(512, 1000, m=2), pkg.ArcFace(512, 1000, s=128.0)):
        e = head._engine
        assert e._stash_kind(*small) == 0 and e._stash_kind(*big) == 2          # auto: guarded stash at scale only
        e.backward_mode = "stash"
        assert e._stash_kind(*small) == 2                                        # forced: at any size
        e.backward_mode = "recompute"
        assert e._stash_kind(*big) == 0
        e.backward_mode = "auto"
        monkeypatch.setenv("MH_STASH_GUARDED", "0")
        assert e._stash_kind(*big) == 0
        monkeypatch.delenv("MH_STASH_GUARDED")
    mv = pkg.MV_Softmax(512, 1000, margin=0.35, mv_weight=1.12, s=32.0, margin_type="am")._engine
    assert mv._stash_kind(*small) == 1
