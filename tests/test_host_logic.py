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
