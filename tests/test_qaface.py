"""QAFace row: reference criterion.py:1331-1520, three consecutive steps (EMA statistics, populated
memory bank, delta = 2 expiry).  CPU: the oracle against the goldens produced from the reference's own autograd
(oracle/make_golden_qaface.py).  GPU: the CUDA path (mh_vpl_mix with a binary mask + the fused tensor-core pipeline with
an external target cosine) against the same goldens - including the gradient w.r.t. ``minput`` on the first step - and
against the oracle at a BASELINE-like shape."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import qaface_oracle as qo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qaface_*.npz")))
LOSS_REL_TC, GRAD_COS_TC, GRAD_NORM_TC = 2e-3, 0.9995, 1e-2


def cfg_of(z):
    return qo.QaConfig(s=float(z["s"]), m=float(z["m"]), easy_margin=bool(z["easy_margin"]), delta=int(z["delta"]),
                       tto=float(z["tto"]), alpha=float(z["alpha"]))


def cosim(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().flatten().cpu(), b.double().flatten().cpu(), dim=0))


def test_goldens_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    cfg, B, Cn, seed, gs = cfg_of(z), int(z["B"]), int(z["C"]), int(z["seed"]), float(z["grad_scale"])
    st = qo.QaState.fresh(Cn)
    for step in range(int(z["n_steps"])):
        x, minput, W, labels = qo.make_inputs(B, Cn, 512, seed * 10 + step)
        r = qo.loss_and_grads(cfg, st, x, minput, W, labels, True, gs, minput_grad=(step == 0))
        st = r["state"]
        assert abs(float(r["loss"]) - float(z[f"s{step}_loss"])) < 1e-10 * abs(float(z[f"s{step}_loss"]))
        assert abs(float(r["acc1"]) - float(z[f"s{step}_acc1"])) < 1e-9 and abs(float(r["acc5"]) - float(z[f"s{step}_acc5"])) < 1e-9
        assert np.allclose(r["dx"].numpy(), z[f"s{step}_dx"], rtol=1e-8, atol=1e-12)
        assert np.allclose(r["dW"].numpy(), z[f"s{step}_dW"], rtol=1e-8, atol=1e-12)
        if step == 0:
            assert np.allclose(r["dminput"].numpy(), z["s0_dminput"], rtol=1e-8, atol=1e-12)
        assert int((st.life > 0).sum()) == int(z[f"s{step}_n_active"])
        assert abs(st.muy - float(z[f"s{step}_muy"])) < 1e-10 and abs(st.std - float(z[f"s{step}_std"])) < 1e-10
        assert abs(float(st.mem.sum()) - float(z[f"s{step}_mem_sum"])) < 1e-9 * max(1.0, abs(float(z[f"s{step}_mem_sum"])))


def test_injection_gate_fires_on_both_sides():
    """The synthetic minput magnitudes must exercise |z| < tto and |z| >= tto (criterion.py:1412)."""
    cfg = qo.QaConfig(tto=1.0)
    x, minput, W, labels = qo.make_inputs(64, 61, 512, 5)
    mag = minput.double().norm(dim=1)
    z = (mag - mag.mean()) / (mag.std() + 1e-6)
    assert int((z.abs() < cfg.tto).sum()) > 0 and int((z.abs() >= cfg.tto).sum()) > 0


def test_module_contract_host():
    import face_recognition_models_b200 as pkg
    h = pkg.QAFace(512, 50, s=64.0, m=0.5, easy_margin=True, delta=1000, tto=2.0, alpha=0.99)
    # criterion.py:1375-1391: parameter and buffer names, shapes, dtypes
    assert list(h.state_dict()) == ["weight", "mem", "life", "muy", "std", "cos_m", "sin_m", "th", "mm"]
    assert tuple(h.weight.shape) == (50, 512) and tuple(h.mem.shape) == (50, 512) and tuple(h.life.shape) == (50,)
    assert float(h.muy) == 0.0 and float(h.std) == 1.0 and h.norm_training_flag is True
    h.change_training_mode(False)
    assert h.norm_training_flag is False
    f = h.injection_cal(torch.tensor([-3.0, -1.0, 0.0, 1.0, 3.0]))
    assert torch.allclose(f, torch.tensor([0.0, float(np.e), 1.0, float(np.exp(-1.0)), 0.0]))
    with pytest.raises(NotImplementedError):
        h(torch.zeros(2, 512), torch.zeros(2, 512), torch.zeros(2, dtype=torch.long))


def test_constructor_signature_equals_reference():
    import inspect
    import sys
    import face_recognition_models_b200 as pkg
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference not present (GPU box)")
    sys.path.insert(0, ref_root)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import criterion as C
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]  # noqa: E731
    assert sig(C.QAFace.__init__) == sig(pkg.QAFace.__init__)


def run_cuda_steps(pkg, cfg, B, Cn, seed, n_steps, gs):
    head = pkg.QAFace(512, Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin, delta=cfg.delta, tto=cfg.tto, alpha=cfg.alpha).cuda()
    out = []
    for step in range(n_steps):
        x, minput, W, labels = qo.make_inputs(B, Cn, 512, seed * 10 + step)
        with torch.no_grad():
            head.weight.copy_(W.cuda())
        head.weight.grad = None
        xg = x.cuda().requires_grad_(True)
        mg = minput.cuda().requires_grad_(step == 0)
        o = head.fused_loss(xg, mg, labels.cuda())
        (o.loss * gs).backward()
        torch.cuda.synchronize()
        out.append(dict(loss=float(o.loss), acc1=float(o.acc1), acc5=float(o.acc5), dx=xg.grad.cpu(), dW=head.weight.grad.cpu(),
                        dminput=mg.grad.cpu() if step == 0 else None, mem=head.mem.cpu().clone(), life=head.life.cpu().clone(),
                        muy=float(head.muy), std=float(head.std), inputs=(x, minput, W, labels)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_matches_reference_golden(path):
    import face_recognition_models_b200 as pkg
    z = np.load(path)
    cfg, B, Cn, seed, gs = cfg_of(z), int(z["B"]), int(z["C"]), int(z["seed"]), float(z["grad_scale"])
    for step, r in enumerate(run_cuda_steps(pkg, cfg, B, Cn, seed, int(z["n_steps"]), gs)):
        ref_loss = float(z[f"s{step}_loss"])
        assert abs(r["loss"] - ref_loss) <= LOSS_REL_TC * abs(ref_loss)
        assert abs(r["acc1"] - float(z[f"s{step}_acc1"])) < 1e-3 and abs(r["acc5"] - float(z[f"s{step}_acc5"])) < 1e-3
        pairs = [(r["dx"], torch.from_numpy(z[f"s{step}_dx"])), (r["dW"], torch.from_numpy(z[f"s{step}_dW"]))]
        if step == 0:
            pairs.append((r["dminput"], torch.from_numpy(z["s0_dminput"])))
        for got, ref in pairs:
            assert cosim(got, ref) >= GRAD_COS_TC
            assert abs(float(got.double().norm()) - float(ref.norm())) <= GRAD_NORM_TC * float(ref.norm())
        assert int((r["life"] > 0).sum()) == int(z[f"s{step}_n_active"])
        assert abs(r["muy"] - float(z[f"s{step}_muy"])) < 1e-4 * abs(float(z[f"s{step}_muy"]))
        assert abs(float(r["mem"].double().sum()) - float(z[f"s{step}_mem_sum"])) < 1e-4 * max(1.0, abs(float(z[f"s{step}_mem_sum"])))


@pytest.mark.gpu
def test_cuda_matches_oracle_at_scale():
    """B = 512, C = 10,575 (BASELINE config 2 shape), three steps with a live memory bank, then a flag-off step."""
    import face_recognition_models_b200 as pkg
    cfg = qo.QaConfig(easy_margin=False, delta=100, tto=1.5, alpha=0.9)
    B, Cn = 512, 10575
    res = run_cuda_steps(pkg, cfg, B, Cn, 4, 3, 1.0)
    st = qo.QaState.fresh(Cn)
    for step, r in enumerate(res):
        x, minput, W, labels = r["inputs"]
        ref = qo.loss_and_grads(cfg, st, x, minput, W, labels, True, 1.0, minput_grad=(step == 0))
        st = ref["state"]
        assert abs(r["loss"] - float(ref["loss"])) <= LOSS_REL_TC * abs(float(ref["loss"]))
        assert abs(r["acc1"] - float(ref["acc1"])) < 0.5 and abs(r["acc5"] - float(ref["acc5"])) < 0.5
        assert cosim(r["dx"], ref["dx"]) >= GRAD_COS_TC and cosim(r["dW"], ref["dW"]) >= GRAD_COS_TC
        if step == 0:
            assert cosim(r["dminput"], ref["dminput"]) >= GRAD_COS_TC
        assert abs(float(r["dW"].double().norm()) - float(ref["dW"].norm())) <= 2e-3 * float(ref["dW"].norm())
        assert torch.allclose(r["mem"].double(), st.mem, atol=1e-5) and torch.equal(r["life"].double(), st.life)
