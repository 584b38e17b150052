"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference

For every in-scope head of main_code/utils/criterion.py this script
  1. builds the reference nn.Module with the wrapper's constructor arguments (config.py:16-70),
  2. loads seeded inputs from oracle.margin_oracle.make_inputs (regenerable anywhere from the seed),
  3. runs reference forward -> nn.CrossEntropyLoss -> accuracy -> autograd backward in float64 on CPU
     (model_utils.py:176-185), and
  4. asserts that the closed-form restatement in oracle/margin_oracle.py reproduces it, then
  5. stores the reference outputs as the golden vectors.

Nothing here is imported by the product or by the GPU tests; the GPU box has no /root/reference.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import margin_oracle as mo  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def load_reference():
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import criterion as C  # type: ignore
        from main_code.utils.metrics import accuracy  # type: ignore
    return C, accuracy


def build_ref_head(C, case):
    fam, cfg = case["family"], case["cfg"]
    D, nC = case["D"], case["C"]
    with contextlib.redirect_stdout(io.StringIO()):
        if fam == "arcface":
            h = C.ArcFace(D, nC, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin)
        elif fam == "cosface":
            h = C.CosFace(D, nC, s=cfg.s, m=cfg.m)
        elif fam == "sphereface":
            h = C.SphereFace(D, nC, m=cfg.sphere_m)
        elif fam in ("mv_am", "mv_arc"):
            h = C.MV_Softmax(D, nC, margin=cfg.m, mv_weight=cfg.mv_weight, s=cfg.s,
                             margin_type="am" if fam == "mv_am" else "arc")
        elif fam == "curricularface":
            h = C.CurricularFace(D, nC, m=cfg.m, s=cfg.s, momentum=cfg.momentum)
        elif fam == "adaface":
            h = C.AdaFace(D, nC, m=cfg.m, h=cfg.h, s=cfg.s, t_alpha=cfg.t_alpha)
        elif fam == "elastic_cos":
            h = C.ElasticCosFace(D, nC, s=cfg.s, m=cfg.m, std=cfg.std, plus=cfg.plus)
        elif fam == "elastic_arc":
            h = C.ElasticArcFace(D, nC, s=cfg.s, m=cfg.m, std=cfg.std, plus=cfg.plus)
        elif fam == "magface":
            h = C.MagFace(D, nC, s=cfg.s, easy_margin=cfg.easy_margin, l_margin=cfg.l_margin,
                          u_margin=cfg.u_margin, l_a=cfg.l_a, u_a=cfg.u_a)
        else:
            raise ValueError(fam)
    return h


def cases():
    out = []

    def add(name, fam, B=8, Cn=61, seed=0, lambda_g=0.0, grad_scale=1.0, state=None, **kw):
        cfg = mo.HeadConfig.default(fam)
        for k, v in kw.items():
            setattr(cfg, k, v)
        out.append(dict(name=name, family=fam, cfg=cfg, B=B, C=Cn, D=512, seed=seed, lambda_g=lambda_g,
                        grad_scale=grad_scale, state=state or mo.HeadState()))

    add("arcface", "arcface", seed=11)
    add("arcface_easy", "arcface", seed=12, easy_margin=True)
    add("cosface", "cosface", seed=13)
    add("sphereface_m2", "sphereface", seed=14)
    add("sphereface_m4_iter", "sphereface", seed=15, sphere_m=4, state=mo.HeadState(sphere_iter=20000))
    add("mv_am", "mv_am", seed=16)
    add("mv_arc", "mv_arc", seed=17)
    add("curricularface", "curricularface", seed=18)
    add("curricularface_t05", "curricularface", seed=19, state=mo.HeadState(t_buf=0.5))
    add("adaface", "adaface", seed=20)
    add("elastic_cos", "elastic_cos", seed=21)
    add("elastic_cos_plus", "elastic_cos", seed=22, plus=True)
    add("elastic_arc", "elastic_arc", seed=23)
    add("elastic_arc_plus", "elastic_arc", seed=24, plus=True)
    add("magface", "magface", seed=25, lambda_g=35.0)
    add("magface_easy", "magface", seed=26, lambda_g=35.0, easy_margin=True)
    add("arcface_gradscale", "arcface", seed=27, grad_scale=65536.0)
    # ragged sizes: C not a multiple of any tile, B not a multiple of a warp
    add("cosface_ragged", "cosface", B=5, Cn=37, seed=28)
    add("arcface_B1", "arcface", B=1, Cn=19, seed=29)
    return out


def run_reference(Cmod, accuracy, case):
    fam = case["family"]
    x, W, labels = mo.make_inputs(fam, case["B"], case["C"], case["D"], case["seed"])
    head = build_ref_head(Cmod, case).double()
    pname = "weight" if mo.LAYOUT[fam] == "CD" else "kernel"
    with torch.no_grad():
        getattr(head, pname).copy_(W.double())
    st = case["state"]
    if fam == "sphereface":
        head.iter = st.sphere_iter
    if fam == "curricularface":
        head.t = torch.full((1,), st.t_buf, dtype=torch.float64)
    if fam == "adaface":
        head.batch_mean = torch.full((1,), st.batch_mean, dtype=torch.float64)
        head.batch_std = torch.full((1,), st.batch_std, dtype=torch.float64)
    xr = x.double().requires_grad_(True)
    torch.manual_seed(99)                       # RNG state seen by ElasticFace's torch.normal
    (pre, logits), norms, loss_g, one_hot = head(xr, labels)
    loss_id = torch.nn.CrossEntropyLoss()(logits, labels)
    loss = loss_id + case["lambda_g"] * loss_g
    acc1, acc5 = accuracy(pre, labels, topk=(1, 5))
    (loss * case["grad_scale"]).backward()
    new_state = mo.HeadState(
        sphere_iter=getattr(head, "iter", 0) if fam == "sphereface" else 0,
        t_buf=float(head.t) if fam == "curricularface" else 0.0,
        batch_mean=float(head.batch_mean) if fam == "adaface" else 20.0,
        batch_std=float(head.batch_std) if fam == "adaface" else 100.0,
    )
    return dict(loss_id=loss_id.detach(), loss_g=torch.as_tensor(loss_g, dtype=torch.float64).detach(),
                loss=loss.detach(), acc1=acc1[0], acc5=acc5[0], norms=norms.detach().reshape(-1),
                dx=xr.grad, dW=getattr(head, pname).grad, pre=pre.detach(), logits=logits.detach(),
                new_state=new_state), (x, W, labels)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def main():
    Cmod, accuracy = load_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    worst = 0.0
    for case in cases():
        ref, (x, W, labels) = run_reference(Cmod, accuracy, case)
        margins = None
        if case["family"].startswith("elastic"):
            torch.manual_seed(99)
            margins = mo.sample_elastic_margins(case["cfg"], case["B"])
        got = mo.loss_and_grads(case["cfg"], case["state"], x, W, labels, margins=margins,
                                lambda_g=case["lambda_g"], grad_scale=case["grad_scale"])
        errs = dict(
            loss=abs(float(got["loss"] - ref["loss"])) / abs(float(ref["loss"])),
            logits=rel(got["logits"], ref["logits"]),
            pre=rel(got["pre"], ref["pre"]),
            dx=rel(got["dx"], ref["dx"]),
            dW=rel(got["dW"], ref["dW"]),
            norms=rel(got["norms"], ref["norms"]),
        )
        ns, rs = got["new_state"], ref["new_state"]
        assert ns.sphere_iter == rs.sphere_iter
        assert abs(ns.t_buf - rs.t_buf) < 1e-12 and abs(ns.batch_mean - rs.batch_mean) < 1e-9
        assert abs(ns.batch_std - rs.batch_std) < 1e-9
        assert abs(float(got["acc1"]) - float(ref["acc1"])) < 1e-9, (case["name"], got["acc1"], ref["acc1"])
        assert abs(float(got["acc5"]) - float(ref["acc5"])) < 1e-9
        w = max(errs.values())
        worst = max(worst, w)
        print(f"{case['name']:22s} loss={float(ref['loss']):.6f} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
        assert w < 1e-6, (case["name"], errs)   # fp32 one_hot in some reference heads (criterion.py:182) limits fp64 agreement to ~1e-8
        cfg = case["cfg"]
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", case["name"] + ".npz"),
            family=case["family"], B=case["B"], C=case["C"], D=case["D"], seed=case["seed"],
            lambda_g=case["lambda_g"], grad_scale=case["grad_scale"],
            cfg=np.array([repr(cfg.__dict__)]),
            state_in=np.array([case["state"].sphere_iter, case["state"].t_buf, case["state"].batch_mean,
                               case["state"].batch_std], dtype=np.float64),
            state_out=np.array([rs.sphere_iter, rs.t_buf, rs.batch_mean, rs.batch_std], dtype=np.float64),
            loss_id=ref["loss_id"].numpy(), loss_g=ref["loss_g"].numpy(), loss=ref["loss"].numpy(),
            acc1=float(ref["acc1"]), acc5=float(ref["acc5"]),
            norms=ref["norms"].numpy(), dx=ref["dx"].numpy().astype(np.float32),
            dW=ref["dW"].numpy().astype(np.float32),
            x_sum=float(x.double().sum()), W_sum=float(W.double().sum()), labels=labels.numpy(),
        )
    print("worst relative deviation oracle vs reference:", worst)


if __name__ == "__main__":
    main()
