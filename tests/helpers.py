"""Shared test helpers: golden loading, head construction, comparison metrics."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import margin_oracle as mo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    """Golden vectors of the margin heads (lfw_synth_* belongs to tests/test_verification.py, vpl_* to tests/test_vpl.py, qaface_* to tests/test_qaface.py)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith(("lfw_", "vpl_", "qaface_")))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    cfg_dict = ast.literal_eval(str(z["cfg"][0]))
    cfg = mo.HeadConfig(**cfg_dict)
    si, so = z["state_in"], z["state_out"]
    state_in = mo.HeadState(int(si[0]), float(si[1]), float(si[2]), float(si[3]))
    state_out = mo.HeadState(int(so[0]), float(so[1]), float(so[2]), float(so[3]))
    fam = str(z["family"])
    B, Cn, D, seed = int(z["B"]), int(z["C"]), int(z["D"]), int(z["seed"])
    x, W, labels = mo.make_inputs(fam, B, Cn, D, seed)
    # the regenerated inputs must be the ones the reference saw
    assert abs(float(x.double().sum()) - float(z["x_sum"])) < 1e-6 * max(1.0, abs(float(z["x_sum"])))
    assert abs(float(W.double().sum()) - float(z["W_sum"])) < 1e-6 * max(1.0, abs(float(z["W_sum"])))
    assert np.array_equal(labels.numpy(), z["labels"])
    margins = None
    if fam.startswith("elastic"):
        torch.manual_seed(99)
        margins = mo.sample_elastic_margins(cfg, B)
    return dict(name=os.path.basename(path)[:-4], family=fam, cfg=cfg, state_in=state_in, state_out=state_out,
                x=x, W=W, labels=labels, margins=margins, lambda_g=float(z["lambda_g"]),
                grad_scale=float(z["grad_scale"]), loss=float(z["loss"]), loss_id=float(z["loss_id"]),
                loss_g=float(z["loss_g"]), acc1=float(z["acc1"]), acc5=float(z["acc5"]),
                norms=torch.from_numpy(z["norms"]), dx=torch.from_numpy(z["dx"]), dW=torch.from_numpy(z["dW"]))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosim(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0))


def build_head(pkg, fam, cfg, Cn):
    if fam == "arcface":
        return pkg.ArcFace(512, Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin)
    if fam == "cosface":
        return pkg.CosFace(512, Cn, s=cfg.s, m=cfg.m)
    if fam == "sphereface":
        return pkg.SphereFace(512, Cn, m=cfg.sphere_m)
    if fam in ("mv_am", "mv_arc"):
        return pkg.MV_Softmax(512, Cn, margin=cfg.m, mv_weight=cfg.mv_weight, s=cfg.s,
                              margin_type="am" if fam == "mv_am" else "arc")
    if fam == "curricularface":
        return pkg.CurricularFace(512, Cn, m=cfg.m, s=cfg.s, momentum=cfg.momentum)
    if fam == "adaface":
        return pkg.AdaFace(512, Cn, m=cfg.m, h=cfg.h, s=cfg.s, t_alpha=cfg.t_alpha)
    if fam == "elastic_cos":
        return pkg.ElasticCosFace(512, Cn, s=cfg.s, m=cfg.m, std=cfg.std, plus=cfg.plus)
    if fam == "elastic_arc":
        return pkg.ElasticArcFace(512, Cn, s=cfg.s, m=cfg.m, std=cfg.std, plus=cfg.plus)
    if fam == "magface":
        return pkg.MagFace(512, Cn, s=cfg.s, easy_margin=cfg.easy_margin, l_margin=cfg.l_margin,
                           u_margin=cfg.u_margin, l_a=cfg.l_a, u_a=cfg.u_a)
    raise ValueError(fam)


def prime_head(head, fam, W, state, margins):
    with torch.no_grad():
        head._param().copy_(W.to(head._param().device))
    if fam == "sphereface":
        head.iter = state.sphere_iter
    if fam == "curricularface":
        head.t.fill_(state.t_buf)
    if fam == "adaface":
        head.batch_mean.fill_(state.batch_mean)
        head.batch_std.fill_(state.batch_std)
    if margins is not None:
        head._margins_override = margins.to(head._param().device)
    return head
