"""CPU oracle for the large-margin cosine-softmax head.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch on CPU, float64 by default) of the
algorithm in the reference's ``main_code/utils/criterion.py`` plus the caller-side
``nn.CrossEntropyLoss`` / top-k accuracy in ``main_code/utils/model_utils.py:176-182`` and
``main_code/utils/metrics.py:3-16``.  It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.  The product
package (``face_recognition_models_b200``) never does.

Parity status: the reference ships no tests or golden vectors for this path ("parity
unpinned" by the reference's own tests).  The oracle is instead pinned against the
reference itself, imported in the build container: ``oracle/make_golden.py`` runs every
in-scope reference head (forward + autograd backward, CPU fp32/fp64), asserts that this
restatement reproduces it, and commits the outputs under ``tests/golden/``.

Two independent pieces live here:

* ``forward_logits``      - materialised ``B x C`` forward, family by family (small sizes).
* ``loss_and_grads``      - closed-form backward (no autograd), the contract the CUDA
                            backward implements (SURVEY.md section 8a "backward contract").
* ``autograd_step``       - same forward but differentiated by torch autograd; used to
                            time a faithful "materialise every B x C temporary" CPU baseline.

Weight layouts follow the reference parameters: ``CD`` = ``weight [C, D]`` (ArcFace,
SphereFace, MV_Softmax), ``DC`` = ``kernel [D, C]`` (all others).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

FAMILIES = (
    "arcface", "cosface", "sphereface", "mv_am", "mv_arc", "curricularface",
    "adaface", "elastic_cos", "elastic_arc", "magface",
)

# parameter layout of each reference head (criterion.py:36,150,243,365,513,831,972,1079,1217)
LAYOUT = {
    "arcface": "CD", "sphereface": "CD", "mv_am": "CD", "mv_arc": "CD",
    "cosface": "DC", "curricularface": "DC", "adaface": "DC",
    "elastic_cos": "DC", "elastic_arc": "DC", "magface": "DC",
}


@dataclass
class HeadConfig:
    """Hyper-parameters of one head; defaults are the wrapper bindings in config.py:16-70."""
    family: str
    s: float = 64.0
    m: float = 0.5
    easy_margin: bool = False
    # SphereFace (criterion.py:29-33): integer m, annealing base/gamma/power/LambdaMin
    sphere_m: int = 2
    # MV-Softmax (criterion.py:334-341)
    mv_weight: float = 1.12
    # CurricularFace (criterion.py:496-501)
    momentum: float = 0.01
    # AdaFace (criterion.py:802-809)
    h: float = 0.333
    t_alpha: float = 0.99
    # ElasticFace (criterion.py:955-962)
    std: float = 0.0125
    plus: bool = False
    # MagFace (criterion.py:1185-1194)
    l_margin: float = 0.45
    u_margin: float = 0.8
    l_a: float = 10.0
    u_a: float = 110.0

    @staticmethod
    def default(family: str) -> "HeadConfig":
        d = {
            "arcface": dict(s=64.0, m=0.5, easy_margin=False),
            "cosface": dict(s=64.0, m=0.35),
            "sphereface": dict(sphere_m=2),
            "mv_am": dict(s=32.0, m=0.35, mv_weight=1.12),
            "mv_arc": dict(s=32.0, m=0.35, mv_weight=1.12),
            "curricularface": dict(s=64.0, m=0.5, momentum=0.01),
            "adaface": dict(s=64.0, m=0.4, h=0.333, t_alpha=0.99),
            "elastic_cos": dict(s=64.0, m=0.35, std=0.0125, plus=False),
            "elastic_arc": dict(s=64.0, m=0.5, std=0.0125, plus=False),
            "magface": dict(s=64.0, easy_margin=False, l_margin=0.45, u_margin=0.8, l_a=10.0, u_a=110.0),
        }[family]
        return HeadConfig(family=family, **d)


@dataclass
class HeadState:
    """Mutable head state the reference keeps between steps."""
    sphere_iter: int = 0                 # SphereFace.iter (criterion.py:33,58)
    t_buf: float = 0.0                   # CurricularFace.t (criterion.py:517,572)
    batch_mean: float = 20.0             # AdaFace.batch_mean (criterion.py:837,881)
    batch_std: float = 100.0             # AdaFace.batch_std  (criterion.py:838,882)


def _clamp_bounds(family: str):
    """(lo, hi) applied to the raw cosine; None = no clamp (ArcFace, criterion.py:267-281)."""
    return {
        "arcface": None,
        "cosface": (-1 + 1e-4, 1 - 1e-4),           # criterion.py:177
        "sphereface": (-1.0, 1.0),                  # criterion.py:81
        "mv_am": (-1 + 1e-7, 1 - 1e-7),             # criterion.py:413
        "mv_arc": (-1 + 1e-7, 1 - 1e-7),
        "curricularface": (-1.0, 1.0),              # criterion.py:546
        "adaface": (-1 + 1e-3, 1 - 1e-3),           # criterion.py:872
        "elastic_cos": (-1 + 1e-7, 1 - 1e-7),       # criterion.py:994
        "elastic_arc": (-1 + 1e-7, 1 - 1e-7),       # criterion.py:1104
        "magface": (-1 + 1e-7, 1 - 1e-7),           # criterion.py:1260
    }[family]


_CHEB = [  # Chebyshev T_m and derivative, criterion.py:40-47
    (lambda c: torch.ones_like(c), lambda c: torch.zeros_like(c)),
    (lambda c: c, lambda c: torch.ones_like(c)),
    (lambda c: 2 * c ** 2 - 1, lambda c: 4 * c),
    (lambda c: 4 * c ** 3 - 3 * c, lambda c: 12 * c ** 2 - 3),
    (lambda c: 8 * c ** 4 - 8 * c ** 2 + 1, lambda c: 32 * c ** 3 - 16 * c),
    (lambda c: 16 * c ** 5 - 20 * c ** 3 + 5 * c, lambda c: 80 * c ** 4 - 60 * c ** 2 + 5),
]


def sphere_lambda(it: int) -> float:
    """criterion.py:60 with base=1000, gamma=0.12, power=1, LambdaMin=5."""
    return max(5.0, 1000.0 * (1 + 0.12 * it) ** (-1))


def normalise(x: torch.Tensor, W: torch.Tensor, layout: str):
    """F.normalize on both sides (criterion.py:65,173-174,263-264,...); returns x_hat, w_hat[C,D], |x|, |w|."""
    xn = x.norm(dim=1, keepdim=True)
    xh = x / xn.clamp_min(1e-12)
    Wc = W if layout == "CD" else W.t()
    wn = Wc.norm(dim=1, keepdim=True)
    wh = Wc / wn.clamp_min(1e-12)
    return xh, wh, xn, wn


def sample_elastic_margins(cfg: HeadConfig, n: int, generator=None, dtype=torch.float32) -> torch.Tensor:
    """Same RNG call as criterion.py:1003-1005 / 1116-1118 (global CPU generator when None)."""
    mg = torch.normal(mean=cfg.m, std=cfg.std, size=(n, 1), generator=generator, dtype=dtype)
    return mg.clamp(cfg.m - cfg.std, cfg.m + cfg.std).squeeze(1)


def row_terms(cfg: HeadConfig, state: HeadState, xn: torch.Tensor, t_raw: torch.Tensor,
              margins: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Everything that depends only on per-row quantities (|x_i| and the raw target cosine).

    Returns a dict with
      scale   [B]  logit scale of the row (s, or |x_i| for SphereFace)
      thr     [B]  hard-negative threshold (MV / Curricular), +inf otherwise
      zt      [B]  target logit z_{i,y_i}
      dzt     [B]  d zt / d t_raw (includes the clamp mask on the target cosine)
      t       [B]  clamped target cosine (pre-margin value at the target column)
      norms   [B]  what the reference returns as `norms`
      loss_g       MagFace regulariser (0 otherwise)
      dzt_dn  [B]  d zt / d |x_i|  through the margin (MagFace) - excludes the SphereFace scale path
      dlg_dn  [B]  d loss_g / d |x_i|
      new_state    HeadState after this forward
    """
    fam = cfg.family
    f64 = xn.dtype
    B = xn.numel()
    xn = xn.reshape(-1)
    bounds = _clamp_bounds(fam)
    if bounds is None:
        t = t_raw
        inside = torch.ones_like(t_raw)
    else:
        lo, hi = bounds
        t = t_raw.clamp(lo, hi)
        inside = ((t_raw >= lo) & (t_raw <= hi)).to(f64)
    s = cfg.s
    new_state = HeadState(state.sphere_iter, state.t_buf, state.batch_mean, state.batch_std)
    scale = torch.full_like(xn, s)
    thr = torch.full_like(xn, float("inf"))
    dzt_dn = torch.zeros_like(xn)
    dlg_dn = torch.zeros_like(xn)
    loss_g = torch.zeros((), dtype=f64)
    norms = xn.clone()
    hard_kind, hard_a, hard_b = 0, 0.0, 0.0

    if fam == "arcface":                                   # criterion.py:280-295
        cm, sm = math.cos(cfg.m), math.sin(cfg.m)
        th, mm = math.cos(math.pi - cfg.m), math.sin(math.pi - cfg.m) * cfg.m
        one_m = 1.0 - t * t
        sine = torch.sqrt(one_m.clamp(1e-9, 1.0))
        dsine = torch.where((one_m >= 1e-9) & (one_m <= 1.0), -t / sine, torch.zeros_like(t))
        phi = t * cm - sine * sm
        dphi = cm - dsine * sm
        take = (t > 0) if cfg.easy_margin else (t > th)
        alt = t if cfg.easy_margin else t - mm
        zt = s * torch.where(take, phi, alt)
        dzt = s * torch.where(take, dphi, torch.ones_like(t))
    elif fam == "cosface":                                 # criterion.py:186-189
        zt = s * (t - cfg.m)
        dzt = s * inside
    elif fam == "sphereface":                              # criterion.py:58-60,85-105
        new_state.sphere_iter = state.sphere_iter + 1
        lamb = sphere_lambda(new_state.sphere_iter)
        Tm, dTm = _CHEB[cfg.sphere_m]
        theta = torch.acos(t)
        k = torch.floor(cfg.sphere_m * theta / math.pi)
        sign = torch.where(k.remainder(2) == 0, torch.ones_like(k), -torch.ones_like(k))
        phi = sign * Tm(t) - 2 * k
        u = (phi - t) / (1 + lamb) + t
        du = (sign * dTm(t) - 1) / (1 + lamb) + 1
        scale = xn.clone()
        zt = u * xn
        dzt = du * xn * inside
    elif fam in ("mv_am", "mv_arc"):                       # criterion.py:420-439
        if fam == "mv_am":
            take = t > cfg.m
            ft = torch.where(take, t - cfg.m, t)
            dft = torch.ones_like(t)
            thr = t - cfg.m
        else:
            cm, sm = math.cos(cfg.m), math.sin(cfg.m)
            sin_t = torch.sqrt(1.0 - t * t + 1e-9)
            ctm = t * cm - sin_t * sm
            take = t > 0
            ft = torch.where(take, ctm, t)
            dft = torch.where(take, cm + t / sin_t * sm, torch.ones_like(t))
            thr = ctm
        zt = s * ft
        dzt = s * dft * inside
        hard_kind, hard_a, hard_b = 1, cfg.mv_weight, cfg.mv_weight - 1.0
    elif fam == "curricularface":                          # criterion.py:552-578
        cm, sm = math.cos(cfg.m), math.sin(cfg.m)
        th, mm = math.cos(math.pi - cfg.m), math.sin(math.pi - cfg.m) * cfg.m
        sin_t = torch.sqrt(1.0 - t * t)
        ctm = t * cm - sin_t * sm
        take = t > th
        ft = torch.where(take, ctm, t - mm)
        dft = torch.where(take, cm + t / sin_t * sm, torch.ones_like(t))
        new_state.t_buf = float(t.mean()) * cfg.momentum + (1 - cfg.momentum) * state.t_buf
        thr = ctm
        zt = s * ft
        dzt = s * dft * inside
        hard_kind, hard_a, hard_b = 2, new_state.t_buf, 0.0
    elif fam == "adaface":                                 # criterion.py:876-904
        eps = 1e-3
        sn = xn.clamp(0.001, 100.0)
        mean = float(sn.mean())
        stdv = float(sn.std()) if B > 1 else float("nan")
        new_state.batch_mean = mean * cfg.t_alpha + (1 - cfg.t_alpha) * state.batch_mean
        new_state.batch_std = stdv * cfg.t_alpha + (1 - cfg.t_alpha) * state.batch_std
        ms = ((sn - new_state.batch_mean) / (new_state.batch_std + eps) * cfg.h).clamp(-1, 1)
        theta = torch.acos(t)
        th_m_raw = theta - cfg.m * ms
        th_m = th_m_raw.clamp(eps, math.pi - eps)
        in2 = ((th_m_raw >= eps) & (th_m_raw <= math.pi - eps)).to(f64)
        zt = s * (torch.cos(th_m) - (cfg.m + cfg.m * ms))
        dzt = s * torch.sin(th_m) / torch.sqrt(1 - t * t) * in2 * inside
    elif fam == "elastic_cos":                             # criterion.py:1003-1015
        assert margins is not None
        mg = _elastic_assign(cfg, margins.to(f64), t)
        zt = s * (t - mg)
        dzt = s * inside
    elif fam == "elastic_arc":                             # criterion.py:1116-1135
        assert margins is not None
        mg = _elastic_assign(cfg, margins.to(f64), t)
        th_m_raw = torch.acos(t) + mg
        th_m = th_m_raw.clamp(0.0, math.pi)
        in2 = ((th_m_raw >= 0) & (th_m_raw <= math.pi)).to(f64)
        zt = s * torch.cos(th_m)
        dzt = s * torch.sin(th_m) / torch.sqrt(1 - t * t) * in2 * inside
    elif fam == "magface":                                 # criterion.py:1244-1290
        xc = xn.clamp(cfg.l_a, cfg.u_a)
        in_n = ((xn >= cfg.l_a) & (xn <= cfg.u_a)).to(f64)
        loss_g = (xc / cfg.u_a ** 2 + 1.0 / xc).mean()
        dlg_dn = (1.0 / cfg.u_a ** 2 - 1.0 / xc ** 2) / B * in_n
        kslope = (cfg.u_margin - cfg.l_margin) / (cfg.u_a - cfg.l_a)
        a = kslope * (xc - cfg.l_a) + cfg.l_margin
        ca, sa = torch.cos(a), torch.sin(a)
        sin_t = torch.sqrt(1.0 - t * t + 1e-9)
        ctm = t * ca - sin_t * sa
        dctm_dt = ca + t / sin_t * sa
        dctm_da = -t * sa - sin_t * ca
        if cfg.easy_margin:
            take = t > 0
            ft = torch.where(take, ctm, t)
            dalt_da = torch.zeros_like(t)
        else:
            take = t > torch.cos(math.pi - a)
            ft = torch.where(take, ctm, t - torch.sin(math.pi - a) * a)
            dalt_da = -(a * ca + sa)
        zt = s * ft
        dzt = s * torch.where(take, dctm_dt, torch.ones_like(t)) * inside
        dzt_dn = s * torch.where(take, dctm_da, dalt_da) * kslope * in_n
        norms = xc
    else:
        raise ValueError(fam)
    return dict(scale=scale, thr=thr, zt=zt, dzt=dzt, t=t, norms=norms, loss_g=loss_g,
                dzt_dn=dzt_dn, dlg_dn=dlg_dn, new_state=new_state,
                hard=(hard_kind, hard_a, hard_b))


def _elastic_assign(cfg: HeadConfig, margins: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """plus=True re-assignment exactly as written at criterion.py:1007-1012 (index by the permutation)."""
    if not cfg.plus:
        return margins
    rank = torch.sort(t, descending=True)[1]
    ms = torch.sort(margins)[0]
    return ms[rank]


def forward_logits(cfg: HeadConfig, state: HeadState, x, W, labels, margins=None, dtype=torch.float64):
    """Materialised forward: returns dict(pre, logits, raw, xh, wh, xn, wn, rows...)."""
    layout = LAYOUT[cfg.family]
    x = x.to(dtype)
    W = W.to(dtype)
    xh, wh, xn, wn = normalise(x, W, layout)
    raw = xh @ wh.t()
    B = x.shape[0]
    ar = torch.arange(B)
    rows = row_terms(cfg, state, xn.reshape(-1), raw[ar, labels], margins)
    bounds = _clamp_bounds(cfg.family)
    if bounds is None:
        c = raw
        inside = torch.ones_like(raw)
    else:
        c = raw.clamp(*bounds)
        inside = ((raw >= bounds[0]) & (raw <= bounds[1])).to(dtype)
    kind, ha, hb = rows["hard"]
    thr = rows["thr"].unsqueeze(1)
    if kind == 1:      # MV-Softmax re-weighting, criterion.py:433-435
        hard = c > thr
        u = torch.where(hard, ha * c + hb, c)
        du = torch.where(hard, torch.full_like(c, ha), torch.ones_like(c))
    elif kind == 2:    # CurricularFace modulation with the *updated* t, criterion.py:575
        hard = c > thr
        u = torch.where(hard, c * (ha + c), c)
        du = torch.where(hard, ha + 2 * c, torch.ones_like(c))
    else:
        u, du = c, torch.ones_like(c)
    scale = rows["scale"].unsqueeze(1)
    z = scale * u
    dz_dc = scale * du * inside
    z[ar, labels] = rows["zt"]
    dz_dc[ar, labels] = rows["dzt"]
    pre = scale * c
    return dict(pre=pre, logits=z, dz_dc=dz_dc, u=u, xh=xh, wh=wh, xn=xn.reshape(-1), wn=wn.reshape(-1),
                rows=rows, layout=layout)


def loss_and_grads(cfg: HeadConfig, state: HeadState, x, W, labels, margins=None,
                   lambda_g: float = 0.0, grad_scale: float = 1.0, dtype=torch.float64):
    """Closed-form forward + backward of loss = CE(logits, labels) + lambda_g * loss_g.

    Returns loss_id, loss_g, loss, acc1, acc5 (percent), norms, dx, dW (in the parameter's own
    layout), rank counts, lse and the new head state.  grad_scale is the upstream gradient
    (e.g. the GradScaler factor, model_utils.py:185).
    """
    f = forward_logits(cfg, state, x, W, labels, margins, dtype)
    z, rows = f["logits"], f["rows"]
    B = z.shape[0]
    ar = torch.arange(B)
    lse = torch.logsumexp(z, dim=1)
    loss_id = (lse - z[ar, labels]).mean()                        # nn.CrossEntropyLoss, model_utils.py:179
    loss_g = rows["loss_g"]
    loss = loss_id + lambda_g * loss_g
    pre = f["pre"]
    cnt = (pre > pre[ar, labels].unsqueeze(1)).sum(dim=1)           # metrics.py:8 (topk on pre-margin logits)
    acc1 = 100.0 * (cnt < 1).to(dtype).mean()
    acc5 = 100.0 * (cnt < 5).to(dtype).mean()

    P = torch.exp(z - lse.unsqueeze(1))
    dz = P.clone()
    dz[ar, labels] -= 1.0
    dz *= grad_scale / B
    dc = dz * f["dz_dc"]
    dxh = dc @ f["wh"]
    dwh = dc.t() @ f["xh"]
    # d loss / d |x_i|: SphereFace scale path (criterion.py:95,105) + MagFace margin / loss_g paths
    dn = dz[ar, labels] * rows["dzt_dn"] + grad_scale * lambda_g * rows["dlg_dn"]
    if cfg.family == "sphereface":
        u = f["u"].clone()
        u[ar, labels] = rows["zt"] / f["xn"]
        dn = dn + (dz * u).sum(dim=1)
    xh, wh, xn, wn = f["xh"], f["wh"], f["xn"], f["wn"]
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / xn.unsqueeze(1) + dn.unsqueeze(1) * xh
    dWc = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) / wn.unsqueeze(1)
    dW = dWc if f["layout"] == "CD" else dWc.t().contiguous()
    return dict(loss_id=loss_id, loss_g=loss_g, loss=loss, acc1=acc1, acc5=acc5, norms=rows["norms"],
                dx=dx, dW=dW, rank_count=cnt, lse=lse, zt=rows["zt"], new_state=rows["new_state"],
                pre=pre, logits=z)


def autograd_step(cfg: HeadConfig, state: HeadState, x, W, labels, margins=None, lambda_g: float = 0.0,
                  dtype=torch.float32):
    """Forward with every B x C temporary materialised + torch autograd backward (CPU baseline leg).

    This is the shape of work the reference performs on CPU: normalise, one B x C GEMM, a chain of
    whole-tensor elementwise passes, CrossEntropyLoss, top-k accuracy, autograd backward.
    """
    x = x.detach().to(dtype).requires_grad_(True)
    W = W.detach().to(dtype).requires_grad_(True)
    layout = LAYOUT[cfg.family]
    xn = x.norm(dim=1, keepdim=True)
    xh = torch.nn.functional.normalize(x, dim=1)
    wh = torch.nn.functional.normalize(W, dim=1 if layout == "CD" else 0)
    raw = xh @ (wh.t() if layout == "CD" else wh)
    B = x.shape[0]
    ar = torch.arange(B)
    bounds = _clamp_bounds(cfg.family)
    c = raw if bounds is None else raw.clamp(*bounds)
    with torch.no_grad():
        rows_ng = row_terms(cfg, state, xn.detach().reshape(-1).double(), raw.detach()[ar, labels].double(), margins)
    # differentiable target logit: recompute zt through autograd from c[ar, labels] and xn
    t = c[ar, labels]
    zt = _target_logit_autograd(cfg, rows_ng, t, xn.reshape(-1), margins)
    kind, ha, hb = rows_ng["hard"]
    thr = rows_ng["thr"].to(dtype).unsqueeze(1)
    if kind == 1:
        u = torch.where(c > thr, ha * c + hb, c)
    elif kind == 2:
        u = torch.where(c > thr, c * (ha + c), c)
    else:
        u = c
    scale = xn if cfg.family == "sphereface" else cfg.s
    one_hot = torch.zeros_like(c)
    one_hot.scatter_(1, labels.view(-1, 1), 1.0)
    z = (1.0 - one_hot) * (u * scale) + one_hot * zt.unsqueeze(1)
    pre = (c * scale).detach()
    loss_id = torch.nn.functional.cross_entropy(z, labels)
    loss_g = _loss_g_autograd(cfg, xn.reshape(-1))
    loss = loss_id + lambda_g * loss_g
    _, pred = pre.topk(min(5, pre.shape[1]), 1, True, True)
    correct = pred.eq(labels.view(-1, 1))
    acc1 = 100.0 * correct[:, :1].any(1).float().mean()
    acc5 = 100.0 * correct.any(1).float().mean()
    loss.backward()
    return dict(loss=loss.detach(), loss_id=loss_id.detach(), acc1=acc1, acc5=acc5, dx=x.grad, dW=W.grad,
                new_state=rows_ng["new_state"])


def _loss_g_autograd(cfg, xn):
    if cfg.family != "magface":
        return torch.zeros((), dtype=xn.dtype)
    xc = xn.clamp(cfg.l_a, cfg.u_a)
    return (xc / cfg.u_a ** 2 + 1.0 / xc).mean()


def _target_logit_autograd(cfg, rows_ng, t, xn, margins):
    fam, s = cfg.family, cfg.s
    dt = t.dtype
    if fam == "arcface":
        cm, sm = math.cos(cfg.m), math.sin(cfg.m)
        sine = torch.sqrt((1.0 - t * t).clamp(1e-9, 1.0))
        phi = t * cm - sine * sm
        if cfg.easy_margin:
            return s * torch.where(t > 0, phi, t)
        return s * torch.where(t > math.cos(math.pi - cfg.m), phi, t - math.sin(math.pi - cfg.m) * cfg.m)
    if fam == "cosface":
        return s * (t - cfg.m)
    if fam == "sphereface":
        lamb = sphere_lambda(rows_ng["new_state"].sphere_iter)
        Tm = _CHEB[cfg.sphere_m][0]
        k = torch.floor(cfg.sphere_m * torch.acos(t.detach()) / math.pi)
        phi = ((-1.0) ** k) * Tm(t) - 2 * k
        return ((phi - t) / (1 + lamb) + t) * xn
    if fam == "mv_am":
        return s * torch.where(t > cfg.m, t - cfg.m, t)
    if fam == "mv_arc":
        ctm = t * math.cos(cfg.m) - torch.sqrt(1.0 - t * t + 1e-9) * math.sin(cfg.m)
        return s * torch.where(t > 0, ctm, t)
    if fam == "curricularface":
        ctm = t * math.cos(cfg.m) - torch.sqrt(1.0 - t * t) * math.sin(cfg.m)
        return s * torch.where(t > math.cos(math.pi - cfg.m), ctm, t - math.sin(math.pi - cfg.m) * cfg.m)
    if fam == "adaface":
        eps = 1e-3
        st = rows_ng["new_state"]
        sn = xn.detach().clamp(0.001, 100.0)
        ms = ((sn - st.batch_mean) / (st.batch_std + eps) * cfg.h).clamp(-1, 1)
        th_m = (torch.acos(t) - cfg.m * ms).clamp(eps, math.pi - eps)
        return s * (torch.cos(th_m) - (cfg.m + cfg.m * ms))
    if fam == "elastic_cos":
        mg = _elastic_assign(cfg, margins.to(dt), t.detach())
        return s * (t - mg)
    if fam == "elastic_arc":
        mg = _elastic_assign(cfg, margins.to(dt), t.detach())
        return s * torch.cos((torch.acos(t) + mg).clamp(0.0, math.pi))
    if fam == "magface":
        xc = xn.clamp(cfg.l_a, cfg.u_a)
        a = (cfg.u_margin - cfg.l_margin) / (cfg.u_a - cfg.l_a) * (xc - cfg.l_a) + cfg.l_margin
        ctm = t * torch.cos(a) - torch.sqrt(1.0 - t * t + 1e-9) * torch.sin(a)
        if cfg.easy_margin:
            return s * torch.where(t > 0, ctm, t)
        return s * torch.where(t > torch.cos(math.pi - a), ctm, t - torch.sin(math.pi - a) * a)
    raise ValueError(fam)


# ---------------------------------------------------------------------------------------------
# deterministic synthetic inputs shared by the golden generator, the tests and the benchmarks
# ---------------------------------------------------------------------------------------------

def make_inputs(family: str, B: int, C: int, D: int = 512, seed: int = 0, trained_frac: float = 0.5,
                x_scale=(0.5, 5.0)):
    """Seeded inputs (numpy PCG64, stable across versions): W in the family's layout, x, labels.

    Half of the rows are "trained-like" (x_i close to its class centre, target cosine ~0.7) so
    that both margin branches and the hard-negative masks fire; row norms span `x_scale`*sqrt(D)
    so MagFace [l_a,u_a] and the AdaFace clamp at 100 are exercised.
    """
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    Wc = rng.standard_normal((C, D)).astype(np.float32) * 0.05
    labels = rng.integers(0, C, size=(B,)).astype(np.int64)
    x = rng.standard_normal((B, D)).astype(np.float32)
    ntr = int(B * trained_frac)
    if ntr > 0:
        wy = Wc[labels[:ntr]]
        wy = wy / np.linalg.norm(wy, axis=1, keepdims=True)
        g = rng.standard_normal((ntr, D)).astype(np.float32)
        g = g / np.linalg.norm(g, axis=1, keepdims=True)
        mix = rng.uniform(0.3, 1.2, size=(ntr, 1)).astype(np.float32)
        v = wy + mix * g
        x[:ntr] = v / np.linalg.norm(v, axis=1, keepdims=True) * math.sqrt(D)
    rs = rng.uniform(x_scale[0], x_scale[1], size=(B, 1)).astype(np.float32)
    x = x * rs
    W = Wc if LAYOUT[family] == "CD" else np.ascontiguousarray(Wc.T)
    return torch.from_numpy(x), torch.from_numpy(W), torch.from_numpy(labels)
