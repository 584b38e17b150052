"""CPU oracle for the VPL-ArcFace head: TEST INFRASTRUCTURE ONLY.

Restates ``VPLArcFace.forward`` (main_code/utils/criterion.py:688-762) + nn.CrossEntropyLoss + accuracy
(model_utils.py:179-182) and their closed-form backward in float64.  Pinned against the reference's own autograd by
oracle/make_golden_vpl.py (tests/golden/vpl_*.npz).  Only tests/ and the golden generator import this module.

Kernel-facing form: with a_j = lamda * 1[life_j > 0] (after the per-step decay, criterion.py:716-717) every non-target
cosine is x^_i . v_j with the mixed class vector v_j = (1 - a_j) w^_j + a_j m^_j (criterion.py:724), the target cosine
is (1 - a_y) x^_i . w^_y + a_y (criterion.py:725), then clamp to +-(1 - 1e-7) (729) and the ArcFace margin with
sqrt(1 - c^2 + 1e-9) (733-739).  The memory bank carries no gradient.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class VplConfig:
    s: float = 64.0
    m: float = 0.5
    easy_margin: bool = True
    lamda: float = 0.15
    delta: int = 100


def make_inputs(B: int, Cn: int, D: int, seed: int, dup_labels: bool = True):
    g = torch.Generator().manual_seed(seed)
    bound = math.sqrt(6.0 / (Cn + D))                                  # xavier_uniform (criterion.py:657)
    W = (torch.rand(Cn, D, generator=g, dtype=torch.float64) * 2 - 1) * bound
    x = torch.randn(B, D, generator=g, dtype=torch.float64) * 2.0
    labels = torch.randint(0, Cn, (B,), generator=g)
    if dup_labels and B >= 4:
        labels[1] = labels[0]                                          # a class with two samples: mem = their mean
    # embeddings correlated with their class centre so that margins and memory matter
    x = x + 6.0 * torch.nn.functional.normalize(W[labels], dim=1)
    return x.float(), W.float(), labels


def update_memory(cfg: VplConfig, mem: torch.Tensor, life: torch.Tensor, x: torch.Tensor, labels: torch.Tensor):
    """criterion.py:703-717: mem[c] = mean of the raw features of class c in the batch, life[c] = delta, life -= 1."""
    mem, life = mem.clone(), life.clone()
    for c in torch.unique(labels):
        mem[c] = x[labels == c].mean(dim=0)
        life[c] = cfg.delta
    life = life - 1
    return mem, life


def loss_and_grads(cfg: VplConfig, x, W, labels, mem, life, training_flag: bool = True, grad_scale: float = 1.0,
                   dtype=torch.float64):
    x, W, mem, life = x.to(dtype), W.to(dtype), mem.to(dtype), life.to(dtype)
    B, Cn = x.shape[0], W.shape[0]
    ar = torch.arange(B)
    if training_flag:
        mem, life = update_memory(cfg, mem, life, x, labels)
        # active_mask is a float32 tensor in the reference, so both interpolation weights are formed in float32
        # (criterion.py:717,724-725): a_j = fl32(mask * lamda), 1 - a_j = fl32(1 - fl32(mask * lamda))
        mask32 = (life > 0).float()
        alpha = (mask32 * cfg.lamda).to(dtype)                         # [C]
        beta = (1 - mask32 * cfg.lamda).to(dtype)                      # [C], the weight of cos(x, w_j)
    else:
        alpha = torch.zeros(Cn, dtype=dtype)
        beta = torch.ones(Cn, dtype=dtype)
    xn = x.norm(dim=1, keepdim=True)
    xh = x / xn.clamp_min(1e-12)
    wn = W.norm(dim=1, keepdim=True)
    wh = W / wn.clamp_min(1e-12)
    mh = mem / mem.norm(dim=1, keepdim=True).clamp_min(1e-12)
    V = beta.unsqueeze(1) * wh + alpha.unsqueeze(1) * mh               # mixed class vectors
    raw = xh @ V.t()                                                   # cosine1 everywhere
    tw = (xh * wh[labels]).sum(1)
    raw[ar, labels] = beta[labels] * tw + alpha[labels]                # cosine2 at the target
    lo, hi = -1 + 1e-7, 1 - 1e-7
    c = raw.clamp(lo, hi)
    inside = ((raw >= lo) & (raw <= hi)).to(dtype)
    t = c[ar, labels]
    # the reference keeps these four constants as float32 buffers (criterion.py:665-668)
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
    cm, sm = f32(math.cos(cfg.m)), f32(math.sin(cfg.m))
    th, mm = f32(math.cos(math.pi - cfg.m)), f32(math.sin(math.pi - cfg.m) * cfg.m)
    sin_t = torch.sqrt(1.0 - t * t + 1e-9)
    phi = t * cm - sin_t * sm
    dphi = cm + t / sin_t * sm
    take = (t > 0) if cfg.easy_margin else (t > th)
    alt = t if cfg.easy_margin else t - mm
    z = cfg.s * c
    z[ar, labels] = cfg.s * torch.where(take, phi, alt)
    dz_dc = cfg.s * inside
    dz_dc[ar, labels] = cfg.s * torch.where(take, dphi, torch.ones_like(t)) * inside[ar, labels]
    pre = cfg.s * c
    lse = torch.logsumexp(z, dim=1)
    loss = (lse - z[ar, labels]).mean()
    cnt = (pre > pre[ar, labels].unsqueeze(1)).sum(1)
    acc1 = 100.0 * (cnt < 1).to(dtype).mean()
    acc5 = 100.0 * (cnt < 5).to(dtype).mean()
    P = torch.exp(z - lse.unsqueeze(1))
    dz = P.clone()
    dz[ar, labels] -= 1.0
    dz *= grad_scale / B
    dc = dz * dz_dc                                                     # d loss / d cosine (final, per entry)
    # non-target entries reach x^ through v_j and w^_j through (1 - a_j) x^_i; the target entry through (1 - a_y) w^_y
    dct = dc[ar, labels].clone()
    dc_nt = dc.clone()
    dc_nt[ar, labels] = 0.0
    dxh = dc_nt @ V + (dct * beta[labels]).unsqueeze(1) * wh[labels]
    dwh = beta.unsqueeze(1) * (dc.t() @ xh)
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / xn
    dW = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) / wn
    return dict(loss=loss, acc1=acc1, acc5=acc5, norms=xn.reshape(-1), dx=dx, dW=dW, logits=z, pre=pre, mem=mem, life=life,
                alpha=alpha, beta=beta)
