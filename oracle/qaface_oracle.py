"""CPU oracle for the QAFace head: TEST INFRASTRUCTURE ONLY.

Restates ``QAFace.forward`` (main_code/utils/criterion.py:1421-1508) + nn.CrossEntropyLoss + accuracy
(model_utils.py:179-182) in float64, in the kernel-facing form the CUDA path uses; gradients come from autograd over this
restatement.  Pinned against the reference's own autograd by oracle/make_golden_qaface.py (tests/golden/qaface_*.npz).
Only tests/ and the golden generator import this module.

Kernel-facing form (norm_training_flag = True):
  * magnitude statistics of ``minput`` (criterion.py:1447-1456): first batch takes the batch mean / unbiased std, later
    batches the EMA with ``alpha``; z_i = (|minput_i| - muy) / (std + 1e-6);
  * injection (1409-1413, 1459-1461): inj_i = [|z_i| < tto] * exp(-z_i) * minput_i / (|minput_i| + 1e-6);
  * memory (1464-1477): mem[c] = mean of inj over the batch rows of class c, life[c] = delta, then life -= 1;
    a_j = 1[life_j > 0] - BINARY, unlike VPL's lamda;
  * non-target cosines (1480-1484): x^_i . v_j with v_j = w^_j when a_j = 0 and m^_j (normalised memory) when a_j = 1;
  * target cosine (1487-1490): x^_i . normalise(w_{y_i} + inj_i) with the RAW class centre - differentiable in x, W
    and minput; the CUDA path takes it as an externally computed per-row term and hands d loss / d t_i back;
  * clamp to +-(1 - 1e-7), ArcFace margin with sqrt(1 - c^2 + 1e-9) (1499-1508).
The reference keeps the autograd graph of ``muy`` / ``std`` across steps (1451-1456), so its second backward fails
whenever ``minput`` requires grad; the statistics are carried detached between steps here (what reloading a checkpoint
does), which is the only way the reference itself can run more than one step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class QaConfig:
    s: float = 64.0
    m: float = 0.5
    easy_margin: bool = True
    delta: int = 1000
    tto: float = 2.0
    alpha: float = 0.99


@dataclass
class QaState:
    mem: torch.Tensor
    life: torch.Tensor
    muy: float = 0.0
    std: float = 1.0

    @staticmethod
    def fresh(Cn: int, D: int = 512, dtype=torch.float64) -> "QaState":
        return QaState(torch.zeros(Cn, D, dtype=dtype), torch.zeros(Cn, dtype=dtype), 0.0, 1.0)


def make_inputs(B: int, Cn: int, D: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    bound = math.sqrt(6.0 / (Cn + D))                                  # xavier_uniform (criterion.py:1376)
    W = (torch.rand(Cn, D, generator=g, dtype=torch.float64) * 2 - 1) * bound
    x = torch.randn(B, D, generator=g, dtype=torch.float64) * 2.0
    labels = torch.randint(0, Cn, (B,), generator=g)
    if B >= 4:
        labels[1] = labels[0]                                          # a class with two samples: mem = their mean
    x = x + 6.0 * torch.nn.functional.normalize(W[labels], dim=1)
    # the "magnitude-sensitive" view: the clean features rescaled row by row (magnitudes spread over +-3 sigma so that
    # both sides of the |z| < tto gate occur) plus noise
    scale = torch.exp(0.5 * torch.randn(B, 1, generator=g, dtype=torch.float64))
    minput = x * scale + 0.5 * torch.randn(B, D, generator=g, dtype=torch.float64)
    return x.float(), minput.float(), W.float(), labels


def forward(cfg: QaConfig, st: QaState, x, minput, W, labels, training_flag: bool = True, dtype=torch.float64):
    """Differentiable restatement.  Returns a dict with loss, pre, logits, t (target cosine before the clamp), the per-class
    mixing mask and the new state (detached)."""
    x, W = x.to(dtype), W.to(dtype)
    B, Cn = x.shape[0], W.shape[0]
    ar = torch.arange(B)
    xn = x.norm(dim=1, keepdim=True)
    xh = x / xn.clamp_min(1e-12)
    wh = W / W.norm(dim=1, keepdim=True).clamp_min(1e-12)
    new = QaState(st.mem.clone().to(dtype), st.life.clone().to(dtype), st.muy, st.std)
    if training_flag:
        minput = minput.to(dtype)
        mag = minput.norm(dim=1, keepdim=True)                          # criterion.py:1447
        mean, sd = mag.mean(), mag.std()
        if st.muy == 0.0:                                                # 1451-1453 (first batch)
            muy, std = mean, sd
        else:                                                            # 1454-1456
            muy = cfg.alpha * st.muy + (1 - cfg.alpha) * mean
            std = cfg.alpha * st.std + (1 - cfg.alpha) * sd
        z = ((mag - muy) / (std + 1e-6)).squeeze(1)                      # 1459
        mask = torch.where(z.abs() < cfg.tto, torch.exp(-z), torch.zeros_like(z))      # 1409-1413
        inj = mask.unsqueeze(1) * minput / (mag + 1e-6)                  # 1461
        with torch.no_grad():                                            # 1468-1474
            for c in torch.unique(labels):
                new.mem[c] = inj[labels == c].mean(dim=0)
                new.life[c] = cfg.delta
        new.life = new.life - 1                                          # 1477
        new.muy, new.std = float(muy.detach()), float(std.detach())
        active = (new.life > 0).to(dtype)                                # 1478
        mh = new.mem / new.mem.norm(dim=1, keepdim=True).clamp_min(1e-12)
        V = (1 - active).unsqueeze(1) * wh + active.unsqueeze(1) * mh    # 1484 as one class vector per column
        raw = xh @ V.t()
        tw = W[labels] + inj                                             # 1487 (raw centre)
        t = (xh * (tw / tw.norm(dim=1, keepdim=True).clamp_min(1e-12))).sum(1)          # 1488-1489
    else:
        active = torch.zeros(Cn, dtype=dtype)
        raw = xh @ wh.t()
        t = raw[ar, labels]
    onehot = torch.zeros(B, Cn, dtype=dtype)
    onehot[ar, labels] = 1.0
    cos = onehot * t.unsqueeze(1) + (1 - onehot) * raw                   # 1493
    lo, hi = -1 + 1e-7, 1 - 1e-7
    c = cos.clamp(lo, hi)                                                # 1499
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731   (float32 buffers, criterion.py:1388-1391)
    cm, sm = f32(math.cos(cfg.m)), f32(math.sin(cfg.m))
    th, mm = f32(math.cos(math.pi - cfg.m)), f32(math.sin(math.pi - cfg.m) * cfg.m)
    sine = torch.sqrt(1.0 - c * c + 1e-9)                                # 1503
    phi = c * cm - sine * sm
    phi = torch.where(c > 0, phi, c) if cfg.easy_margin else torch.where(c > th, phi, c - mm)   # 1506-1509
    logits = cfg.s * (onehot * phi + (1 - onehot) * c)
    pre = cfg.s * c
    loss = torch.nn.functional.cross_entropy(logits, labels)
    cnt = (pre.detach() > pre.detach()[ar, labels].unsqueeze(1)).sum(1)
    return dict(loss=loss, pre=pre, logits=logits, t=t, active=active, norms=xn.reshape(-1), state=new,
                acc1=100.0 * (cnt < 1).to(dtype).mean(), acc5=100.0 * (cnt < 5).to(dtype).mean())


def loss_and_grads(cfg: QaConfig, st: QaState, x, minput, W, labels, training_flag: bool = True, grad_scale: float = 1.0,
                   minput_grad: bool = True, dtype=torch.float64):
    xr = x.to(dtype).clone().requires_grad_(True)
    Wr = W.to(dtype).clone().requires_grad_(True)
    mr = minput.to(dtype).clone().requires_grad_(bool(minput_grad and training_flag))
    out = forward(cfg, st, xr, mr, Wr, labels, training_flag, dtype)
    (out["loss"] * grad_scale).backward()
    out.update(dx=xr.grad, dW=Wr.grad, dminput=mr.grad if mr.requires_grad else None)
    out["loss"] = out["loss"].detach()
    return out
