"""TEST INFRASTRUCTURE (checker only; never on the product path): plain PyTorch fp32 restatement of the ArcFace / CosFace
head + mean cross-entropy + gradients, evaluated in class chunks on the GPU.

The CPU oracle (oracle/margin_oracle.py) materialises B x C in fp64 and cannot hold the bench sizes (B=1024..8192,
C=2,000,000), so the full-size parity tests (tests/test_gpu_fullsize.py) and bench.py's out-of-timed-region parity
self-check for N>1 use this chunked form instead.  It follows criterion.py:262-300 (ArcFace, easy_margin=False) and
criterion.py:161-195 (CosFace; W is the class-major view of its [D, C] kernel) + nn.CrossEntropyLoss
(model_utils.py:179): fp32 GEMMs with TF32 off, fp64 softmax statistics.  It is pinned to the reference through the
small-size goldens: tests/test_oracle_golden.py::test_chunked_reference_matches_golden (CPU) compares it with
tests/golden/arcface.npz / cosface.npz (outputs of the unmodified reference).
"""
import math

import torch


def chunked_reference(x, W, y, family="arcface", s=64.0, m=0.5, chunk=125_000):
    """loss (fp64 scalar), dx [B, D], dW [C, D] (class-major) for x [B, D] fp32, W [C, D] fp32, y [B] int64."""
    assert not torch.backends.cuda.matmul.allow_tf32, "the checker needs true fp32 GEMMs"
    assert family in ("arcface", "cosface")
    B, D = x.shape
    CN = W.shape[0]
    cos_m, sin_m = math.cos(m), math.sin(m)
    th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
    xn = x.norm(dim=1, keepdim=True)
    xh = x / xn.clamp_min(1e-12)
    inv_w = 1.0 / W.norm(dim=1).clamp_min(1e-12)                       # [C]
    wy = W[y] * inv_w[y, None]
    t = (xh * wy).sum(1)                                                # target cosine
    if family == "arcface":
        sine = torch.sqrt((1.0 - t * t).clamp(0, 1))
        hard = t > th
        phi = torch.where(hard, t * cos_m - sine * sin_m, t - mm)
        dphi = torch.where(hard, cos_m + sin_m * t / sine.clamp_min(1e-12), torch.ones_like(t))
    else:
        phi = t - m                                                     # |cos| < 1 - 1e-4 assumed: the clamp is inactive
        dphi = torch.ones_like(t)
    zt = s * phi
    rows = torch.arange(B, device=x.device)
    # pass 1: log-sum-exp over all classes with the target column replaced by the margin logit
    mx = torch.full((B,), -float("inf"), dtype=torch.float64, device=x.device)
    sm = torch.zeros(B, dtype=torch.float64, device=x.device)
    for c0 in range(0, CN, chunk):
        c1 = min(CN, c0 + chunk)
        Sc = (xh @ (W[c0:c1] * inv_w[c0:c1, None]).t()) * s
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        Sd = Sc.double()
        m_new = torch.maximum(mx, Sd.max(1).values)
        sm = sm * torch.exp(mx - m_new) + torch.exp(Sd - m_new[:, None]).sum(1)
        mx = m_new
        del Sc, Sd
    lse = mx + torch.log(sm)
    loss = (lse - zt.double()).mean()
    # pass 2: gradients
    dxh = torch.zeros(B, D, dtype=torch.float32, device=x.device)
    dW = torch.empty_like(W)
    for c0 in range(0, CN, chunk):
        c1 = min(CN, c0 + chunk)
        wh = W[c0:c1] * inv_w[c0:c1, None]
        Sc = (xh @ wh.t()) * s
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        P = torch.exp(Sc.double() - lse[:, None]).float()
        del Sc
        G = P * (s / B)                                                  # dL/dcos_ij off the target
        G[rows[own], y[own] - c0] = (P[rows[own], y[own] - c0] - 1.0) * (s / B) * dphi[own]
        del P
        dxh += G @ wh
        dwh = G.t() @ xh
        dW[c0:c1] = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) * inv_w[c0:c1, None]
        del G, dwh, wh
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / xn
    return loss, dx, dW


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
