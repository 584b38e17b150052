"""TEST INFRASTRUCTURE (checker only; never on the product path): plain PyTorch fp32 restatement of the ArcFace / CosFace /
CurricularFace / SphereFace heads + mean cross-entropy + gradients, evaluated in class chunks on the GPU.

The CPU oracle (oracle/margin_oracle.py) materialises B x C in fp64 and cannot hold the bench sizes (B=1024..8192,
C=2,000,000), so the full-size parity tests (tests/test_gpu_fullsize.py) and bench.py's out-of-timed-region parity
self-check for N>1 use this chunked form instead.  It follows criterion.py:262-300 (ArcFace, easy_margin=False) and
criterion.py:161-195 (CosFace; W is the class-major view of its [D, C] kernel), criterion.py:527-587 (CurricularFace,
hard-negative re-weighting with the updated t buffer) and criterion.py:57-107 (SphereFace, |x|-scaled logits and the
d|x| path) + nn.CrossEntropyLoss
(model_utils.py:179): fp32 GEMMs with TF32 off, fp64 softmax statistics.  It is pinned to the reference through the
small-size goldens: tests/test_oracle_golden.py::test_chunked_reference_matches_golden (CPU) compares it with
tests/golden/arcface.npz / cosface.npz / curricularface*.npz / sphereface_m*.npz (outputs of the unmodified reference).
"""
import math

import torch


def _cheb(m, c):
    """Chebyshev T_m(c) and its derivative, m in 0..5 (criterion.py:40-47)."""
    return {0: (torch.ones_like(c), torch.zeros_like(c)), 1: (c, torch.ones_like(c)), 2: (2 * c ** 2 - 1, 4 * c),
            3: (4 * c ** 3 - 3 * c, 12 * c ** 2 - 3), 4: (8 * c ** 4 - 8 * c ** 2 + 1, 32 * c ** 3 - 16 * c),
            5: (16 * c ** 5 - 20 * c ** 3 + 5 * c, 80 * c ** 4 - 60 * c ** 2 + 5)}[int(m)]


def chunked_reference(x, W, y, family="arcface", s=64.0, m=0.5, chunk=125_000, t_buf=0.0, momentum=0.01, sphere_lambda=5.0):
    """loss (fp64 scalar), dx [B, D], dW [C, D] (class-major) for x [B, D] fp32, W [C, D] fp32, y [B] int64.

    family: arcface (easy_margin=False, criterion.py:262-300), cosface (161-195), curricularface (527-587; t_buf is the
    buffer BEFORE this step, the hard negatives use the updated value as the reference does) or sphereface (57-107; m is
    the integer margin, sphere_lambda the annealing value of this step).  Returns the updated t_buf as a 4th value for
    curricularface."""
    assert not torch.backends.cuda.matmul.allow_tf32, "the checker needs true fp32 GEMMs"
    assert family in ("arcface", "cosface", "curricularface", "sphereface")
    B, D = x.shape
    CN = W.shape[0]
    xn = x.norm(dim=1, keepdim=True)
    xh = x / xn.clamp_min(1e-12)
    inv_w = 1.0 / W.norm(dim=1).clamp_min(1e-12)                       # [C]
    wy = W[y] * inv_w[y, None]
    t = (xh * wy).sum(1)                                                # target cosine
    lo, hi = {"arcface": (-2.0, 2.0), "cosface": (-1 + 1e-4, 1 - 1e-4), "curricularface": (-1.0, 1.0), "sphereface": (-1.0, 1.0)}[family]
    scale = torch.full_like(t, s)                                       # per-row logit scale
    thr = None                                                          # hard-negative threshold (curricularface)
    tb_new = None
    if family == "arcface":
        cos_m, sin_m = math.cos(m), math.sin(m)
        th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
        sine = torch.sqrt((1.0 - t * t).clamp(0, 1))
        hard = t > th
        phi = torch.where(hard, t * cos_m - sine * sin_m, t - mm)
        dphi = torch.where(hard, cos_m + sin_m * t / sine.clamp_min(1e-12), torch.ones_like(t))
    elif family == "cosface":
        phi = t.clamp(lo, hi) - m
        dphi = ((t >= lo) & (t <= hi)).float()
    elif family == "curricularface":
        cos_m, sin_m = math.cos(m), math.sin(m)
        th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
        inside_t = ((t >= lo) & (t <= hi)).float()
        t = t.clamp(lo, hi)
        sin_t = torch.sqrt(1.0 - t * t)
        ctm = t * cos_m - sin_t * sin_m
        take = t > th
        phi = torch.where(take, ctm, t - mm)
        dphi = torch.where(take, cos_m + t / sin_t * sin_m, torch.ones_like(t)) * inside_t
        tb_new = float(t.double().mean()) * momentum + (1 - momentum) * float(t_buf)      # criterion.py:570-573
        thr = ctm
    else:                                                               # sphereface
        inside_t = ((t >= lo) & (t <= hi)).float()
        t = t.clamp(lo, hi)
        theta = torch.acos(t)
        k = torch.floor(float(m) * theta / math.pi)
        sign = 1.0 - 2.0 * torch.remainder(k, 2.0)
        Tm, dTm = _cheb(m, t)
        phi_s = sign * Tm - 2.0 * k
        phi = (phi_s - t) / (1.0 + sphere_lambda) + t                   # u at the target column
        dphi = ((sign * dTm - 1.0) / (1.0 + sphere_lambda) + 1.0) * inside_t
        scale = xn.reshape(-1)
    zt = scale * phi

    def tile(c0, c1):
        """u = z / scale and du/dcos_raw on the class block [c0, c1) (target column still untreated)."""
        wh = W[c0:c1] * inv_w[c0:c1, None]
        raw = xh @ wh.t()
        c = raw if family == "arcface" else raw.clamp(lo, hi)
        du = torch.ones_like(raw) if family == "arcface" else ((raw >= lo) & (raw <= hi)).float()
        u = c
        if family == "curricularface":
            hardn = c > thr[:, None]
            u = torch.where(hardn, c * (tb_new + c), c)                 # criterion.py:559, 575
            du = du * torch.where(hardn, tb_new + 2.0 * c, torch.ones_like(c))
        return wh, u, du

    rows = torch.arange(B, device=x.device)
    # pass 1: log-sum-exp over all classes with the target column replaced by the margin logit
    mx = torch.full((B,), -float("inf"), dtype=torch.float64, device=x.device)
    sm = torch.zeros(B, dtype=torch.float64, device=x.device)
    for c0 in range(0, CN, chunk):
        c1 = min(CN, c0 + chunk)
        _wh, u, _du = tile(c0, c1)
        Sc = u * scale[:, None]
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        Sd = Sc.double()
        m_new = torch.maximum(mx, Sd.max(1).values)
        sm = sm * torch.exp(mx - m_new) + torch.exp(Sd - m_new[:, None]).sum(1)
        mx = m_new
        del Sc, Sd, u, _du, _wh
    lse = mx + torch.log(sm)
    loss = (lse - zt.double()).mean()
    # pass 2: gradients
    dxh = torch.zeros(B, D, dtype=torch.float32, device=x.device)
    dn = torch.zeros(B, dtype=torch.float32, device=x.device)           # d loss / d |x_i| (sphereface: z = u |x|)
    dW = torch.empty_like(W)
    for c0 in range(0, CN, chunk):
        c1 = min(CN, c0 + chunk)
        wh, u, du = tile(c0, c1)
        Sc = u * scale[:, None]
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        P = torch.exp(Sc.double() - lse[:, None]).float()
        del Sc
        Gz = P / B                                                       # d loss / d z
        Gz[rows[own], y[own] - c0] -= 1.0 / B
        del P
        if family == "sphereface":
            u[rows[own], y[own] - c0] = phi[own]
            dn += (Gz * u).sum(1)
        G = Gz * scale[:, None] * du                                     # d loss / d cos_raw off the target
        G[rows[own], y[own] - c0] = Gz[rows[own], y[own] - c0] * scale[own] * dphi[own]
        del Gz, u, du
        dxh += G @ wh
        dwh = G.t() @ xh
        dW[c0:c1] = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) * inv_w[c0:c1, None]
        del G, dwh, wh
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / xn + dn[:, None] * xh
    if family == "curricularface":
        return loss, dx, dW, tb_new
    return loss, dx, dW


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-300))
