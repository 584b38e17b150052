"""GPU: the bench workload itself (BASELINE cfg4: ArcFace s=64 m=0.5, B=1024, d=512, C=2,000,000) against a plain
PyTorch fp32 restatement of criterion.py:262-300 + nn.CrossEntropyLoss, evaluated in class chunks on the same GPU
(the CPU oracle cannot hold B x C at this size).  Tolerances are north_star's bf16 bar: loss 2e-3 relative,
gradient norms 2e-3 relative, gradient cosine >= 0.9995."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

B, CN, D = 1024, 2_000_000, 512
S, M = 64.0, 0.5
CHUNK = 125_000


def _fp32_chunked_reference(x, W, y, family="arcface", m=M):
    """loss, dx, dW of ArcFace(easy_margin=False) (criterion.py:262-300) or CosFace (criterion.py:161-195; W is the
    class-major view of its [D, C] kernel) + mean cross-entropy; fp32 GEMMs (TF32 off), fp64 softmax statistics."""
    assert not torch.backends.cuda.matmul.allow_tf32
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
        phi = t - m                                                     # |cos| < 1 here: the reference's clamp is inactive
        dphi = torch.ones_like(t)
    zt = S * phi
    rows = torch.arange(B, device=x.device)
    # pass 1: log-sum-exp over all classes with the target column replaced by the margin logit
    mx = torch.full((B,), -float("inf"), dtype=torch.float64, device=x.device)
    sm = torch.zeros(B, dtype=torch.float64, device=x.device)
    for c0 in range(0, CN, CHUNK):
        c1 = min(CN, c0 + CHUNK)
        Sc = (xh @ (W[c0:c1] * inv_w[c0:c1, None]).t()) * S
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        Sd = Sc.double()
        m_new = torch.maximum(mx, Sd.max(1).values)
        sm = sm * torch.exp(mx - m_new) + torch.exp(Sd - m_new[:, None]).sum(1)
        mx = m_new
    lse = mx + torch.log(sm)
    loss = (lse - zt.double()).mean()
    # pass 2: gradients
    dxh = torch.zeros(B, D, dtype=torch.float32, device=x.device)
    dW = torch.empty_like(W)
    for c0 in range(0, CN, CHUNK):
        c1 = min(CN, c0 + CHUNK)
        wh = W[c0:c1] * inv_w[c0:c1, None]
        Sc = (xh @ wh.t()) * S
        own = (y >= c0) & (y < c1)
        Sc[rows[own], y[own] - c0] = zt[own]
        P = torch.exp(Sc.double() - lse[:, None]).float()
        G = P * (S / B)                                                  # dL/dcos_ij off the target
        G[rows[own], y[own] - c0] = (P[rows[own], y[own] - c0] - 1.0) * (S / B) * dphi[own]
        dxh += G @ wh
        dwh = G.t() @ xh
        dW[c0:c1] = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) * inv_w[c0:c1, None]
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / xn
    return loss, dx, dW


def test_bench_size_cosface_dc_layout_matches_fp32_torch():
    """Same size through the [D, C] parameter layout (CosFace kernel): DC W prologue, stash backward, DC dW epilogue."""
    import face_recognition_models_b200 as pkg
    g = torch.Generator(device="cuda").manual_seed(6)
    head = pkg.CosFace(D, CN, s=S, m=0.35).cuda()
    with torch.no_grad():
        head.kernel.normal_(0, 0.01, generator=g)
    y = torch.randint(0, CN, (B,), device="cuda", generator=g)
    x = torch.randn(B, D, device="cuda", generator=g)
    with torch.no_grad():
        near = torch.arange(B, device="cuda") % 2 == 0
        centre = torch.nn.functional.normalize(head.kernel[:, y[near]].t(), dim=1)
        x[near] = 20.0 * torch.nn.functional.normalize(centre + torch.nn.functional.normalize(x[near], dim=1), dim=1)
    x.requires_grad_(True)
    out = head.fused_loss(x, y)
    out.loss.backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        Wc = head.kernel.detach().t().contiguous()                     # class-major copy for the restatement
        loss, dx, dWc = _fp32_chunked_reference(x.detach(), Wc, y, family="cosface", m=0.35)
        del Wc
    lf = float(out.loss.detach())
    assert abs(lf - float(loss)) <= 2e-3 * abs(float(loss)), (lf, float(loss))
    gx, gw = x.grad, head.kernel.grad.t()
    assert _cos(gx, dx) >= 0.9995 and _cos(gw, dWc) >= 0.9995, (_cos(gx, dx), _cos(gw, dWc))
    assert abs(float(gx.norm()) / float(dx.norm()) - 1.0) <= 2e-3
    assert abs(float(gw.norm()) / float(dWc.norm()) - 1.0) <= 2e-3
    print(f"[cosface DC] loss {lf:.6f} vs {float(loss):.6f}; cos dx {_cos(gx, dx):.7f} dW {_cos(gw, dWc):.7f}")


def _cos(a, b):
    return float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()))


@pytest.mark.parametrize("bmode", ["auto", "recompute"])
def test_bench_workload_matches_fp32_torch(bmode):
    import face_recognition_models_b200 as pkg
    g = torch.Generator(device="cuda").manual_seed(4)
    head = pkg.ArcFace(D, CN, s=S, m=M, easy_margin=False).cuda()
    head.backward_mode = bmode
    with torch.no_grad():
        head.weight.normal_(0, 0.01, generator=g)
    # half of the rows sit near their class centre (t ~ 0.7: the margin branch and a peaked softmax), half are random
    y = torch.randint(0, CN, (B,), device="cuda", generator=g)
    x = torch.randn(B, D, device="cuda", generator=g)
    with torch.no_grad():
        near = torch.arange(B, device="cuda") % 2 == 0
        centre = torch.nn.functional.normalize(head.weight[y[near]], dim=1)
        noise = torch.nn.functional.normalize(x[near], dim=1)
        x[near] = 20.0 * torch.nn.functional.normalize(centre + noise, dim=1)
    x.requires_grad_(True)
    out = head.fused_loss(x, y)
    out.loss.backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        loss, dx, dW = _fp32_chunked_reference(x.detach(), head.weight.detach(), y)
    lf = float(out.loss.detach())
    assert abs(lf - float(loss)) <= 2e-3 * abs(float(loss)), (lf, float(loss))
    gx, gw = x.grad, head.weight.grad
    assert _cos(gx, dx) >= 0.9995 and _cos(gw, dW) >= 0.9995, (_cos(gx, dx), _cos(gw, dW))
    assert abs(float(gx.norm()) / float(dx.norm()) - 1.0) <= 2e-3
    assert abs(float(gw.norm()) / float(dW.norm()) - 1.0) <= 2e-3
    # accuracy: the near rows are classified correctly, the random rows are not
    assert 45.0 <= float(out.acc1) <= 55.0, float(out.acc1)
    print(f"[{bmode}] loss {lf:.6f} vs {float(loss):.6f}; cos dx {_cos(gx, dx):.7f} dW {_cos(gw, dW):.7f}; "
          f"norm ratio dx {float(gx.norm()) / float(dx.norm()):.6f} dW {float(gw.norm()) / float(dW.norm()):.6f}")
