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
    """oracle/chunked_fp32.py at this module's sizes (kept under its old name for the tests below)."""
    from oracle.chunked_fp32 import chunked_reference
    return chunked_reference(x, W, y, family=family, s=S, m=m, chunk=CHUNK)


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


def _near_centre_inputs(Wc, g, scale=20.0):
    """Half of the rows near their class centre (margin branch, peaked softmax), half random; labels uniform."""
    y = torch.randint(0, CN, (B,), device="cuda", generator=g)
    x = torch.randn(B, D, device="cuda", generator=g)
    near = torch.arange(B, device="cuda") % 2 == 0
    centre = torch.nn.functional.normalize(Wc[y[near]], dim=1)
    x[near] = scale * torch.nn.functional.normalize(centre + torch.nn.functional.normalize(x[near], dim=1), dim=1)
    return x, y


def test_bench_size_curricularface_matches_fp32_torch():
    """CurricularFace at C = 2,000,000 (recompute backward: its hard-negative map is the common path - at random init
    nearly every entry is 'hard' - and is not stash-eligible at s = 64): loss, dx, dW and the t buffer."""
    import face_recognition_models_b200 as pkg
    from oracle.chunked_fp32 import chunked_reference
    g = torch.Generator(device="cuda").manual_seed(8)
    head = pkg.CurricularFace(D, CN, m=0.5, s=S, momentum=0.01).cuda()
    with torch.no_grad():
        head.kernel.normal_(0, 0.01, generator=g)
        head.t.fill_(0.3)
        Wc = head.kernel.detach().t().contiguous()
    x, y = _near_centre_inputs(Wc, g)
    x.requires_grad_(True)
    out = head.fused_loss(x, y)
    out.loss.backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        loss, dx, dWc, t_new = chunked_reference(x.detach(), Wc, y, family="curricularface", s=S, m=0.5, chunk=CHUNK,
                                                 t_buf=0.3, momentum=0.01)
    lf = float(out.loss.detach())
    assert abs(lf - float(loss)) <= 2e-3 * abs(float(loss)), (lf, float(loss))
    gx, gw = x.grad, head.kernel.grad.t()
    assert _cos(gx, dx) >= 0.9995 and _cos(gw, dWc) >= 0.9995, (_cos(gx, dx), _cos(gw, dWc))
    assert abs(float(gx.norm()) / float(dx.norm()) - 1.0) <= 2e-3
    assert abs(float(gw.norm()) / float(dWc.norm()) - 1.0) <= 2e-3
    assert abs(float(head.t) - t_new) < 1e-5
    print(f"[curricularface] loss {lf:.6f} vs {float(loss):.6f}; cos dx {_cos(gx, dx):.7f} dW {_cos(gw, dWc):.7f}")


def test_bench_size_sphereface_matches_fp32_torch():
    """SphereFace (m = 2, annealed to lambda = 5) at C = 2,000,000: |x|-scaled logits through the online-max forward,
    recompute backward, the d|x| path."""
    import face_recognition_models_b200 as pkg
    from oracle.chunked_fp32 import chunked_reference
    g = torch.Generator(device="cuda").manual_seed(9)
    head = pkg.SphereFace(D, CN, m=2).cuda()
    head.iter = 20_000                                               # criterion.py:58-60: lambda = max(5, 1000 / (1 + 0.12 it))
    with torch.no_grad():
        head.weight.normal_(0, 0.01, generator=g)
    x, y = _near_centre_inputs(head.weight.detach(), g, scale=30.0)
    x.requires_grad_(True)
    out = head.fused_loss(x, y)
    out.loss.backward()
    torch.cuda.synchronize()
    assert head.lamb == 5.0
    with torch.no_grad():
        loss, dx, dW = chunked_reference(x.detach(), head.weight.detach(), y, family="sphereface", m=2, chunk=CHUNK,
                                         sphere_lambda=5.0)
    lf = float(out.loss.detach())
    assert abs(lf - float(loss)) <= 2e-3 * abs(float(loss)), (lf, float(loss))
    gx, gw = x.grad, head.weight.grad
    assert _cos(gx, dx) >= 0.9995 and _cos(gw, dW) >= 0.9995, (_cos(gx, dx), _cos(gw, dW))
    assert abs(float(gx.norm()) / float(dx.norm()) - 1.0) <= 2e-3
    assert abs(float(gw.norm()) / float(dW.norm()) - 1.0) <= 2e-3
    print(f"[sphereface] loss {lf:.6f} vs {float(loss):.6f}; cos dx {_cos(gx, dx):.7f} dW {_cos(gw, dW):.7f}")
