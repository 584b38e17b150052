"""Staged GPU diagnostics (each stage in its own process so one CUDA fault does not mask the rest).

    python scripts/gpu_diag.py            # run every stage, print a summary
    python scripts/gpu_diag.py <stage>    # run one stage in-process
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def _imports():
    import face_recognition_models_b200 as pkg
    from face_recognition_models_b200 import _lib as L
    from face_recognition_models_b200.functional import _ptr, _stream
    from oracle import margin_oracle as mo
    return pkg, L, _ptr, _stream, mo


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosim(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0))


def stage_prologue():
    pkg, L, _ptr, _stream, mo = _imports()
    dev = "cuda"
    for layout, name in ((L.LAYOUT_CD, "CD"), (L.LAYOUT_DC, "DC")):
        for Cn in (61, 1000, 4097):
            Wc = torch.randn(Cn, 512, device=dev) * 0.05
            W = Wc.contiguous() if layout == L.LAYOUT_CD else Wc.t().contiguous()
            C_pad = (Cn + 255) // 256 * 256
            wh = torch.full((C_pad, 512), 7.0, dtype=torch.bfloat16, device=dev)
            wh32 = torch.empty(Cn, 512, device=dev)
            inv = torch.empty(Cn, device=dev)
            L.call("mh_prologue_w", _ptr(W), layout, Cn, W.shape[1], _ptr(wh), C_pad, _ptr(wh32), _ptr(inv), _stream())
            torch.cuda.synchronize()
            ref = torch.nn.functional.normalize(Wc, dim=1)
            print(f"prologue_w {name} C={Cn}: w32 rel={rel(wh32, ref):.2e} bf16 rel={rel(wh[:Cn].float(), ref):.2e} "
                  f"inv rel={rel(inv, 1 / Wc.norm(dim=1)):.2e} pad_zero={bool((wh[Cn:] == 0).all())}")
            assert rel(wh32, ref) < 1e-6 and rel(wh[:Cn].float(), ref) < 5e-3 and bool((wh[Cn:] == 0).all())


def stage_sgemm():
    pkg, L, _ptr, _stream, mo = _imports()
    dev = "cuda"
    M, N, K = 70, 133, 512
    A = torch.randn(M, K, device=dev)
    Bm = torch.randn(K, N, device=dev)
    Cm = torch.empty(M, N, device=dev)
    L.call("mh_sgemm_strided", M, N, K, _ptr(A), K, 1, _ptr(Bm), N, 1, _ptr(Cm), N, _stream())
    print("sgemm NN rel", rel(Cm, A.double() @ Bm.double()))
    Bt = Bm.t().contiguous()  # [N,K]
    L.call("mh_sgemm_strided", M, N, K, _ptr(A), K, 1, _ptr(Bt), 1, K, _ptr(Cm), N, _stream())
    print("sgemm NT rel", rel(Cm, A.double() @ Bm.double()))
    At = A.t().contiguous()  # [K,M]
    L.call("mh_sgemm_strided", M, N, K, _ptr(At), 1, M, _ptr(Bm), N, 1, _ptr(Cm), N, _stream())
    r = rel(Cm, A.double() @ Bm.double())
    print("sgemm TN rel", r)
    assert r < 1e-5


def _run_family(pkg, mo, fam, B, Cn, seed, mode, lambda_g=0.0, state=None, grad_scale=1.0, **kw):
    cfg = mo.HeadConfig.default(fam)
    for k, v in kw.items():
        setattr(cfg, k, v)
    state = state or mo.HeadState()
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed)
    margins = None
    if fam.startswith("elastic"):
        torch.manual_seed(99)
        margins = mo.sample_elastic_margins(cfg, B)
    ref = mo.loss_and_grads(cfg, state, x, W, labels, margins=margins, lambda_g=lambda_g, grad_scale=grad_scale)
    head = build_head(pkg, fam, cfg, Cn).cuda()
    head.mode = mode
    with torch.no_grad():
        head._param().copy_(W.cuda())
    if fam == "sphereface":
        head.iter = state.sphere_iter
    if fam == "curricularface":
        head.t.fill_(state.t_buf)
    if fam == "adaface":
        head.batch_mean.fill_(state.batch_mean)
        head.batch_std.fill_(state.batch_std)
    if margins is not None:
        head._margins_override = margins.cuda()
    xg = x.cuda().requires_grad_(True)
    out = head.fused_loss(xg, labels.cuda())
    loss = out.loss + lambda_g * out.loss_g
    (loss * grad_scale).backward()
    torch.cuda.synchronize()
    res = dict(
        loss=abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"])),
        dx=rel(xg.grad, ref["dx"]), dW=rel(head._param().grad, ref["dW"]),
        cdx=cosim(xg.grad, ref["dx"]), cdW=cosim(head._param().grad, ref["dW"]),
        acc1=abs(float(out.acc1) - float(ref["acc1"])), acc5=abs(float(out.acc5) - float(ref["acc5"])),
        norms=rel(out.norms.flatten(), ref["norms"]),
    )
    return res, head, ref


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


def stage_exact():
    pkg, L, _ptr, _stream, mo = _imports()
    bad = 0
    for fam in mo.FAMILIES:
        for (B, Cn) in ((8, 61), (64, 1000)):
            lg = 35.0 if fam == "magface" else 0.0
            res, _, _ = _run_family(pkg, mo, fam, B, Cn, seed=100 + B, mode="exact", lambda_g=lg)
            ok = res["loss"] < 1e-4 and res["cdx"] > 0.9999 and res["cdW"] > 0.9999 and res["acc1"] < 1e-3
            bad += 0 if ok else 1
            print(f"exact {fam:15s} B={B:3d} C={Cn:5d} " + " ".join(f"{k}={v:.2e}" for k, v in res.items()) +
                  ("" if ok else "   <-- FAIL"))
    assert bad == 0, f"{bad} exact-mode cases failed"


def stage_tcgemm():
    """Descriptor validation: the two plain tensor-core GEMMs against torch.matmul."""
    pkg, L, _ptr, _stream, mo = _imports()
    dev = "cuda"
    torch.manual_seed(0)
    for (B_pad, C_pad) in ((256, 256), (256, 1024), (1024, 4096)):
        G = (torch.randn(B_pad, C_pad, device=dev) * 0.5).to(torch.bfloat16)
        wh = (torch.randn(C_pad, 512, device=dev) * 0.1).to(torch.bfloat16)
        xh = (torch.randn(B_pad, 512, device=dev) * 0.1).to(torch.bfloat16)
        Gl = G                                                       # logical [B_pad, C_pad]
        G = G.view(B_pad, C_pad // 128, 128).permute(1, 0, 2).contiguous()   # class-tiled storage the kernels expect
        ns = C.c_int(0)
        L.call("mh_tc_backward_dx", _ptr(G), B_pad, C_pad, _ptr(wh), C.c_void_p(0), C.byref(ns), C.c_void_p(0), _stream())
        part = torch.zeros(ns.value, B_pad, 512, device=dev)
        L.call("mh_tc_backward_dx", _ptr(G), B_pad, C_pad, _ptr(wh), _ptr(part), C.byref(ns), C.c_void_p(0), _stream())
        torch.cuda.synchronize()
        got = part.sum(0)
        ref = Gl.double() @ wh.double()
        print(f"tc dx  B_pad={B_pad} C_pad={C_pad} n_split={ns.value} rel={rel(got, ref):.3e}")
        dw = torch.zeros(C_pad, 512, device=dev)
        L.call("mh_tc_backward_dw", _ptr(G), B_pad, C_pad, _ptr(xh), _ptr(dw), _stream())
        torch.cuda.synchronize()
        refw = Gl.double().t() @ xh.double()
        print(f"tc dw  B_pad={B_pad} C_pad={C_pad} rel={rel(dw, refw):.3e}")
        assert rel(got, ref) < 1e-3 and rel(dw, refw) < 1e-3


def stage_tcfwd():
    pkg, L, _ptr, _stream, mo = _imports()
    bad = 0
    for fam in ("arcface", "cosface", "curricularface", "mv_am", "sphereface"):
        for (B, Cn) in ((8, 61), (64, 1000), (200, 3000)):
            res, _, _ = _run_family(pkg, mo, fam, B, Cn, seed=300 + B, mode="tc")
            ok = res["loss"] < 2e-3 and res["cdx"] > 0.9995 and res["cdW"] > 0.9995
            bad += 0 if ok else 1
            print(f"tc {fam:15s} B={B:3d} C={Cn:5d} " + " ".join(f"{k}={v:.2e}" for k, v in res.items()) +
                  ("" if ok else "   <-- FAIL"))
    assert bad == 0


def stage_tcall():
    pkg, L, _ptr, _stream, mo = _imports()
    bad = 0
    for fam in mo.FAMILIES:
        lg = 35.0 if fam == "magface" else 0.0
        res, _, _ = _run_family(pkg, mo, fam, 512, 10575, seed=500, mode="tc", lambda_g=lg)
        ok = res["loss"] < 2e-3 and res["cdx"] > 0.9995 and res["cdW"] > 0.9995
        bad += 0 if ok else 1
        print(f"tc {fam:15s} B=512 C=10575 " + " ".join(f"{k}={v:.2e}" for k, v in res.items()) +
              ("" if ok else "   <-- FAIL"))
    assert bad == 0


def stage_time():
    """Per-kernel CUDA-event timing of the ArcFace tc path at growing C."""
    pkg, L, _ptr, _stream, mo = _imports()
    dev = "cuda"
    for Cn in (100_000, 2_000_000):
        B = 1024
        head = pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False)
        head = head.cuda()
        with torch.no_grad():
            head.weight.normal_(0, 0.01)
        x = torch.randn(B, 512, device=dev, requires_grad=True)
        labels = torch.randint(0, Cn, (B,), device=dev)
        for it in range(3):
            out = head.fused_loss(x, labels)
            out.loss.backward()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        out = head.fused_loss(x, labels)
        ev[1].record()
        out.loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        flops = 6.0 * B * Cn * 512
        print(f"time C={Cn}: fwd {f:.3f} ms bwd {b:.3f} ms total {f + b:.3f} ms -> "
              f"{flops / (f + b) / 1e9:.1f} TFLOP/s algorithmic, {B / (f + b) * 1e3:.0f} samples/s, loss={float(out.loss):.4f}")
        del head, x
        torch.cuda.empty_cache()


STAGES = dict(prologue=stage_prologue, sgemm=stage_sgemm, exact=stage_exact, tcgemm=stage_tcgemm,
              tcfwd=stage_tcfwd, tcall=stage_tcall, time=stage_time)


def main():
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]]()
        print(f"[stage {sys.argv[1]}] OK")
        return
    summary = {}
    for name in STAGES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=300)
            out = p.stdout + p.stderr
            rc = p.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            out += "\nTIMEOUT"
            rc = -9
        print(f"===== stage {name}: rc={rc} ({time.time() - t0:.1f}s) =====")
        print(out[-6000:])
        summary[name] = rc
    print("SUMMARY", summary)


if __name__ == "__main__":
    main()
