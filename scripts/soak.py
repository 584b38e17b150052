"""Soak run: many back-to-back fused steps per family / backward mode; checks that nothing hangs or drifts (same inputs ->
same loss every step) and prints the sustained step time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import face_recognition_models_b200 as pkg

B, Cn, STEPS = 1024, int(os.environ.get("C", 500_000)), int(os.environ.get("STEPS", 400))
for fam, ctor, bmode in (("arcface", lambda: pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False), "auto"),
                         ("arcface", lambda: pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False), "recompute"),
                         ("cosface", lambda: pkg.CosFace(512, Cn), "auto"),
                         ("magface", lambda: pkg.MagFace(512, Cn), "auto"),
                         ("curricularface", lambda: pkg.CurricularFace(512, Cn), "auto"),
                         ("sphereface", lambda: pkg.SphereFace(512, Cn, m=2), "auto"),
                         ("vpl_arcface", lambda: pkg.VPLArcFace(512, Cn), "auto")):
    head = ctor().cuda()
    head.backward_mode = bmode
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        head._param().normal_(0, 0.01, generator=g)
    x = (torch.randn(B, 512, device="cuda", generator=g) * 3).requires_grad_(True)
    y = torch.randint(0, Cn, (B,), device="cuda", generator=g)
    losses = []
    torch.cuda.synchronize(); t0 = time.time()
    for i in range(STEPS):
        x.grad = None; head._param().grad = None
        out = head.fused_loss(x, y)
        out.loss.backward()
        if i % 50 == 0:
            losses.append(float(out.loss.detach()))
    torch.cuda.synchronize(); dt = time.time() - t0
    ok = all(torch.isfinite(torch.tensor(losses))) and bool(torch.isfinite(x.grad).all()) and bool(torch.isfinite(head._param().grad).all())
    stateless = fam in ("arcface", "cosface", "magface")
    same = (max(losses) - min(losses)) < 1e-6 * abs(losses[0]) if stateless else True
    print(f"{fam:15s} {bmode:9s} {STEPS} steps  {1e3 * dt / STEPS:.3f} ms/step  loss {losses[0]:.4f} .. {losses[-1]:.4f}  finite={ok} constant={same}", flush=True)
    assert ok and same
    del head
    torch.cuda.empty_cache()
print("soak ok")
