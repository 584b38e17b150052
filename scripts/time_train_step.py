"""Head training step (fused_loss + backward + optimizer step of the head parameter) at the bench workload:
torch.optim.SGD as the reference builds it (model_utils.py:557) against HeadSGD (update fused with the next W prologue).
Secondary measurement; bench.py's metric excludes the optimizer step by definition.
  python scripts/time_train_step.py [--C 2000000] [--B 1024] [--family arcface|cosface] [--steps 30]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_recognition_models_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=2_000_000)
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--family", default="arcface")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
a = ap.parse_args()


def make():
    torch.manual_seed(0)
    if a.family == "arcface":
        h = pkg.ArcFace(512, a.C, s=64.0, m=0.5, easy_margin=False).cuda()
    else:
        h = pkg.CosFace(512, a.C, s=64.0, m=0.35).cuda()
    with torch.no_grad():
        h.head_parameter().normal_(0, 0.01)
    return h


x = torch.randn(a.B, 512, device="cuda", requires_grad=True)
y = torch.randint(0, a.C, (a.B,), device="cuda")
res = {"family": a.family, "B": a.B, "C": a.C, "steps": a.steps}
for name in ("torch_sgd", "head_sgd"):
    head = make()
    if name == "torch_sgd":
        opt = torch.optim.SGD(head.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    else:
        opt = pkg.HeadSGD([head], lr=0.01, momentum=0.9, weight_decay=5e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        x.grad = None
        out = head.fused_loss(x, y)
        out.loss.backward()
        opt.step()
        return out.loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    res[name] = {"ms_per_step": round(ms, 4), "samples_per_s": round(a.B / ms * 1e3, 1), "last_loss": float(loss)}
    del head, opt
    torch.cuda.empty_cache()
print(json.dumps(res))
