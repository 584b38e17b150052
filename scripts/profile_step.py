"""One fused head step for ncu (profiling range limited to the last step); --family selects the head (default ArcFace)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_recognition_models_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=2_000_000)
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--family", default="arcface", choices=["arcface", "curricularface", "sphereface", "cosface"])
a = ap.parse_args()
head = {"arcface": lambda: pkg.ArcFace(512, a.C, s=64.0, m=0.5, easy_margin=False),
        "curricularface": lambda: pkg.CurricularFace(512, a.C),
        "sphereface": lambda: pkg.SphereFace(512, a.C, m=2),
        "cosface": lambda: pkg.CosFace(512, a.C, s=64.0, m=0.35)}[a.family]().cuda()
with torch.no_grad():
    head._param().normal_(0, 0.01)
x = torch.randn(a.B, 512, device="cuda", requires_grad=True)
y = torch.randint(0, a.C, (a.B,), device="cuda")
def zero():                         # optimizer.zero_grad(set_to_none=True): no gradient-accumulation kernels
    x.grad = None
    head._param().grad = None


for _ in range(a.warmup):
    zero()
    head.fused_loss(x, y).loss.backward()
zero()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = head.fused_loss(x, y)
out.loss.backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(out.loss))
