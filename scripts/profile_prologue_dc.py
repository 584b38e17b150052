"""One DC-layout W prologue (CosFace kernel [512, C]) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_recognition_models_b200 as pkg  # noqa: E402

Cn = int(os.environ.get("C", 2_000_000))
head = pkg.CosFace(512, Cn, s=64.0, m=0.35).cuda()
x = torch.randn(256, 512, device="cuda")
y = torch.randint(0, Cn, (256,), device="cuda")
with torch.no_grad():
    for _ in range(3):
        out = head.fused_loss(x, y)
torch.cuda.synchronize()
print("loss", float(out.loss))
