"""Context numbers on the same B200 (not part of bench.py's contract):
 (1) the three bf16 GEMMs of a head step as plain library calls (torch.matmul -> cuBLAS), no epilogue work at all;
 (2) the ArcFace head written the way the reference writes it (criterion.py:234-300: F.normalize, F.linear, the
     B x C elementwise margin passes, nn.CrossEntropyLoss) in stock PyTorch on the GPU, fp32 and bf16-autocast,
     forward + backward;
 (3) this package's fused step.
  python scripts/gpu_library_baselines.py [--C 2000000] [--B 1024]"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=2_000_000)
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
B, Cn, D = a.B, a.C, 512
dev = "cuda"


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"B": B, "C": Cn, "d": D}

# (1) library GEMMs ------------------------------------------------------------------------------
xh = torch.randn(B, D, device=dev, dtype=torch.bfloat16)
wh = torch.randn(Cn, D, device=dev, dtype=torch.bfloat16)
S = torch.empty(B, Cn, device=dev, dtype=torch.bfloat16)
dx = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
dw = torch.empty(Cn, D, device=dev, dtype=torch.bfloat16)
flop = 2.0 * B * Cn * D
t_s = timed(lambda: torch.matmul(xh, wh.t(), out=S), a.iters)
t_dx = timed(lambda: torch.matmul(S, wh, out=dx), a.iters)
t_dw = timed(lambda: torch.matmul(S.t(), xh, out=dw), a.iters)
res["library_gemms_bf16"] = {
    "S=x.wT_ms": round(t_s, 3), "dx=G.w_ms": round(t_dx, 3), "dW=GT.x_ms": round(t_dw, 3),
    "sum_ms": round(t_s + t_dx + t_dw, 3), "tflops": [round(flop / t / 1e9, 1) for t in (t_s, t_dx, t_dw)],
    "note": "bf16 in / bf16 out, no normalisation, margin, softmax or gradient epilogues"}
del S, dx, dw, xh, wh
torch.cuda.empty_cache()

# (2) the reference's formulation in stock PyTorch ---------------------------------------------------------------
s_, m_ = 64.0, 0.5
cos_m, sin_m, th, mm = math.cos(m_), math.sin(m_), math.cos(math.pi - m_), math.sin(math.pi - m_) * m_


def stock_arcface(x, W, labels):
    cosine = F.linear(F.normalize(x), F.normalize(W))
    sine = torch.sqrt((1.0 - torch.pow(cosine, 2)).clamp(0, 1))
    phi = cosine * cos_m - sine * sin_m
    phi = torch.where(cosine > th, phi, cosine - mm)
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, labels.view(-1, 1).long(), 1)
    output = (one_hot * phi) + ((1.0 - one_hot) * cosine)
    output = output * s_
    return F.cross_entropy(output, labels)


W = (torch.randn(Cn, D, device=dev) * 0.01).requires_grad_(True)
x = torch.randn(B, D, device=dev, requires_grad=True)
y = torch.randint(0, Cn, (B,), device=dev)
for name, ctx in (("stock_pytorch_fp32", torch.autocast("cuda", enabled=False)),
                  ("stock_pytorch_bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
    def step():
        W.grad = None
        x.grad = None
        with ctx:
            loss = stock_arcface(x, W, y)
        loss.backward()
        return loss
    try:
        torch.cuda.reset_peak_memory_stats()
        ms = timed(step, max(3, a.iters // 6), warm=2)
        res[name] = {"ms_per_step": round(ms, 2), "samples_per_s": round(B / ms * 1e3, 1),
                     "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    except torch.cuda.OutOfMemoryError as ex:
        res[name] = {"error": "out of memory", "detail": str(ex)[:80]}
    torch.cuda.empty_cache()
del W, x
torch.cuda.empty_cache()

# (3) this package ------------------------------------------------------------------------------------------------
import face_recognition_models_b200 as pkg  # noqa: E402

head = pkg.ArcFace(512, Cn, s=s_, m=m_, easy_margin=False).cuda()
with torch.no_grad():
    head.weight.normal_(0, 0.01)
x = torch.randn(B, D, device=dev, requires_grad=True)


def fused():
    head.weight.grad = None
    x.grad = None
    head.fused_loss(x, y).loss.backward()


ms = timed(fused, a.iters, warm=5)
res["fused_head"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1)}
print(json.dumps(res))
