"""Where the end-to-end arm of bench.py spends its extra time over the device-resident arm (cfg4 shape by default).

Each variant: 1 s idle, 5 warm-up steps, 30 timed steps (CUDA events on the launching stream), repeated REPS times,
interleaved.  Variants add one ingredient of the e2e step at a time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import face_recognition_models_b200 as pkg

B, Cn = int(os.environ.get("B", 1024)), int(os.environ.get("C", 2_000_000))
K = int(os.environ.get("STEPS", 30))
head = pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False).cuda()
g = torch.Generator(device="cuda").manual_seed(4)
with torch.no_grad():
    head.weight.normal_(0, 0.01, generator=g)
x = torch.randn(B, 512, device="cuda", generator=g) * 3.0
y = torch.randint(0, Cn, (B,), device="cuda", generator=g)
x_host, y_host = x.cpu().pin_memory(), y.cpu().pin_memory()
res_host = torch.empty(3, dtype=torch.float32).pin_memory()
res_dev = torch.empty(3, dtype=torch.float32, device="cuda")
x_dev, y_dev = torch.empty_like(x), torch.empty_like(y)
copy_stream = torch.cuda.Stream()
copied = torch.cuda.Event()


def core(xx, yy):
    xg = xx.detach().requires_grad_(True)
    head.weight.grad = None
    out = head.fused_loss(xg, yy)
    out.loss.backward()
    return out


def resident():
    core(x, y)


def resident_sync():
    core(x, y)
    torch.cuda.current_stream().synchronize()


def d2h3():
    out = core(x, y)
    res_host[0:1].copy_(out.loss.detach().reshape(1), non_blocking=True)
    res_host[1:2].copy_(out.acc1.reshape(1), non_blocking=True)
    res_host[2:3].copy_(out.acc5.reshape(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()


def d2h1():
    out = core(x, y)
    torch.stack((out.loss.detach(), out.acc1, out.acc5), out=res_dev)
    res_host.copy_(res_dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def h2d_same_stream():
    x_dev.copy_(x_host, non_blocking=True)
    y_dev.copy_(y_host, non_blocking=True)
    out = core(x_dev, y_dev)
    res_host[0:1].copy_(out.loss.detach().reshape(1), non_blocking=True)
    res_host[1:2].copy_(out.acc1.reshape(1), non_blocking=True)
    res_host[2:3].copy_(out.acc5.reshape(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()


def h2d_copy_stream_prefetch():
    head.prefetch()
    with torch.cuda.stream(copy_stream):
        x_dev.copy_(x_host, non_blocking=True)
        y_dev.copy_(y_host, non_blocking=True)
        copied.record()
    torch.cuda.current_stream().wait_event(copied)
    out = core(x_dev, y_dev)
    res_host[0:1].copy_(out.loss.detach().reshape(1), non_blocking=True)
    res_host[1:2].copy_(out.acc1.reshape(1), non_blocking=True)
    res_host[2:3].copy_(out.acc5.reshape(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()


def prefetch_only_sync():
    head.prefetch()
    core(x, y)
    torch.cuda.current_stream().synchronize()


VARIANTS = [("resident (no sync)", resident), ("resident + sync per step", resident_sync),
            ("prefetch + resident + sync", prefetch_only_sync), ("+ 3 small D2H", d2h3), ("+ 1 packed D2H", d2h1),
            ("+ H2D, same stream", h2d_same_stream), ("+ H2D, copy stream + prefetch (bench e2e)", h2d_copy_stream_prefetch)]
for rep in range(int(os.environ.get("REPS", 2))):
    for name, fn in VARIANTS:
        torch.cuda.synchronize()
        time.sleep(1.0)
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(f"{name:46s} {e0.elapsed_time(e1) / K:7.3f} ms/step (events)  {1e3 * (t1 - t0) / K:7.3f} ms/step (wall)", flush=True)
