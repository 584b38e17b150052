"""Per-kernel sustained time, SM clock, board power and energy per launch at the bench shape (cfg4 by default).

Every C-ABI call of one fused step is recorded (name + ctypes arguments, the buffers stay alive in the engine's
workspaces) and then replayed alone in a loop for DUR seconds while NVML is sampled, so a kernel's power state is its
own and not the step's average.  Modes: stash step, recompute step, and a no-grad forward (FWD without stash).

    B=1024 C=2000000 DUR=1.0 python scripts/kernel_power.py > gpurun_out/kernel_power.txt
"""
import os, sys, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import face_recognition_models_b200 as pkg
from face_recognition_models_b200 import _lib as L

B, Cn = int(os.environ.get("B", 1024)), int(os.environ.get("C", 2_000_000))
DUR = float(os.environ.get("DUR", 1.0))
FAM = os.environ.get("FAM", "arcface")
ONLY = [s for s in os.environ.get("ONLY", "").split(",") if s]
head = {"arcface": lambda: pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False),
        "cosface": lambda: pkg.CosFace(512, Cn, s=64.0, m=0.35),
        "curricularface": lambda: pkg.CurricularFace(512, Cn)}[FAM]().cuda()
g = torch.Generator(device="cuda").manual_seed(4)
with torch.no_grad():
    head._param().normal_(0, 0.01, generator=g)
x = torch.randn(B, 512, device="cuda", generator=g)
y = torch.randint(0, Cn, (B,), device="cuda", generator=g)


class Sampler:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.n = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0] or 0))
        self.on = False
        self.s = []
        threading.Thread(target=self.run, daemon=True).start()

    def run(self):
        n = self.n
        while True:
            if self.on:
                try:
                    self.s.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM), n.nvmlDeviceGetPowerUsage(self.h) / 1e3))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.005)


smp = Sampler()
_orig_call = L.call
rec = None


def hook(name, *args):
    if rec is not None:
        rec.append((name, args))
    _orig_call(name, *args)


L.call = hook
import face_recognition_models_b200.functional as F  # noqa: E402
F.L.call = hook
keep = []


def record(mode, grad=True):
    """One step through the per-kernel driver (MH_STEP_API=0) so that every entry point is seen by the hook."""
    global rec
    head.backward_mode = mode
    os.environ["MH_STEP_API"] = "0"
    for _ in range(2):
        xg = x.detach().requires_grad_(grad)
        head._param().grad = None
        with torch.set_grad_enabled(grad):
            out = head.fused_loss(xg, y)
        if grad:
            out.loss.backward()
    rec = []
    xg = x.detach().requires_grad_(grad)
    head._param().grad = None
    with torch.set_grad_enabled(grad):
        out = head.fused_loss(xg, y)
    if grad:
        out.loss.backward()
    keep.append((xg, out, head._param().grad))
    torch.cuda.synchronize()
    os.environ.pop("MH_STEP_API", None)
    r, rec = rec, None
    return r


def sustained(tag, name, fn):
    """Loop fn for DUR seconds with a bounded host run-ahead (20 launches), sampling NVML after 30 % of the time."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    n = 0
    smp.s = []
    prev = None
    e0.record()
    while time.time() - t0 < DUR:
        for _ in range(10):
            fn()
        n += 10
        ev = torch.cuda.Event()
        ev.record()
        if prev is not None:
            prev.synchronize()
        prev = ev
        if not smp.on and time.time() - t0 > 0.3 * DUR:
            smp.on = True
    e1.record()
    torch.cuda.synchronize()
    smp.on = False
    ms = e0.elapsed_time(e1) / n
    clk = statistics.median(s[0] for s in smp.s) if smp.s else 0
    pw = statistics.mean(s[1] for s in smp.s) if smp.s else 0
    print(f"{tag:10s} {name:28s} {ms:8.4f} ms  sm {clk:6.0f} MHz  {pw:7.1f} W  {ms * pw:8.3f} mJ/launch  (n={n}, samples={len(smp.s)})", flush=True)
    time.sleep(0.3)


def replay(tag, name, args):
    f = getattr(L.load(), name)
    sustained(tag, name, lambda: f(*args))


BIG = ("mh_prologue_w", "mh_tc_forward", "mh_tc_backward_g", "mh_tc_backward_dx", "mh_tc_backward_dx_stash",
       "mh_tc_backward_dw_fused", "mh_tc_backward_dw", "mh_tc_backward_dw_proj", "mh_tc_backward_dxdw")
for tag, mode, grad in (("stash", "auto", True), ("recompute", "recompute", True), ("nograd", "auto", False)):
    if ONLY and tag not in ONLY:
        continue
    calls = record(mode, grad)
    seen = set()
    for name, args in calls:
        if name not in BIG or name in seen:
            continue
        if name == "mh_tc_backward_dxdw":
            if not getattr(args[11], "value", None):
                continue                              # the eligibility query, no launch
        elif name.startswith("mh_tc_backward_dx") and not getattr(args[4 if name == "mh_tc_backward_dx" else 9], "value", None):
            continue                                  # the split-count query, no launch
        if tag != "stash" and name == "mh_prologue_w":
            continue
        seen.add(name)
        replay(tag, name, args)
    # the whole step, sustained
    head.backward_mode = mode

    def whole():
        xg = x.detach().requires_grad_(grad)
        head._param().grad = None
        with torch.set_grad_enabled(grad):
            out = head.fused_loss(xg, y)
        if grad:
            out.loss.backward()
    sustained(tag, "WHOLE STEP", whole)

if os.environ.get("CUBLAS", "1") == "1":
    # the same three GEMM shapes as bare library calls (bf16 in / bf16 out), for energy per GEMM
    xh = torch.randn(B, 512, device="cuda").bfloat16()
    wh = torch.randn(Cn, 512, device="cuda").bfloat16()
    S = torch.empty(B, Cn, device="cuda", dtype=torch.bfloat16)
    dxo = torch.empty(B, 512, device="cuda", dtype=torch.bfloat16)
    dwo = torch.empty(Cn, 512, device="cuda", dtype=torch.bfloat16)
    sustained("cublas", "S = x w^T", lambda: torch.matmul(xh, wh.t(), out=S))
    sustained("cublas", "dx = G w", lambda: torch.matmul(S, wh, out=dxo))
    sustained("cublas", "dW = G^T x", lambda: torch.matmul(S.t(), xh, out=dwo))
