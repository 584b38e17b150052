"""Benchmark of the margin-softmax head: BASELINE.json metric "margin-head fwd+bwd samples/s at C=2M".

    python bench.py [--gpus N] [--steps K] [--warmup W]                    # this repo's CUDA path, cfg4 (the headline)
    python bench.py --config cfg2|cfg3|cfg1 ...                            # the other BASELINE configs, same JSON contract
    python bench.py --impl reference [--config ...] [--steps K]            # the reference's own CPU implementation

A step = one pass of the hot path over one batch: prologues (normalise W and x), fused forward
(cos-GEMM + margin + softmax-CE + top-1/5), backward (dx, dW, normalise-backward).  The headline runs the default
backward mode of the head ("stash": the training forward also writes bf16 exp2(z - ref), the backward is two GEMMs);
`alt_backward` times the "recompute" mode (north_star's formulation: nothing B x C written by the forward, logit tiles
recomputed) the same way.

Workloads (BASELINE.json configs):
  cfg4 (default)  ArcFace(s=64, m=0.5, easy_margin=False), d=512, C=2,000,000 synthetic identities, B=1024 per GPU; with
                  N>1 the class dimension is sharded over the N ranks (weak scaling: per-GPU tensor work 6*B*C*d is
                  constant) with NCCL all-gather / all-reduce / reduce-scatter.
  cfg2            CosFace / SphereFace / CurricularFace sweep, B=512, C=10,575 (value = samples over the sweep / time).
  cfg3            AdaFace / MagFace / ElasticArcFace / ElasticCosFace sweep, B=1024, C=85,742.
  cfg1            full training step: random-init ResNet-50 + ArcFace head, synthetic 112x112 faces, B=64, C=10,575
                  (img/s; the reference arm runs its CPU case, this repo's arm the same step on the GPU).

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over ranks.
`e2e` = same metric through the public nn.Module API with HOST (pinned) buffers: H2D of x and labels and
D2H of loss/accuracy inside the timed region every step.

Reference arm / cpu_baseline: when baseline/_ref holds the staged reference (baseline/stage_ref.py) the UNMODIFIED
reference head (`main_code.utils.criterion.<Head>`) + nn.CrossEntropyLoss + `metrics.accuracy` + autograd backward is
timed on the host cores at the full class count (kind "reference"); otherwise the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 512
METRIC = "margin-head fwd+bwd samples/s at C=2M"

# family -> (this package's ctor kwargs, reference class name, reference ctor kwargs); values of main_code/utils/config.py
FAMILIES = {
    "arcface": (dict(s=64.0, m=0.5, easy_margin=False), "ArcFace", dict(s=64.0, m=0.5, easy_margin=False)),
    "cosface": (dict(s=64.0, m=0.35), "CosFace", dict(s=64.0, m=0.35)),
    "sphereface": (dict(m=2), "SphereFace", dict(m=2)),
    "curricularface": (dict(m=0.5, s=64.0, momentum=0.01), "CurricularFace", dict(m=0.5, s=64.0, momentum=0.01)),
    "adaface": (dict(m=0.4, h=0.333, s=64.0, t_alpha=0.99), "AdaFace", dict(m=0.4, h=0.333, s=64.0, t_alpha=0.99)),
    "magface": (dict(s=64.0, easy_margin=False, l_margin=0.45, u_margin=0.8, l_a=10.0, u_a=110.0), "MagFace",
                dict(s=64.0, easy_margin=False, l_margin=0.45, u_margin=0.8, l_a=10.0, u_a=110.0)),
    "elastic_arc": (dict(s=64.0, m=0.5, std=0.0125, plus=False), "ElasticArcFace", dict(s=64.0, m=0.5, std=0.0125, plus=False)),
    "elastic_cos": (dict(s=64.0, m=0.35, std=0.0125, plus=False), "ElasticCosFace", dict(s=64.0, m=0.35, std=0.0125, plus=False)),
}
CONFIGS = {
    "cfg4": dict(families=["arcface"], B=1024, C=2_000_000, x_scale=1.0,
                 workload="cfg4: ArcFace(s=64,m=0.5,easy_margin=False) head fwd+bwd, d=512, C=2,000,000 synthetic "
                          "identities, B=1024 per GPU, class-sharded Partial-FC style when N>1"),
    "cfg2": dict(families=["cosface", "sphereface", "curricularface"], B=512, C=10_575, x_scale=1.0,
                 workload="cfg2: CosFace / SphereFace(m=2) / CurricularFace head fwd+bwd sweep, B=512, d=512, C=10,575 "
                          "(CASIA-WebFace classes); value = samples over the whole sweep / time"),
    "cfg3": dict(families=["adaface", "magface", "elastic_arc", "elastic_cos"], B=1024, C=85_742, x_scale=3.0,
                 workload="cfg3: AdaFace / MagFace / ElasticArcFace / ElasticCosFace norm-adaptive heads fwd+bwd sweep, "
                          "B=1024, d=512, C=85,742 (MS1MV2 scale); value = samples over the whole sweep / time"),
    "cfg1": dict(families=["arcface"], B=64, C=10_575, x_scale=1.0,
                 workload="cfg1: full training step, random-init ResNet-50 + ArcFace(s=64,m=0.5) head, synthetic 112x112 "
                          "faces, B=64, 512-d embeddings, C=10,575, SGD(momentum 0.9, wd 5e-4); img/s"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"],
                    tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe), through NVML every
    ~10 ms (the timed regions are short); falls back to an `nvidia-smi -lms 50` pipe when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, set(reasons))
        self.proc = None
        self.active = False
        self.armed = False         # set by the bench around the arms that count: timed() then switches `active` on and off
        self._stop = False
        self._t = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
            self._t = threading.Thread(target=self._poll_nvml, daemon=True)
            self._t.start()
            return
        except Exception:  # noqa: BLE001
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None
            return
        self._t = threading.Thread(target=self._pump, daemon=True)
        self._t.start()

    def _poll_nvml(self):
        n = self._nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            mx = 0.0
        while not self._stop:
            if self.active:
                try:
                    sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    pw = n.nvmlDeviceGetPowerUsage(self._h) / 1e3
                    self.samples.append((sm, mx, pw, {k for k, b in bits.items() if mask & b}))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.01)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if not self.active:
                continue
            f = [x.strip() for x in line.strip().split(",")]
            try:
                self.samples.append((float(f[1]), float(f[2]), float(f[3]),
                                     {n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")}))
            except (ValueError, IndexError):
                continue

    def stop(self):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for s_ in self.samples:
            reasons |= s_[3]
        return {"sm_mhz": statistics.median(s_[0] for s_ in self.samples), "sm_max_mhz": max(s_[1] for s_ in self.samples),
                "power_w_max": round(max(s_[2] for s_ in self.samples), 1), "reasons": sorted(reasons),
                "samples": len(self.samples), "source": "nvml" if self._nvml else "nvidia-smi"}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path on the host cores
# -------------------------------------------------------------------------------------------------
def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:  # noqa: BLE001
        return 0.0


def reference_criterion():
    """`main_code.utils.criterion` + `metrics.accuracy` of the staged, unmodified reference (baseline/_ref), or None."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        from stage_ref import reference_path
    finally:
        sys.path.pop(0)
    p = reference_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")                     # autocast(device_type='cuda') warns on a CUDA-less process
    from main_code.utils import criterion as crit
    from main_code.utils.metrics import accuracy
    return crit, accuracy


def _cpu_inputs(family, B, Cn, x_scale, seed=4):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, generator=g) * x_scale
    y = torch.randint(0, Cn, (B,), generator=g)
    return x, y, g


def cpu_head_run(family, B, Cn, steps, warmup, x_scale=1.0, budget_s=240.0):
    """Times head forward + nn.CrossEntropyLoss + accuracy + backward (criterion.py + model_utils.py:177-185) on the host
    cores, fp32, every host thread.  Returns value (samples/s), the steps actually timed, and what ran."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ours_kw, ref_cls, ref_kw = FAMILIES[family]
    x, y, g = _cpu_inputs(family, B, Cn, x_scale)
    ref = reference_criterion()
    if ref is not None:
        crit, accuracy = ref
        head = getattr(crit, ref_cls)(D, Cn, **ref_kw)          # the reference's own initialiser (values do not change the cost)
        ce = torch.nn.CrossEntropyLoss()
        kind = "reference"
        what = ("unmodified reference head modules (baseline/_ref/main_code/utils/criterion.py) + nn.CrossEntropyLoss + "
                "metrics.accuracy + autograd backward")

        def step():
            head.zero_grad(set_to_none=True)
            xg = x.clone().requires_grad_(True)
            out, _norms, loss_g, _oh = head(xg, y)
            cos_s, logits = out
            loss = ce(logits, y) + 0.0 * loss_g
            accuracy(cos_s, y, topk=(1, 5))
            loss.backward()
            return float(loss.detach())
    else:
        from oracle import margin_oracle as mo
        cfg = mo.HeadConfig.default(family)
        W = torch.randn((Cn, D) if mo.LAYOUT[family] == "CD" else (D, Cn), generator=g) * 0.01
        kind = "port"
        what = "oracle port (baseline/_ref not staged): torch CPU fp32 autograd, all B x C temporaries materialised"

        def step():
            return float(mo.autograd_step(cfg, mo.HeadState(), x, W, y, dtype=torch.float32)["loss"])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    loss = None
    while n < steps:
        loss = step()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=B * n / dt, ms_per_step=1e3 * dt / n, steps=n, warmup=warmup, cores=cores, kind=kind, loss=loss,
                sample=f"{what}; {family} B={B}, C={Cn} (full class count, no extrapolation), fp32, {cores} threads, "
                       f"{n} timed step(s) after {warmup} warm-up")


def cpu_cfg1_run(steps, warmup, budget_s=240.0):
    """BASELINE configs[0]: random-init ResNet-50 + fc->512 + the reference ArcFace + CE + SGD, one train step on CPU."""
    import torch
    import torchvision
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CONFIGS["cfg1"]
    B, Cn = cfg["B"], cfg["C"]
    torch.manual_seed(1)
    net = torchvision.models.resnet50(weights=None)                     # backbones.py:16-18 without the download
    net.fc = torch.nn.Linear(net.fc.in_features, D)
    ref = reference_criterion()
    if ref is not None:
        crit, accuracy = ref
        head = crit.ArcFace(D, Cn, s=64.0, m=0.5, easy_margin=False)
        kind = "reference"
    else:
        head, accuracy, kind = None, None, "port"
    images = torch.randn(B, 3, 112, 112)
    y = torch.randint(0, Cn, (B,))
    ce = torch.nn.CrossEntropyLoss()
    if head is None:
        from oracle import margin_oracle as mo
        Wp = torch.nn.Parameter(torch.randn(Cn, D) * 0.01)
        params = list(net.parameters()) + [Wp]
    else:
        params = list(net.parameters()) + list(head.parameters())
    opt = torch.optim.SGD(params, lr=0.1, momentum=0.9, weight_decay=5e-4)      # model_utils.py:557

    def step():
        opt.zero_grad(set_to_none=True)
        feats = net(images)
        if head is not None:
            out, _n, loss_g, _oh = head(feats, y)
            loss = ce(out[1], y)
            accuracy(out[0], y, topk=(1, 5))
        else:
            o = mo.forward_logits(mo.HeadConfig.default("arcface"), mo.HeadState(), feats, Wp, y, dtype=torch.float32)
            loss = ce(o["logits"], y)
        loss.backward()
        opt.step()
        return float(loss.detach())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=B * n / dt, ms_per_step=1e3 * dt / n, steps=n, warmup=warmup, cores=cores, kind=kind,
                sample=f"torchvision resnet50(weights=None)+fc512 + {'unmodified reference ArcFace' if kind == 'reference' else 'oracle port'} "
                       f"+ CrossEntropyLoss + SGD on CPU fp32, B={B}, C={Cn}, {cores} threads, {n} timed step(s) after {warmup} warm-up")


def reference_sweep(cfg_name, steps, warmup, budget_s):
    """The reference arm of one config: every family of the sweep at the config's own B and C (cfg4: B=32 rows, the
    reference holds ~25 fp32 B x C temporaries = 200 MB per sample at C=2M)."""
    cfg = CONFIGS[cfg_name]
    if cfg_name == "cfg1":
        r = cpu_cfg1_run(steps, warmup, budget_s)
        r["families"] = {"arcface+resnet50": {"ms_per_step": r["ms_per_step"], "steps": r["steps"]}}
        return r
    B = cfg["B"]
    note = ""
    if cfg_name == "cfg4":
        ram = host_ram_gb()
        B = 32 if ram > 40 else (16 if ram > 22 else 8)
        note = (f" [B={B} rows per step instead of 1024: the reference keeps ~25 fp32 B x C tensors alive (27.5 GB RSS at B=32, "
                f"C=2M); its W-proportional costs are amortised over fewer rows than on the GPU arm]")
    fams = cfg["families"]
    per = {}
    tot_samples = tot_s = 0.0
    last = None
    for f in fams:
        r = cpu_head_run(f, B, cfg["C"], steps, warmup, cfg["x_scale"], budget_s / len(fams))
        per[f] = {"ms_per_step": r["ms_per_step"], "steps": r["steps"], "samples_per_s": r["value"]}
        tot_samples += B * r["steps"]
        tot_s += r["ms_per_step"] * r["steps"] * 1e-3
        last = r
    steps_done = min(p["steps"] for p in per.values())
    return dict(value=tot_samples / tot_s, ms_per_step=1e3 * tot_s / sum(p["steps"] for p in per.values()), steps=steps_done,
                warmup=warmup, cores=last["cores"], kind=last["kind"], families=per,
                sample=last["sample"].replace(f"{fams[-1]} B=", f"{'/'.join(fams)} B=") + note)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    warm = min(args.warmup, 1)
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)                  # the reference's constructors print banners: keep stdout for the JSON line
    os.dup2(2, 1)
    r = reference_sweep(args.config, args.steps, warm, budget_s=float(os.environ.get("MH_REF_BUDGET_S", 150)))
    unit = "img/s" if args.config == "cfg1" else "samples/s"
    line = {
        "impl": "reference", "metric": METRIC if args.config != "cfg1" else "training-step img/s (cfg1)",
        "value": r["value"], "unit": unit,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warm, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": r["sample"], "requested_steps": args.steps,
                   "families": r["families"]},
        "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout_fd, 1)
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------
L2_FLUSH_BYTES = 256 << 20


def _parity_selfcheck(pkg, dev, world, rank, dist, torch):
    """Outside every timed region: the class-sharded head of THIS build against the chunked fp32 restatement
    (oracle/chunked_fp32.py, pinned to the reference goldens) on a problem every rank can hold - ArcFace, B=1024 per
    GPU, C=500,000 (ragged shards when world does not divide it) - loss, dx after the reduce-scatter, the shard's dW."""
    from oracle.chunked_fp32 import chunked_reference, cosine
    Cp, Bl = 500_003, 1024
    head = pkg.ShardedMarginHead("arcface", Cp, s=64.0, m=0.5, easy_margin=False).to(dev)
    g = torch.Generator(device=dev).manual_seed(1234)                      # same stream on every rank: identical full W, x, y
    Wfull = torch.randn(Cp, D, device=dev, generator=g) * 0.01
    y = torch.randint(0, Cp, (Bl * world,), device=dev, generator=g)
    x = torch.randn(Bl * world, D, device=dev, generator=g)
    near = torch.arange(Bl * world, device=dev) % 2 == 0                   # half the rows near their centre (margin branch)
    x[near] = 20.0 * torch.nn.functional.normalize(
        torch.nn.functional.normalize(Wfull[y[near]], dim=1) + torch.nn.functional.normalize(x[near], dim=1), dim=1)
    with torch.no_grad():
        head.shard_parameter().copy_(Wfull[head.c_begin:head.c_end])
    xl = x[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
    out = head.fused_loss(xl, y[rank * Bl:(rank + 1) * Bl])
    out.loss.backward()
    torch.cuda.synchronize()
    loss_ref, dx_ref, dW_ref = chunked_reference(x, Wfull, y, "arcface", 64.0, 0.5, chunk=62_500)
    res = torch.tensor([abs(float(out.loss) - float(loss_ref)) / abs(float(loss_ref)),
                        cosine(xl.grad, dx_ref[rank * Bl:(rank + 1) * Bl]),
                        cosine(head.shard_parameter().grad, dW_ref[head.c_begin:head.c_end]),
                        abs(float(xl.grad.norm()) / float(dx_ref[rank * Bl:(rank + 1) * Bl].norm()) - 1.0)],
                       device=dev, dtype=torch.float64)
    worst = res.clone()
    if world > 1:
        lo = res.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        worst[1], worst[2] = lo[1], lo[2]                                 # cosines: the worst rank is the minimum
    del head, Wfull, dx_ref, dW_ref
    torch.cuda.empty_cache()
    return {"loss_rel": float(worst[0]), "cos_dx": float(worst[1]), "cos_dw": float(worst[2]), "dx_norm_rel": float(worst[3]),
            "what": f"worst rank of {world}: ShardedMarginHead(arcface) B={Bl}/GPU C={Cp} vs chunked fp32 restatement "
                    f"(oracle/chunked_fp32.py); bar: loss_rel <= 2e-3, cos >= 0.9995",
            "pass": bool(worst[0] <= 2e-3 and worst[1] >= 0.9995 and worst[2] >= 0.9995)}


def _gpu_context(dev, B, Cn, torch):
    """The same GPU running library code, measured in this run (context for the CPU ratio, not a target): the
    reference's ArcFace module as written (its own autocast fp16 path) + CrossEntropyLoss + accuracy + backward at the
    bench shape, and the three bare cuBLAS bf16 GEMMs of a step."""
    ctx = {}

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    try:
        xh = torch.randn(B, D, device=dev).bfloat16()
        wh = torch.randn(Cn, D, device=dev).bfloat16()
        S = torch.empty(B, Cn, device=dev, dtype=torch.bfloat16)
        dxo = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
        dwo = torch.empty(Cn, D, device=dev, dtype=torch.bfloat16)
        t = [timed(lambda: torch.matmul(xh, wh.t(), out=S), 10), timed(lambda: torch.matmul(S, wh, out=dxo), 10),
             timed(lambda: torch.matmul(S.t(), xh, out=dwo), 10)]
        ctx["cublas_three_bare_gemms_ms"] = {"S=x.wT": round(t[0], 3), "dx=G.w": round(t[1], 3), "dW=GT.x": round(t[2], 3),
                                            "sum": round(sum(t), 3), "note": "torch.matmul bf16 in / bf16 out, nothing else"}
        del xh, wh, S, dxo, dwo
        torch.cuda.empty_cache()
    except Exception as ex:  # noqa: BLE001
        ctx["cublas_three_bare_gemms_ms"] = f"failed: {ex!r}"
    try:
        ref = reference_criterion()
        if ref is None:
            ctx["reference_module_on_gpu"] = "baseline/_ref not staged"
        else:
            crit, accuracy = ref
            head = crit.ArcFace(D, Cn, s=64.0, m=0.5, easy_margin=False).to(dev)
            with torch.no_grad():
                head.weight.normal_(0, 0.01)
            x = torch.randn(B, D, device=dev)
            y = torch.randint(0, Cn, (B,), device=dev)
            ce = torch.nn.CrossEntropyLoss()

            def step():
                head.zero_grad(set_to_none=True)
                xg = x.clone().requires_grad_(True)
                out, _n, _lg, _oh = head(xg, y)
                loss = ce(out[1], y)
                accuracy(out[0], y, topk=(1, 5))
                loss.backward()
            torch.cuda.reset_peak_memory_stats(dev)
            ms = timed(step, 3)
            ctx["reference_module_on_gpu"] = {"ms_per_step": round(ms, 2), "samples_per_s": round(B / ms * 1e3, 1),
                                              "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
                                              "note": "unmodified reference ArcFace (criterion.py:232-301, autocast fp16 as written) + "
                                                      "CrossEntropyLoss + accuracy + autograd, stock PyTorch on this B200"}
            del head, x, y
            torch.cuda.empty_cache()
    except Exception as ex:  # noqa: BLE001
        ctx["reference_module_on_gpu"] = f"failed: {ex!r}"
        torch.cuda.empty_cache()
    return ctx


def run_b200(args):
    import torch
    import torch.distributed as dist
    import face_recognition_models_b200 as pkg
    from face_recognition_models_b200 import _lib as L

    # Library banners (NCCL_DEBUG=VERSION prints "NCCL version ..." on stdout at communicator creation) must not
    # precede the JSON line: route fd 1 to stderr for the duration of the run, restore it to print the result.
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    if args.config == "cfg1":
        return run_b200_cfg1(args, torch, dist, pkg, dev, world, rank, saved_stdout_fd)
    Cn = args.C or cfg["C"]
    B = args.B or cfg["B"]
    peaks = load_peaks()
    small = 6.0 * Cn * D < 2.5 * (126 << 20)              # fp32 W + bf16 w^ fit in (or near) the 126 MB L2: flush between steps
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev) if small else None
    flush_sink = torch.zeros((), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events on the launching stream.  Large
        working sets: one event pair around the K steps.  Small ones: the L2 is flushed (256 MB write) before every
        step, outside that step's event pair, and the K per-step times are summed."""
        barrier()
        was_active = sampler.active if sampler else False
        if sampler and sampler.armed:
            sampler.active = True                         # clocks are sampled inside the timed regions only (not the idle gaps)
        try:
            return _timed_region(fn, steps)
        finally:
            if sampler:
                sampler.active = was_active

    def _timed_region(fn, steps):
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(steps):
                flush.zero_()
                if args.flush == "clean":
                    flush_sink.copy_(flush.sum())      # read pass: the dirty lines of the write are written back HERE, not inside the step
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                evs.append((e0, e1))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    fam_res = {}
    kern_all = {}
    tot_ms = tot_ms_e2e = tot_ms_alt = 0.0
    alt_ok = True
    n_launch = 0
    loss_val = None
    backward_desc = None
    for fam in cfg["families"]:
        ours_kw = FAMILIES[fam][0]
        if world > 1:
            head = pkg.ShardedMarginHead(fam, Cn, **ours_kw).to(dev)
            W = head.shard_parameter()
            eng = head.engine
        else:
            head = pkg.HEAD_CLASSES[fam](D, Cn, **ours_kw).to(dev)
            W = head._param()
            eng = head._engine
        g = torch.Generator(device=dev).manual_seed(4 + rank)
        with torch.no_grad():
            W.normal_(0, 0.01, generator=g)            # generated on the device per shard, never shipped through the host
        x = torch.randn(B, D, device=dev, generator=g) * cfg["x_scale"]
        y = torch.randint(0, Cn, (B,), device=dev, generator=g)
        x_host = x.cpu().pin_memory()
        y_host = y.cpu().pin_memory()
        res_host = torch.empty(3, dtype=torch.float32).pin_memory()
        x_dev, y_dev = torch.empty_like(x), torch.empty_like(y)

        def step_resident():
            xg = x.detach().requires_grad_(True)
            W.grad = None
            out = head.fused_loss(xg, y)
            out.loss.backward()
            return out

        copy_stream = torch.cuda.Stream(device=dev)
        copied = torch.cuda.Event()

        def step_e2e():
            # the W prologue does not depend on the batch: enqueue it first (head.prefetch()), and bring the batch in on a
            # copy stream beside it -- what a training loop with a prefetching data loader does; the step waits for the copy
            if world == 1:
                head.prefetch()
            with torch.cuda.stream(copy_stream):
                x_dev.copy_(x_host, non_blocking=True)
                y_dev.copy_(y_host, non_blocking=True)
                copied.record()
            torch.cuda.current_stream().wait_event(copied)
            xg = x_dev.detach().requires_grad_(True)
            W.grad = None
            out = head.fused_loss(xg, y_dev)
            out.loss.backward()
            res_host[0:1].copy_(out.loss.detach().reshape(1), non_blocking=True)
            res_host[1:2].copy_(out.acc1.reshape(1), non_blocking=True)
            res_host[2:3].copy_(out.acc5.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()        # the caller reads loss/acc every step (model_utils.py:190)
            return float(res_host[0])

        def arm(fn, steps):
            """One timed arm: idle gap, W warm-up steps, K timed steps.  Every arm (device-resident, end-to-end, the other
            backward mode) starts from the same power / thermal state: the part is power-capped, its clock under load
            settles within a few hundred milliseconds, and an arm that simply ran after another one would be measured
            in a different regime (the resident arm in the boost, the next one throttled)."""
            torch.cuda.synchronize()
            time.sleep(args.idle_s)
            for _ in range(max(args.warmup, 3)):
                fn()
            return timed(fn, steps)

        out = step_resident()
        torch.cuda.synchronize()
        loss_val = float(out.loss.detach())
        if sampler:
            sampler.armed = True
        # ---- device-resident arm -----------------------------------------------------------------------------------
        if os.environ.get("BENCH_E2E_FIRST") == "1":      # diagnostic: does the order of the arms matter on this box?
            ms_e2e = arm(step_e2e, args.steps)
            ms = arm(step_resident, args.steps)
        else:
            ms = arm(step_resident, args.steps)
            # ---- end-to-end arm (host buffers): same protocol ---------------------------------------------------------
            ms_e2e = arm(step_e2e, args.steps)
        # ---- the same K steps again with CUDA events around every C-ABI entry point on the launching stream (the
        # product path issues a step as two calls, mh_step_forward / mh_step_backward; MH_STEP_API=0 drives the same
        # kernels one entry point at a time so that each can be timed): feeds `kernels` and `roofline`, not `value`
        prev_api = os.environ.get("MH_STEP_API")
        os.environ["MH_STEP_API"] = "0"
        for _ in range(2):
            step_resident()
        L.PROFILE = []
        timed(step_resident, args.steps)
        prof, L.PROFILE = L.PROFILE, None
        if prev_api is None:
            os.environ.pop("MH_STEP_API", None)
        else:
            os.environ["MH_STEP_API"] = prev_api
        for _ in range(2):
            step_resident()
        # ---- the other backward mode (recompute: north_star's "backward recomputes logit tiles"), same protocol ----
        ms_alt = None
        # 0 recompute, 1 the proven stash, 2 the guarded stash (step API, single GPU; CurricularFace / SphereFace at scale)
        stash = 1 if eng.stash_ok() else 0
        if not stash and world == 1:
            stash = eng._stash_kind((B + 255) // 256 * 256, (eng.C + 255) // 256 * 256)
        if stash:
            eng.backward_mode = "recompute"
            ms_alt = arm(step_resident, args.steps)
            eng.backward_mode = "auto"
            for _ in range(2):
                step_resident()
        else:
            alt_ok = False
        if sampler:
            sampler.armed = False
        kern = {}
        for name, e0, e1, launches in prof:
            name = name[:-3] if name.endswith("_ex") else name        # gated forms (guarded stash): same kernels
            k = kern.setdefault(name, [0.0, 0, 0])
            k[0] += e0.elapsed_time(e1)
            k[1] += 1 if launches else 0
            k[2] += launches
        n_launch += sum(k[2] for k in kern.values())      # launches of one K-step pass (same kernels in both drivers)
        kern_all[fam] = kern
        tot_ms += ms
        tot_ms_e2e += ms_e2e
        tot_ms_alt += ms_alt or 0.0
        backward_desc = ("stash (forward writes a bf16 B x C stash; 3 GEMM passes per step)" if stash == 1 else
                         "guarded stash (speculative fixed-reference stash checked on the device; 3 GEMM passes per step)"
                         if stash == 2 else "recompute (4 GEMM passes per step)")
        fam_res[fam] = {"ms_per_step": round(ms / args.steps, 4), "e2e_ms_per_step": round(ms_e2e / args.steps, 4),
                        "backward": ("recompute", "stash", "guarded stash")[stash], "loss": loss_val,
                        **({"recompute_ms_per_step": round(ms_alt / args.steps, 4)} if ms_alt else {})}
        W_rows = W.shape[0] if eng.layout == L.LAYOUT_CD else W.shape[1]
        del head, W, eng, x, y, x_dev, y_dev
        torch.cuda.empty_cache()
    if sampler:
        sampler.stop()

    # ---- per-kernel summary (all families of the sweep pooled by entry point) ----------------------------------
    B_g = B * world
    C_loc = W_rows
    n_fam = len(cfg["families"])
    gemm_flops = 2.0 * B_g * C_loc * D                       # algorithmic FLOPs of ONE cos/dx/dw GEMM on this rank
    algo = {
        "mh_tc_forward": ("tensor", gemm_flops), "mh_tc_backward_g": ("tensor", gemm_flops),
        "mh_tc_backward_dx": ("tensor", gemm_flops), "mh_tc_backward_dw": ("tensor", gemm_flops),
        "mh_tc_backward_dw_fused": ("tensor", gemm_flops), "mh_tc_backward_dx_stash": ("tensor", gemm_flops),
        "mh_tc_backward_dw_proj": ("tensor", gemm_flops),
        "mh_tc_backward_dxdw": ("tensor", 2.0 * gemm_flops),       # both backward GEMMs in one kernel
        "mh_prologue_w": ("hbm", 6.0 * C_loc * D + 4.0 * C_loc),
        "mh_norm_backward_w": ("hbm", (4.0 + 2.0 + 4.0) * C_loc * D),
    }
    pooled = {}
    for fam, kern in kern_all.items():
        for name, (tot, calls, _l) in kern.items():
            p = pooled.setdefault(name, [0.0, 0])
            p[0] += tot
            p[1] += calls
    kernels = {}
    for name, (tot, calls) in pooled.items():
        if calls == 0:
            continue
        avg_ms = tot / calls
        ent = {"avg_ms": round(avg_ms, 4), "calls_per_step": round(calls / (args.steps * n_fam), 3)}
        if name in algo:
            kind, work = algo[name]
            ent["bound"] = kind
            ent["achieved"] = round(work / (avg_ms * 1e-3) / (1e12 if kind == "tensor" else 1e9), 2)
            ent["unit"] = "TFLOP/s" if kind == "tensor" else "GB/s"
        kernels[name] = ent
    dom = max((n for n in kernels if kernels[n].get("bound") == "tensor"), key=lambda n: kernels[n]["avg_ms"])
    # DRAM bytes of the dominant kernel per launch come from one `ncu --set full` capture of THIS shape at N=1
    # (profiles/kernel_traffic.json); the per-rank shapes at N>1 differ (B_g x C/R), so no number is claimed there.
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp) and world == 1 and args.config == "cfg4" and Cn == CONFIGS["cfg4"]["C"] and B == CONFIGS["cfg4"]["B"]:
        traffic = json.load(open(tp)).get(dom.replace("_stash", ""))
    if dom == "mh_tc_backward_dxdw":
        roofline_flops = 2.0 * gemm_flops
    else:
        roofline_flops = gemm_flops
    roofline = {
        "kernel": dom, "bound": "tensor", "achieved": kernels[dom]["achieved"], "peak": peaks["tflops_sustained"],
        "unit": "TFLOP/s", "frac": round(kernels[dom]["achieved"] / peaks["tflops_sustained"], 4), "traffic": traffic,
        "peak_source": f"{peaks['source']} sustained bf16 GEMM (kernel timed inside a long step); burst "
                       f"{peaks['tflops_burst']}",
        "frac_of_burst": round(kernels[dom]["achieved"] / peaks["tflops_burst"], 4),
        "algorithmic_flops_per_launch": roofline_flops,
    }
    steps_tot = args.steps * n_fam
    ms_per_step = tot_ms / steps_tot
    step_flops_rank = 6.0 * B_g * C_loc * D
    value = B_g * steps_tot / (tot_ms * 1e-3)
    e2e_value = B_g * steps_tot / (tot_ms_e2e * 1e-3)
    alt = None
    if alt_ok and tot_ms_alt > 0:
        alt = {"backward": "recompute (forward writes nothing of size B x C; backward recomputes the logit tiles, 4 GEMM passes)",
               "value": B_g * steps_tot / (tot_ms_alt * 1e-3), "unit": "samples/s", "ms_per_step": tot_ms_alt / steps_tot,
               "algorithmic_tflops": round(step_flops_rank / (tot_ms_alt / steps_tot * 1e-3) / 1e12, 2)}

    parity = None
    if world > 1 and not args.no_parity:
        parity = _parity_selfcheck(pkg, dev, world, rank, dist, torch)
    gpu_ctx = None
    if rank == 0 and world == 1 and args.config == "cfg4" and not args.no_gpu_context:
        gpu_ctx = _gpu_context(dev, B, Cn, torch)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                # bounded sample of the same workload: 1 warm-up + 2 timed steps per family at the full class count
                r = reference_sweep(args.config, 2, 1, budget_s=30.0)
                cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": f"failed: {ex!r}"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "B_per_gpu": B, "C": Cn, "d": D, "parallelism": f"class-shard x{world}",
                       "families": fam_res, "backward": backward_desc,
                       "l2": (f"flushed between steps (256 MB write{' + 256 MB read, so that the step does not pay the write-back of the flush buffer' if args.flush == 'clean' else ''}"
                              f" outside each step's CUDA-event pair; fp32 W + bf16 w^ = "
                              f"{6.0 * Cn * D / 2 ** 20:.0f} MB vs 126 MB L2); per-step times summed") if small else
                             "inputs_exceed_l2 (W fp32 4.1 GB + bf16 2 GB per step vs 126 MB L2)",
                       "loss": loss_val},
            "pct_of_bf16_peak": {"algorithmic_tflops": round(step_flops_rank / (ms_per_step * 1e-3) / 1e12, 2),
                                 "of_burst": round(step_flops_rank / (ms_per_step * 1e-3) / 1e12 / peaks["tflops_burst"], 4),
                                 "of_sustained": round(step_flops_rank / (ms_per_step * 1e-3) / 1e12 / peaks["tflops_sustained"], 4)},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": tot_ms_e2e / steps_tot,
                    "h2d_bytes_per_step": int(B * D * 4 + B * 8) * world, "d2h_bytes_per_step": 12 * world},
            "gpu_launches": n_launch,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "alt_backward": alt, "parity": parity,
            "gpu_context": gpu_ctx,
            "clocks": sampler.summary() if sampler else None,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def run_b200_cfg1(args, torch, dist, pkg, dev, world, rank, saved_stdout_fd):
    """BASELINE configs[0] / [4] on the GPU: the reference-style training step (autocast backbone, GradScaler, SGD;
    model_utils.py:168-192) around the fused head; DDP backbone + class-sharded head when N>1.  img/s."""
    import torchvision
    cfg = CONFIGS["cfg1"]
    B, Cn = args.B or cfg["B"], args.C or cfg["C"]
    torch.manual_seed(1 + rank)
    net = torchvision.models.resnet50(weights=None)
    net.fc = torch.nn.Linear(net.fc.in_features, D)
    net = net.to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[dev.index])
        head = pkg.ShardedMarginHead("arcface", Cn, s=64.0, m=0.5, easy_margin=False, dx_scale=world).to(dev)
    else:
        head = pkg.ArcFace(D, Cn, s=64.0, m=0.5, easy_margin=False).to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    opt_head = pkg.HeadSGD([head], lr=0.01, momentum=0.9, weight_decay=5e-4)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    images = torch.randn(B, 3, 112, 112, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, Cn, (B,), device=dev)
    img_host, y_host = images.cpu().pin_memory(), y.cpu().pin_memory()
    res_host = torch.empty(1).pin_memory()
    t_head = []

    def step(e2e=False, prof=False):
        if e2e:
            images.copy_(img_host, non_blocking=True)
            y.copy_(y_host, non_blocking=True)
        with torch.autocast("cuda"):
            feats = net(images)
        if prof:
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
        out = head.fused_loss(feats, y)
        if prof:
            h1.record()
            t_head.append((h0, h1))
        opt.zero_grad(set_to_none=True)
        opt_head.zero_grad(set_to_none=True)
        scaler.scale(out.loss).backward()
        scaler.step(opt)
        scaler.step(opt_head)
        scaler.update()
        if e2e:
            res_host.copy_(out.loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sampler = ClockSampler(dev.index) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    if sampler:
        sampler.active = True
    ms = timed(lambda: step(prof=True), args.steps)
    head_fwd_ms = sum(a.elapsed_time(b) for a, b in t_head) / len(t_head)
    for _ in range(2):
        step(e2e=True)
    ms_e2e = timed(lambda: step(e2e=True), args.steps)
    if sampler:
        sampler.active = False
        sampler.stop()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_cfg1_run(2, 1, budget_s=30.0)
                cpu = {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": "img/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex!r}"}
        line = {
            "metric": "training-step img/s (cfg1)", "value": B * world * args.steps / (ms * 1e-3), "unit": "img/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 head / fp16-autocast backbone",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "B_per_gpu": B, "C": Cn, "d": D,
                       "parallelism": f"DDP backbone x{world} + class-sharded head x{world}" if world > 1 else "single GPU",
                       "head_forward_ms": round(head_fwd_ms, 4),
                       "l2": "backbone activations exceed L2; no flush"},
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(img_host.numel() * 4 + y_host.numel() * 8) * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": None, "roofline": None, "cpu_baseline": cpu,
            "clocks": sampler.summary() if sampler else None,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--C", type=int, default=0, help="override the config's class count")
    ap.add_argument("--B", type=int, default=0, help="override the config's per-GPU batch")
    ap.add_argument("--flush", default="clean", choices=["clean", "write"],
                    help="small configs (cfg2/cfg3), L2 flush before every step: 'write' = a 256 MB write; 'clean' = the write "
                         "followed by a 256 MB read, so that the first kernels of the step do not pay the write-back of the "
                         "flush buffer's dirty lines")
    ap.add_argument("--idle-s", type=float, default=1.0, dest="idle_s",
                    help="idle gap before the warm-up of every timed arm (same power / thermal starting state for each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N>1 parity self-check (outside the timed region)")
    ap.add_argument("--no-gpu-context", action="store_true", help="skip the stock-PyTorch / cuBLAS context block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
