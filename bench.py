"""Benchmark of the margin-softmax head: BASELINE.json metric "margin-head fwd+bwd samples/s at C=2M".

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPUs

A step = one pass of the hot path over one batch: prologues (normalise W and x), fused forward
(cos-GEMM + margin + softmax-CE + top-1/5), backward (dx, dW, normalise-backward).  The headline runs the default
backward mode of the head ("stash": the training forward also writes bf16 exp2(z - ref), the backward is two GEMMs);
`alt_backward` times the "recompute" mode (nothing B x C written by the forward, logit tiles recomputed) the same way.
Workload (BASELINE configs[3]): ArcFace(s=64, m=0.5, easy_margin=False), d=512, C=2,000,000 synthetic
identities, B=1024 per GPU; with N>1 the class dimension is sharded over the N ranks (weak scaling:
per-GPU tensor work 6*B*C*d is constant) with NCCL all-gather / all-reduce / reduce-scatter.

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over ranks.
`e2e` = same metric through the public nn.Module API with HOST (pinned) buffers: H2D of x and labels and
D2H of loss/accuracy inside the timed region every step.
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
C_TOTAL = 2_000_000
B_PER_GPU = 1024
WORKLOAD = ("cfg4: ArcFace(s=64,m=0.5,easy_margin=False) head fwd+bwd, d=512, C=2,000,000 synthetic identities, "
            "B=1024 per GPU, class-sharded Partial-FC style when N>1")


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
# reference / CPU-baseline arm: the oracle's materialising autograd formulation on the host cores
# -------------------------------------------------------------------------------------------------
C_SAMPLE = 250_000      # CPU arms time a 1/8 slice of the class dimension (cost is linear in C for fixed B)


def cpu_reference_run(steps: int, warmup: int, B: int, Cn: int = C_SAMPLE, scale_to: int = C_TOTAL):
    """Times the reference algorithm (oracle port: normalise, B x C GEMM, materialised elementwise margin
    passes, CrossEntropyLoss, top-k, autograd backward) in PyTorch CPU fp32 with every host thread.

    Bounded sample: B rows against Cn classes; every term of the cost (normalise W, the B x C passes, dW) is
    linear in C, so samples/s at C = scale_to is the measured rate divided by scale_to / Cn."""
    import torch
    from oracle import margin_oracle as mo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = mo.HeadConfig.default("arcface")
    g = torch.Generator().manual_seed(4)
    W = torch.randn(Cn, D, generator=g) * 0.01
    x = torch.randn(B, D, generator=g)
    y = torch.randint(0, Cn, (B,), generator=g)
    for _ in range(warmup):
        mo.autograd_step(cfg, mo.HeadState(), x, W, y, dtype=torch.float32)
    t0 = time.perf_counter()
    loss = None
    for _ in range(steps):
        loss = mo.autograd_step(cfg, mo.HeadState(), x, W, y, dtype=torch.float32)["loss"]
    dt = time.perf_counter() - t0
    k = scale_to / Cn
    return dict(value=B * steps / dt / k, ms_per_step=1e3 * dt / steps * k, cores=cores, loss=float(loss),
                sample=f"oracle port (torch CPU fp32, autograd, all B x C temporaries materialised), ArcFace "
                       f"B={B}, C={Cn} (1/{k:g} of the {scale_to} classes; time scaled x{k:g}), "
                       f"{steps} timed step(s) after {warmup} warm-up")


def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:  # noqa: BLE001
        return 0.0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32 if host_ram_gb() > 24 else 8
    # bounded sample: ~0.3 s per step at C_SAMPLE on 16 cores; shrink the class slice when many steps are requested
    Cn = C_SAMPLE if args.steps <= 30 else max(C_TOTAL // 64, C_TOTAL // (8 * -(-args.steps // 30)))
    r = cpu_reference_run(args.steps, min(args.warmup, 1), B, Cn)
    line = {
        "impl": "reference", "metric": "margin-head fwd+bwd samples/s at C=2M", "value": r["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------
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
    Cn, B = args.C, args.B
    peaks = load_peaks()

    if world > 1:
        head = pkg.ShardedMarginHead("arcface", Cn, s=64.0, m=0.5, easy_margin=False).to(dev)
        W = head.shard_parameter()
    else:
        head = pkg.ArcFace(D, Cn, s=64.0, m=0.5, easy_margin=False).to(dev)
        W = head.weight
    eng = head.engine if world > 1 else head._engine
    g = torch.Generator(device=dev).manual_seed(4 + rank)
    with torch.no_grad():
        W.normal_(0, 0.01, generator=g)            # generated on the device per shard, never shipped through the host
    x = torch.randn(B, D, device=dev, generator=g)
    y = torch.randint(0, Cn, (B,), device=dev, generator=g)
    x_host = x.cpu().pin_memory()
    y_host = y.cpu().pin_memory()
    res_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def step_resident():
        xg = x.detach().requires_grad_(True)
        W.grad = None
        out = head.fused_loss(xg, y)
        out.loss.backward()
        return out

    def step_e2e(x_dev, y_dev):
        x_dev.copy_(x_host, non_blocking=True)
        y_dev.copy_(y_host, non_blocking=True)
        xg = x_dev.detach().requires_grad_(True)
        W.grad = None
        out = head.fused_loss(xg, y_dev)
        out.loss.backward()
        res_host[0:1].copy_(out.loss.detach().reshape(1), non_blocking=True)
        res_host[1:2].copy_(out.acc1.reshape(1), non_blocking=True)
        res_host[2:3].copy_(out.acc5.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller reads loss/acc every step (model_utils.py:190)
        return float(res_host[0])

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
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    torch.cuda.synchronize()
    loss_val = float(out.loss.detach())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
        sampler.active = True

    # ---- device-resident arm, with per-kernel CUDA events on the launching stream ----------------------
    L.PROFILE = []
    ms_total = timed(step_resident, args.steps)
    prof, L.PROFILE = L.PROFILE, None
    # ---- the other backward mode (recompute: north_star's "backward recomputes logit tiles"), same protocol ---------
    alt = None
    if eng.stash_ok():
        eng.backward_mode = "recompute"
        for _ in range(3):
            step_resident()
        ms_alt = timed(step_resident, args.steps)
        eng.backward_mode = "auto"
        alt = {"backward": "recompute (forward writes nothing of size B x C; backward recomputes the logit tiles, 4 GEMM passes)",
               "value": B * world * args.steps / (ms_alt * 1e-3), "unit": "samples/s", "ms_per_step": ms_alt / args.steps}
        for _ in range(2):
            step_resident()
    # ---- end-to-end arm (host buffers) ------------------------------------------------------------------
    x_dev, y_dev = torch.empty_like(x), torch.empty_like(y)
    for _ in range(2):
        step_e2e(x_dev, y_dev)
    ms_e2e = timed(lambda: step_e2e(x_dev, y_dev), args.steps)
    if sampler:
        sampler.active = False
        sampler.stop()

    # ---- per-kernel summary -----------------------------------------------------------------------------
    kern = {}
    for name, e0, e1, launches in prof:
        k = kern.setdefault(name, [0.0, 0, 0])
        k[0] += e0.elapsed_time(e1)
        k[1] += 1 if launches else 0
        k[2] += launches
    n_launch = sum(k[2] for k in kern.values())
    B_g = B * world
    C_loc = W.shape[0]
    gemm_flops = 2.0 * B_g * C_loc * D                       # algorithmic FLOPs of ONE cos/dx/dw GEMM on this rank
    algo = {
        "mh_tc_forward": ("tensor", gemm_flops), "mh_tc_backward_g": ("tensor", gemm_flops),
        "mh_tc_backward_dx": ("tensor", gemm_flops), "mh_tc_backward_dw": ("tensor", gemm_flops),
        "mh_tc_backward_dw_fused": ("tensor", gemm_flops), "mh_tc_backward_dx_stash": ("tensor", gemm_flops),
        "mh_prologue_w": ("hbm", 6.0 * C_loc * D + 4.0 * C_loc),
        "mh_norm_backward_w": ("hbm", (4.0 + 2.0 + 4.0) * C_loc * D),
    }
    kernels = {}
    for name, (tot, calls, _l) in kern.items():
        if calls == 0:
            continue
        avg_ms = tot / calls
        ent = {"avg_ms": round(avg_ms, 4), "calls_per_step": calls / args.steps}
        if name in algo:
            kind, work = algo[name]
            ent["bound"] = kind
            ent["achieved"] = round(work / (avg_ms * 1e-3) / (1e12 if kind == "tensor" else 1e9), 2)
            ent["unit"] = "TFLOP/s" if kind == "tensor" else "GB/s"
        kernels[name] = ent
    dom = max((n for n in kernels if kernels[n].get("bound") == "tensor"), key=lambda n: kernels[n]["avg_ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom.replace("_stash", ""))
    roofline = {
        "kernel": dom, "bound": "tensor", "achieved": kernels[dom]["achieved"], "peak": peaks["tflops_sustained"],
        "unit": "TFLOP/s", "frac": round(kernels[dom]["achieved"] / peaks["tflops_sustained"], 4), "traffic": traffic,
        "peak_source": f"{peaks['source']} sustained bf16 GEMM (kernel timed inside a long step); burst "
                       f"{peaks['tflops_burst']}",
        "algorithmic_flops_per_launch": gemm_flops,
    }
    step_flops = 6.0 * B_g * C_loc * D * world
    value = B_g * args.steps / (ms_total * 1e-3)
    e2e_value = B_g * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                Bc = 32 if host_ram_gb() > 24 else 8
                r = cpu_reference_run(3, 1, Bc, min(C_SAMPLE, Cn), Cn)
                cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                       "sample": f"failed: {ex!r}"}
        line = {
            "metric": "margin-head fwd+bwd samples/s at C=2M", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "B_per_gpu": B, "C": Cn, "d": D, "parallelism": f"class-shard x{world}",
                       "backward": ("stash (forward writes bf16 exp2(z-ref); 3 GEMM passes per step)"
                                    if eng.stash_ok() else "recompute (4 GEMM passes per step)"),
                       "l2": "inputs_exceed_l2 (W fp32 4.1 GB + bf16 2 GB per step vs 126 MB L2)", "loss": loss_val},
            "pct_of_bf16_peak": {"algorithmic_tflops": round(step_flops / (ms_total / args.steps * 1e-3) / 1e12 / world, 2),
                                 "of_burst": round(step_flops / world / (ms_total / args.steps * 1e-3) / 1e12 / peaks["tflops_burst"], 4),
                                 "of_sustained": round(step_flops / world / (ms_total / args.steps * 1e-3) / 1e12 / peaks["tflops_sustained"], 4)},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8) * world, "d2h_bytes_per_step": 12 * world},
            "gpu_launches": n_launch,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "alt_backward": alt,
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
    ap.add_argument("--C", type=int, default=C_TOTAL)
    ap.add_argument("--B", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
