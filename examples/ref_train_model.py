"""BASELINE config 5: the reference's OWN `train_model` (main_code/utils/model_utils.py:147-216), unmodified, driving the
B200 head - DDP backbone + class-sharded head on N GPUs, or one GPU.

    python baseline/stage_ref.py                                                   # once, in the build container
    torchrun --nproc-per-node 8 examples/ref_train_model.py --classes 2000000 --steps 30
    python examples/ref_train_model.py --check-grads                               # + the DDP scale-convention check

What stays the reference's: `train_model` itself (imported from the staged copy under baseline/_ref: autocast, GradScaler,
`criterion(logits, target)`, `accuracy(cosine_s, target, (1, 5))`, `scaler.step(optimizer)`, the `.item()` reads, the
meters and `wandb.log`), `nn.CrossEntropyLoss` semantics, `optim.SGD(model.parameters(), lr, momentum=0.9,
weight_decay=5e-4)` (model_utils.py:557) and `GradScaler()` (:559).

What the shim supplies (the "3 lines" of INTEGRATION.md section 1, done from outside so that train_model's source is
untouched): train_model wants materialised `logits` / `cosine_s` of size B x C (model_utils.py:177-182); at C = 2M that is
8 GB per tensor per GPU, so the model hands out two light proxies instead -
  * `FusedLogits`  - carries the fused loss;  `FusedCriterion()(proxy, target)` returns it (and falls back to
                     nn.CrossEntropyLoss for real tensors);
  * `FusedCosine`  - carries acc@1 / acc@5;   `model_utils.accuracy` is rebound to a version that unwraps it.
Environment stubs: `alive_progress` (missing in the image, imported by utils/dataset.py at module scope) and a no-op
`wandb.log` (no network).  Backbone: torchvision resnet50(weights=None) + fc -> 512, i.e. backbones.py:16-18 without the
checkpoint download.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import types

import torch
import torch.distributed as dist
import torch.nn as nn
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import face_recognition_models_b200 as pkg  # noqa: E402


def import_reference_model_utils():
    ref = os.path.join(ROOT, "baseline", "_ref", "main_code")
    if not os.path.isfile(os.path.join(ref, "utils", "model_utils.py")):
        raise SystemExit("baseline/_ref is empty: run `python baseline/stage_ref.py` in the build container first")
    sys.modules.setdefault("alive_progress", types.SimpleNamespace(alive_bar=lambda *a, **k: None))
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.path.insert(0, ref)                               # the reference runs from main_code/: `from utils.config import *`
    import warnings
    warnings.filterwarnings("ignore")
    import utils.model_utils as mu
    mu.wandb = types.SimpleNamespace(log=lambda *a, **k: None)        # train_model's only use of wandb (:199-208)
    ref_accuracy = mu.accuracy

    def accuracy(output, target, topk=(1,)):
        if isinstance(output, FusedCosine):
            assert tuple(topk) == (1, 5)
            return [output.acc1.reshape(1), output.acc5.reshape(1)]   # percent, shape [1]: what metrics.accuracy returns
        return ref_accuracy(output, target, topk)
    mu.accuracy = accuracy
    return mu


class FusedLogits:
    def __init__(self, loss):
        self.loss = loss


class FusedCosine:
    def __init__(self, acc1, acc5):
        self.acc1, self.acc5 = acc1, acc5


class FusedCriterion(nn.Module):
    """`criterion(logits, target)` of model_utils.py:179: the fused head already reduced the cross-entropy."""

    def __init__(self):
        super().__init__()
        self.ce = nn.CrossEntropyLoss()

    def forward(self, logits, target):
        return logits.loss if isinstance(logits, FusedLogits) else self.ce(logits, target)


class FusedFaceNet(nn.Module):
    """The reference's `ArcFaceNet` (criterion.py:303-325) with the fused head: same attribute names (`backbone`,
    `arcface`), same call protocol `model(images, labels) -> (output, norms, loss_g, one_hot)`."""

    def __init__(self, num_classes, backbone="resnet50", world=1, local_rank=0, time_head=False):
        super().__init__()
        net = getattr(torchvision.models, backbone)(weights=None)
        net.fc = nn.Linear(net.fc.in_features, 512)
        self.backbone = net
        self.world = world
        if world > 1:
            self.arcface = pkg.ShardedMarginHead("arcface", num_classes, s=64.0, m=0.5, easy_margin=False, dx_scale=world)
        else:
            self.arcface = pkg.ArcFace(512, num_classes, s=64.0, m=0.5, easy_margin=False)
        self.loss_model = "arcface"
        self._ddp = None
        self.time_head = time_head
        self.head_events = []

    def wrap_ddp(self, local_rank):
        # only the backbone is data-parallel; the class shard is model-parallel and stays out of DDP
        self._ddp = nn.parallel.DistributedDataParallel(self.backbone, device_ids=[local_rank])

    def forward(self, x, labels=None):
        features = (self._ddp or self.backbone)(x)
        if not self.training:
            return features
        assert labels is not None
        if self.time_head:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        out = self.arcface.fused_loss(features, labels)
        if self.time_head:
            e[1].record()
            out.loss.register_hook(lambda g, ev=e[2]: (ev.record(), g)[1])          # backward reaches the head
            features.register_hook(lambda g, ev=e[3]: (ev.record(), g)[1])          # d(features) is ready: head backward done
            self.head_events.append(e)
        return [FusedCosine(out.acc1, out.acc5), FusedLogits(out.loss)], out.norms, out.loss_g, None


def check_grads(dev, world, rank, local):
    """DDP scale convention: backbone gradients of DDP(backbone) + ShardedMarginHead(dx_scale=world) equal the
    single-process gradient on the concatenated batch; so does every rank's dW shard.  resnet18 in eval() mode (batch
    statistics would differ between B_local and B_global by construction), fp32 features."""
    torch.manual_seed(0)
    Bl, Cn = 16, 4099
    full = FusedFaceNet(Cn, "resnet18", world=1).to(dev)
    full.backbone.eval()
    g = torch.Generator(device=dev).manual_seed(7)
    images = torch.randn(Bl * world, 3, 112, 112, device=dev, generator=g)
    y = torch.randint(0, Cn, (Bl * world,), device=dev, generator=g)
    if world > 1:                                              # same weights / data on every rank
        for p in full.parameters():
            dist.broadcast(p.data, 0)
        dist.broadcast(images, 0)
        dist.broadcast(y, 0)
    out = full.arcface.fused_loss(full.backbone(images), y)
    out.loss.backward()
    ref_grads = [p.grad.clone() for p in full.backbone.parameters()]
    ref_dW = full.arcface.weight.grad.clone()
    if world == 1:
        print("check-grads: single GPU, nothing to compare (run under torchrun)")
        return True
    shard = FusedFaceNet(Cn, "resnet18", world=world).to(dev)
    shard.backbone.load_state_dict(full.backbone.state_dict())
    shard.backbone.eval()
    shard.wrap_ddp(local)
    b, e = shard.arcface.c_begin, shard.arcface.c_end
    with torch.no_grad():
        shard.arcface.shard_parameter().copy_(full.arcface.weight[b:e])
    xl, yl = images[rank * Bl:(rank + 1) * Bl], y[rank * Bl:(rank + 1) * Bl]
    o2 = shard.arcface.fused_loss(shard._ddp(xl), yl)
    o2.loss.backward()                                         # DDP averages the backbone grads over the ranks
    cos = torch.nn.functional.cosine_similarity
    a = torch.cat([p.grad.flatten() for p in shard.backbone.parameters()]).double()
    r = torch.cat([gr.flatten() for gr in ref_grads]).double()
    cb = float(cos(a, r, dim=0))
    nb = float(a.norm() / r.norm())
    cw = float(cos(shard.arcface.shard_parameter().grad.flatten().double(), ref_dW[b:e].flatten().double(), dim=0))
    lr = abs(float(o2.loss) - float(out.loss)) / abs(float(out.loss))
    res = torch.tensor([cb, nb, cw, lr], device=dev, dtype=torch.float64)
    lo, hi = res.clone(), res.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = bool(lo[0] > 0.9995 and abs(lo[1] - 1) < 5e-3 and abs(hi[1] - 1) < 5e-3 and lo[2] > 0.9995 and hi[3] < 2e-3)
    if rank == 0:
        print(f"check-grads (world {world}, B_local {Bl}, C {Cn}): backbone grad cos >= {float(lo[0]):.6f}, norm ratio in "
              f"[{float(lo[1]):.5f}, {float(hi[1]):.5f}], dW shard cos >= {float(lo[2]):.6f}, loss rel <= {float(hi[3]):.2e}  "
              f"-> {'OK' if ok else 'FAIL'}")
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=64, help="per GPU (BASELINE cfg1/cfg5: 64)")
    ap.add_argument("--classes", type=int, default=10575)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--check-grads", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    mu = import_reference_model_utils()
    ok = True
    if a.check_grads:
        ok = check_grads(dev, world, rank, local)

    torch.manual_seed(5 + rank)
    model = FusedFaceNet(a.classes, a.backbone, world=world, time_head=True).to(dev)
    if world > 1:
        model.wrap_ddp(local)
    criterion = FusedCriterion().to(dev)                                    # model_utils.py:556
    optimizer = torch.optim.SGD(model.parameters(), lr=a.lr, momentum=0.9, weight_decay=5e-4)   # :557, one SGD over everything
    scaler = torch.amp.GradScaler("cuda")                                   # :559 (default init_scale 65536)
    images = torch.randn(a.batch, 3, 112, 112)
    target = torch.randint(0, a.classes, (a.batch,))
    warm, timed = [(images, target)] * 5, [(images, target)] * a.steps      # len() and iteration are all train_model needs
    args = types.SimpleNamespace(lambda_g=0.0, print_freq=10 if rank == 0 else 10 ** 9)
    # warm-up epoch (GradScaler's first overflow skips, cuDNN autotune, workspaces), then the timed epoch
    mu.train_model(model, warm, criterion, optimizer, scaler, dev, 0, 1, args)
    model.head_events.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    loss_avg = mu.train_model(model, timed, criterion, optimizer, scaler, dev, 1, 1, args)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.time() - t0
    fwd = sum(e[0].elapsed_time(e[1]) for e in model.head_events) / len(model.head_events)
    bwd = sum(e[2].elapsed_time(e[3]) for e in model.head_events) / len(model.head_events)
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t)
    if rank == 0:
        step_ms = 1e3 * dt / a.steps
        print(f"ref_train_model: reference train_model (unmodified, baseline/_ref) + fused ArcFace head, {a.backbone}, "
              f"world {world}, B {a.batch}/GPU, C {a.classes}: {a.steps} steps, {step_ms:.2f} ms/step, "
              f"{a.batch * world * a.steps / dt:.1f} img/s, avg loss {float(loss_avg):.4f}; head fwd {fwd:.3f} ms + bwd {bwd:.3f} ms "
              f"= {100 * (fwd + bwd) / step_ms:.1f} % of the step (the torch.optim.SGD update of the head parameter is outside this share)")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
