"""Reference-style training step with the B200 head as a drop-in (BASELINE configs 1 and 5).

    python examples/train_step.py --steps 5                                   # 1 GPU, ArcFace, C=10,575
    torchrun --nproc-per-node 8 examples/train_step.py --classes 2000000      # DDP backbone + class-sharded head

Mirrors main_code/utils/model_utils.py:168-192 (autocast backbone, GradScaler, SGD momentum 0.9 / wd 5e-4,
loss.item() every step) with the three-line change of INTEGRATION.md: the loss and the top-1/5 accuracies
come from head.fused_loss instead of materialised logits.  Synthetic 112x112 faces, random-init ResNet-50.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_recognition_models_b200 as pkg  # noqa: E402


def build_backbone(name="resnet50"):
    net = getattr(torchvision.models, name)(weights=None)          # backbones.py:13-18 without the download
    net.fc = nn.Linear(net.fc.in_features, 512)
    return net


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--classes", type=int, default=10575)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--lambda_g", type=float, default=0.0)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--init_scale", type=float, default=1024.0)
    ap.add_argument("--lfw_pairs", type=int, default=6000, help="synthetic verification pairs (0 = skip)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1 + rank)
    backbone = build_backbone(a.backbone).to(dev)
    if world > 1:
        backbone = nn.parallel.DistributedDataParallel(backbone, device_ids=[local])
        head = pkg.ShardedMarginHead("arcface", a.classes, s=64.0, m=0.5, easy_margin=False, dx_scale=world).to(dev)
    else:
        head = pkg.ArcFace(512, a.classes, s=64.0, m=0.5, easy_margin=False).to(dev)
    # model_utils.py:557 builds one SGD over everything; here the head parameter goes to HeadSGD (same update, fused
    # with the next step's W prologue), the backbone keeps torch.optim.SGD
    opt = torch.optim.SGD(backbone.parameters(), lr=a.lr, momentum=0.9, weight_decay=5e-4)
    opt_head = pkg.HeadSGD([head], lr=a.lr, momentum=0.9, weight_decay=5e-4)
    # model_utils.py:559 uses the default GradScaler (init_scale 65536): with fp16 features its first steps overflow and
    # are skipped while the scale halves; start lower so that a short demo run shows the loss moving
    scaler = torch.amp.GradScaler("cuda", init_scale=a.init_scale)
    images = torch.randn(a.batch, 3, 112, 112, device=dev)
    target = torch.randint(0, a.classes, (a.batch,), device=dev)
    for step in range(a.steps):
        t0 = time.time()
        with torch.autocast("cuda"):
            feats = backbone(images)                                   # fp16 features, as under the reference's autocast
        out = head.fused_loss(feats, target)
        loss = out.loss + a.lambda_g * out.loss_g
        opt.zero_grad(set_to_none=True)
        opt_head.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.step(opt_head)
        scaler.update()
        lv = loss.item()
        if rank == 0:
            print(f"step {step}: loss {lv:.4f} acc@1 {float(out.acc1):.2f} acc@5 {float(out.acc5):.2f} "
                  f"({a.batch * world / (time.time() - t0):.1f} img/s)")
    if rank == 0 and a.lfw_pairs > 0:
        # LFW-style 10-fold verification (model_utils.py:416-474) on synthetic embedding pairs: the per-pair cosine runs
        # on the GPU (mh_pair_cosine), the fold statistics are the reference's scikit-learn calls
        g = torch.Generator(device=dev).manual_seed(5)
        half = a.lfw_pairs // 2

        def unit(n):
            return torch.nn.functional.normalize(torch.randn(n, 512, device=dev, generator=g), dim=1)

        centre = unit(half)                                       # same-identity pairs share a centre
        e1 = torch.cat([centre + 1.5 * unit(half), unit(half) + 1.5 * unit(half)])
        e2 = torch.cat([centre + 1.5 * unit(half), unit(half) + 1.5 * unit(half)])
        same = torch.cat([torch.ones(half, dtype=torch.long), torch.zeros(half, dtype=torch.long)])
        acc, acc_std, auc, auc_std = pkg.verification.cross_validate_kfold(e1, e2, same, k_fold=10)
        print(f"LFW-protocol 10-fold on {a.lfw_pairs} synthetic pairs: accuracy {acc:.3f}% +- {acc_std:.3f}, AUC {auc:.4f} +- {auc_std:.4f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
