"""Scratch probe: loss trajectories of the GradScaler training-step test under a few optimiser settings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torchvision
import face_recognition_models_b200 as pkg
for lr, mom, scale0 in ((0.02, 0.0, 1024.0), (0.05, 0.0, 1024.0), (0.01, 0.9, 1024.0), (0.005, 0.0, 65536.0)):
    torch.manual_seed(0)
    net = torchvision.models.resnet18(weights=None); net.fc = nn.Linear(512, 512); net = net.cuda()
    head = pkg.ArcFace(512, 1000, s=64.0, m=0.5, easy_margin=False).cuda()
    opt = torch.optim.SGD(list(net.parameters()) + list(head.parameters()), lr=lr, momentum=mom, weight_decay=5e-4)
    scaler = torch.amp.GradScaler("cuda", init_scale=scale0)
    images = torch.randn(32, 3, 112, 112, device="cuda"); target = torch.randint(0, 1000, (32,), device="cuda")
    losses = []
    for _ in range(14):
        with torch.autocast("cuda"):
            feats = net(images)
        out = head.fused_loss(feats, target)
        opt.zero_grad(set_to_none=True)
        scaler.scale(out.loss).backward(); scaler.step(opt); scaler.update()
        losses.append(round(out.loss.item(), 2))
    print(lr, mom, scale0, losses)
