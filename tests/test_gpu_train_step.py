"""GPU: the head as a drop-in inside a reference-style training step (BASELINE config 1 shape, small backbone)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _backbone():
    import torchvision
    net = torchvision.models.resnet18(weights=None)
    net.fc = nn.Linear(net.fc.in_features, 512)
    return net


def test_training_steps_reduce_loss_with_gradscaler():
    """autocast backbone (fp16 features) + GradScaler + SGD as model_utils.py:176-187; the fused head must train."""
    import face_recognition_models_b200 as pkg
    torch.manual_seed(0)
    net = _backbone().cuda()
    head = pkg.ArcFace(512, 1000, s=64.0, m=0.5, easy_margin=False).cuda()
    # plain SGD with a small step: on a fixed batch the loss must then fall monotonically-ish (momentum 0.9 at
    # lr 0.05 overshoots on the s=64 logits within the first steps, which says nothing about the head)
    opt = torch.optim.SGD(list(net.parameters()) + list(head.parameters()), lr=0.02, momentum=0.0, weight_decay=5e-4)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    images = torch.randn(32, 3, 112, 112, device="cuda")
    target = torch.randint(0, 1000, (32,), device="cuda")
    losses = []
    for _ in range(12):
        with torch.autocast("cuda"):
            feats = net(images)
        assert feats.dtype == torch.float16
        out = head.fused_loss(feats, target)
        opt.zero_grad(set_to_none=True)
        scaler.scale(out.loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(out.loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0] - 0.5, losses          # memorising 32 samples: the loss must fall


def test_compat_and_fused_paths_agree_inside_a_model():
    """model(images, target) -> 4-tuple -> CrossEntropyLoss (unchanged train_model) == fused_loss on the same batch."""
    import face_recognition_models_b200 as pkg
    torch.manual_seed(1)
    net = _backbone().cuda()
    head = pkg.CurricularFace(512, 700, m=0.5, s=64.0, momentum=0.01).cuda()
    images = torch.randn(16, 3, 112, 112, device="cuda")
    target = torch.randint(0, 700, (16,), device="cuda")

    def run(fused):
        net.zero_grad(set_to_none=True)
        head.zero_grad(set_to_none=True)
        head.t.zero_()
        feats = net(images)
        if fused:
            head.mode = "exact"
            loss = head.fused_loss(feats, target).loss
        else:
            (cos_s, logits), norms, loss_g, one_hot = head(feats, target)
            loss = nn.CrossEntropyLoss()(logits, target)
        loss.backward()
        return loss.item(), net.fc.weight.grad.clone(), head.kernel.grad.clone(), head.t.clone()

    l0, g0, k0, t0 = run(False)
    l1, g1, k1, t1 = run(True)
    assert abs(l0 - l1) < 1e-4 * abs(l0)
    assert torch.allclose(t0, t1, atol=1e-7)
    cs = torch.nn.functional.cosine_similarity
    assert cs(g0.flatten(), g1.flatten(), dim=0) > 0.99999 and cs(k0.flatten(), k1.flatten(), dim=0) > 0.99999
