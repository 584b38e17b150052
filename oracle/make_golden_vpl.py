"""Generate tests/golden/vpl_*.npz from the UNMODIFIED reference VPLArcFace (build container only).

    python oracle/make_golden_vpl.py            # needs /root/reference

Runs reference forward -> nn.CrossEntropyLoss -> accuracy -> autograd backward in float64 on CPU for two consecutive
steps (so that the second step sees a populated memory bank with decayed lifetimes), asserts that
oracle/vpl_oracle.py reproduces every output, and stores the reference's outputs.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import vpl_oracle as vo  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def main():
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import criterion as C  # type: ignore
        from main_code.utils.metrics import accuracy  # type: ignore
    out_dir = os.path.join(ROOT, "tests", "golden")
    cases = [("vpl_easy", dict(easy_margin=True, lamda=0.15, delta=100), 8, 61, 0, 1.0),
             ("vpl_hard_margin", dict(easy_margin=False, lamda=0.3, delta=2), 8, 61, 1, 1.0),
             ("vpl_gradscale", dict(easy_margin=True, lamda=0.15, delta=100), 16, 130, 2, 1024.0)]
    for name, kw, B, Cn, seed, gs in cases:
        cfg = vo.VplConfig(**kw)
        with contextlib.redirect_stdout(io.StringIO()):
            head = C.VPLArcFace(512, Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin, lamda=cfg.lamda, delta=cfg.delta).double()
        mem, life = torch.zeros(Cn, 512, dtype=torch.float64), torch.zeros(Cn, dtype=torch.float64)
        steps = []
        for step in range(3):
            x, W, labels = vo.make_inputs(B, Cn, 512, seed * 10 + step)
            with torch.no_grad():
                head.weight.copy_(W.double())
            head.weight.grad = None
            xr = x.double().requires_grad_(True)
            (pre, logits), norms, loss_g, one_hot = head(xr, labels)
            loss = torch.nn.CrossEntropyLoss()(logits, labels)
            a1, a5 = accuracy(pre, labels, (1, 5))
            (loss * gs).backward()
            mine = vo.loss_and_grads(cfg, x, W, labels, mem, life, True, gs)
            mem, life = mine["mem"], mine["life"]
            assert torch.allclose(mine["mem"], head.mem) and torch.allclose(mine["life"], head.life)
            assert torch.allclose(mine["logits"], logits.detach(), rtol=1e-12, atol=1e-11) and torch.allclose(mine["pre"], pre.detach(), rtol=1e-12, atol=1e-11)
            assert abs(float(mine["loss"]) - float(loss)) < 1e-12 * abs(float(loss))
            assert torch.allclose(mine["dx"], xr.grad, rtol=1e-9, atol=1e-13), (mine["dx"] - xr.grad).abs().max()
            assert torch.allclose(mine["dW"], head.weight.grad, rtol=1e-9, atol=1e-13)
            assert abs(float(mine["acc1"]) - float(a1)) < 1e-9 and abs(float(mine["acc5"]) - float(a5)) < 1e-9
            steps.append(dict(loss=float(loss), acc1=float(a1), acc5=float(a5), dx=xr.grad.numpy().copy(),
                              dW=head.weight.grad.numpy().copy(), life_sum=float(head.life.sum()),
                              mem_sum=float(head.mem.sum()), n_active=int((head.life > 0).sum())))
        flat = {}
        for i, st in enumerate(steps):
            for k, v in st.items():
                flat[f"s{i}_{k}"] = v
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), B=B, C=Cn, seed=seed, grad_scale=gs, n_steps=len(steps),
                            easy_margin=cfg.easy_margin, lamda=cfg.lamda, delta=cfg.delta, s=cfg.s, m=cfg.m, **flat)
        print(name, [round(st["loss"], 4) for st in steps], [st["n_active"] for st in steps])


if __name__ == "__main__":
    main()
