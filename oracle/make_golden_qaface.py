"""Generate tests/golden/qaface_*.npz from the UNMODIFIED reference QAFace (build container only).

    python oracle/make_golden_qaface.py            # needs /root/reference

Runs the reference forward -> nn.CrossEntropyLoss -> accuracy -> autograd backward in float64 on CPU for three
consecutive steps (the second and third see a populated memory bank, EMA statistics and decayed lifetimes), asserts that
oracle/qaface_oracle.py reproduces every output, and stores the reference's outputs.  Step 0 differentiates through
``minput`` as well; on later steps ``minput`` carries no gradient and the reference's ``muy`` / ``std`` buffers are
detached after each step (its own second backward fails otherwise, see the oracle's header).
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
from oracle import qaface_oracle as qo  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def main():
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import criterion as C  # type: ignore
        from main_code.utils.metrics import accuracy  # type: ignore
    out_dir = os.path.join(ROOT, "tests", "golden")
    cases = [("qaface_easy", dict(easy_margin=True, delta=1000, tto=2.0, alpha=0.99), 8, 61, 0, 1.0),
             ("qaface_hard_margin", dict(easy_margin=False, delta=2, tto=1.0, alpha=0.9), 8, 61, 1, 1.0),
             ("qaface_gradscale", dict(easy_margin=True, delta=1000, tto=2.0, alpha=0.99), 16, 130, 2, 1024.0)]
    for name, kw, B, Cn, seed, gs in cases:
        cfg = qo.QaConfig(**kw)
        with contextlib.redirect_stdout(io.StringIO()):
            head = C.QAFace(512, Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin, delta=cfg.delta, tto=cfg.tto,
                            alpha=cfg.alpha).double()
        st = qo.QaState.fresh(Cn)
        steps = []
        for step in range(3):
            x, minput, W, labels = qo.make_inputs(B, Cn, 512, seed * 10 + step)
            with torch.no_grad():
                head.weight.copy_(W.double())
            head.weight.grad = None
            mg = step == 0
            xr = x.double().requires_grad_(True)
            mr = minput.double().requires_grad_(mg)
            (pre, logits), norms, loss_g, one_hot = head(xr, mr, labels)
            loss = torch.nn.CrossEntropyLoss()(logits, labels)
            a1, a5 = accuracy(pre, labels, (1, 5))
            (loss * gs).backward()
            head.muy, head.std = head.muy.detach(), head.std.detach()       # see the module docstring
            mine = qo.loss_and_grads(cfg, st, x, minput, W, labels, True, gs, minput_grad=mg)
            st = mine["state"]
            assert torch.allclose(st.mem, head.mem, rtol=1e-12, atol=1e-14) and torch.allclose(st.life, head.life)
            assert abs(st.muy - float(head.muy)) < 1e-12 and abs(st.std - float(head.std)) < 1e-12
            assert torch.allclose(mine["logits"].detach(), logits.detach(), rtol=1e-12, atol=1e-10)
            assert torch.allclose(mine["pre"].detach(), pre.detach(), rtol=1e-12, atol=1e-10)
            assert abs(float(mine["loss"]) - float(loss)) < 1e-12 * abs(float(loss))
            assert torch.allclose(mine["dx"], xr.grad, rtol=1e-8, atol=1e-12), (mine["dx"] - xr.grad).abs().max()
            assert torch.allclose(mine["dW"], head.weight.grad, rtol=1e-8, atol=1e-12)
            if mg:
                assert torch.allclose(mine["dminput"], mr.grad, rtol=1e-8, atol=1e-12)
            assert abs(float(mine["acc1"]) - float(a1)) < 1e-9 and abs(float(mine["acc5"]) - float(a5)) < 1e-9
            rec = dict(loss=float(loss), acc1=float(a1), acc5=float(a5), dx=xr.grad.numpy().copy(),
                       dW=head.weight.grad.numpy().copy(), life_sum=float(head.life.sum()), mem_sum=float(head.mem.sum()),
                       n_active=int((head.life > 0).sum()), muy=float(head.muy), std=float(head.std))
            if mg:
                rec["dminput"] = mr.grad.numpy().copy()
            steps.append(rec)
        flat = {}
        for i, s_ in enumerate(steps):
            for k, v in s_.items():
                flat[f"s{i}_{k}"] = v
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), B=B, C=Cn, seed=seed, grad_scale=gs, n_steps=len(steps),
                            easy_margin=cfg.easy_margin, delta=cfg.delta, tto=cfg.tto, alpha=cfg.alpha, s=cfg.s, m=cfg.m, **flat)
        print(name, [round(s_["loss"], 4) for s_ in steps], [s_["n_active"] for s_ in steps],
              [round(s_["muy"], 3) for s_ in steps])


if __name__ == "__main__":
    main()
