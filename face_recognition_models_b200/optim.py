"""SGD for the class-centre parameter of a margin head, fused with the next step's W prologue.

The reference trains everything with one `optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=5e-4)`
(main_code/utils/model_utils.py:557), stepped through `GradScaler.step` (model_utils.py:186).  For the head parameter
that is a 20 B/element streaming pass over C x 512 floats, and the next forward then streams W again to normalise it.
`HeadSGD` does both in one pass (`mh_sgd_step_w`): the update is torch.optim.SGD's (dampening 0, no nesterov), and
the same kernel leaves the bf16 `w_hat` and `inv_norm` of the updated W in the head's workspace, so the next
tensor-core forward skips `mh_prologue_w`.

Use it next to the unchanged optimizer of the backbone:

    opt_backbone = optim.SGD(model.backbone.parameters(), lr=lr, momentum=0.9, weight_decay=5e-4)
    opt_head = HeadSGD([model.arcface], lr=lr, momentum=0.9, weight_decay=5e-4)
    ...
    scaler.step(opt_backbone); scaler.step(opt_head); scaler.update()

The shadow is keyed on the parameter's storage and autograd version counter, so `load_state_dict`, `W.copy_()`,
`W.mul_()` ... under `torch.no_grad()` drop it automatically; a write through `W.data` does not move the counter -
call `head.head_engine().invalidate_shadow()` after one.  `lr` is passed to the kernel by value (a CUDA-graph capture
of `step()` freezes it).

It is a `torch.optim.Optimizer`: LR schedulers, `zero_grad`, `state_dict` (`momentum_buffer`, interchangeable with
torch.optim.SGD's entry for the same parameter) and GradScaler work as usual.  GradScaler hands `grad_scale` /
`found_inf` to `step()` as device scalars (no host sync) and the kernel divides by the scale itself, so the head
gradient is never rewritten by an unscale pass; GradScaler's own inf/nan scan of the gradient still runs.

Divergence from the reference's single optimizer: `GradScaler.step` decides the inf-skip PER OPTIMIZER, so an overflow
that shows up only in the backbone gradients skips the backbone step while the head still updates (and the reverse);
the reference skips the whole step.  The head side of that is covered by `coupled=[opt_backbone]` + `couple_scaler(scaler)`:
`step()` then ORs the other optimizers' `found_inf` (recorded by GradScaler at `scaler.unscale_(opt_backbone)`, which must
be called before `scaler.step(opt_head)`) into its own before the kernel runs, still without a host sync.  The reverse
(an overflow seen only in the 4 GB head gradient skipping the backbone step) is not covered: the scale is halved for
the next step either way, and the backbone gradient of such a step is finite by construction.
"""
from typing import Iterable

import torch

from . import _lib as L


class HeadSGD(torch.optim.Optimizer):
    _step_supports_amp_scaling = True      # GradScaler hands over grad_scale / found_inf instead of unscaling first

    def __init__(self, heads: Iterable[torch.nn.Module], lr: float, momentum: float = 0.9, weight_decay: float = 5e-4,
                 coupled: Iterable[torch.optim.Optimizer] = ()):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("lr, momentum and weight_decay must be >= 0")
        heads = list(heads)
        if not heads:
            raise ValueError("HeadSGD needs at least one margin head")
        self._engine_of = {}
        self._coupled = list(coupled)          # optimizers sharing this head's GradScaler (see the module docstring)
        self._scaler = None
        params = []
        for h in heads:
            if not hasattr(h, "head_parameter") or not hasattr(h, "head_engine"):
                raise TypeError(f"{type(h).__name__} is not a margin head of this package")
            p = h.head_parameter()
            self._engine_of[id(p)] = h
            params.append(p)
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    def couple_scaler(self, scaler) -> None:
        """Give step() access to the GradScaler so it can read the coupled optimizers' found_inf."""
        self._scaler = scaler

    def _combined_found_inf(self, found_inf):
        """OR of this optimizer's found_inf with the coupled optimizers' (device tensors, no host sync)."""
        if found_inf is None or not self._coupled or self._scaler is None:
            return found_inf
        states = getattr(self._scaler, "_per_optimizer_states", {})
        out = found_inf.clone()
        for opt in self._coupled:
            st = states.get(id(opt))
            if st:
                for v in st.get("found_inf_per_device", {}).values():
                    out = torch.maximum(out, v.to(out.device))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise L.MarginHeadError("HeadSGD runs on CUDA parameters only (no CPU fallback)")
                g = p.grad
                if g.is_sparse:
                    raise L.MarginHeadError("HeadSGD does not take sparse gradients")
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                engine = self._engine_of[id(p)].head_engine()
                engine.sgd_step(p, g.contiguous(), st["momentum_buffer"], float(group["lr"]), float(group["momentum"]),
                                float(group["weight_decay"]), getattr(self, "grad_scale", None),
                                self._combined_found_inf(getattr(self, "found_inf", None)))   # set by GradScaler.step

        return loss
