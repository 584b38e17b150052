"""Drop-in nn.Modules for the reference's margin heads (main_code/utils/criterion.py).

Each class has the reference's constructor signature and defaults, the same parameter / buffer
names, shapes, layouts and initialisation (so ``state_dict``s interchange and a wrapper such as
``ArcFaceNet`` can assign ``self.arcface = ArcFace(...)`` unchanged), and two call styles:

* ``head(feats, labels)``            -> ``([pre_margin_logits, logits], norms, loss_g, one_hot)``
  the reference's 4-tuple (compat mode, materialises B x C in fp32 - small class counts only;
  this is what the unchanged ``train_model`` consumes, model_utils.py:177-182).
* ``head.fused_loss(feats, labels)`` -> ``FusedOutput(loss, loss_g, acc1, acc5, norms)``
  the fused path: cross-entropy (mean), top-1/top-5 and the gradients come from the sm_100a
  kernels and nothing of size B x C is materialised in the forward.

All compute goes through libmargin_head.so; there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .functional import DenseMarginLogitsFn, FusedMarginLossFn, HeadEngine, ShardInfo, sphere_lambda


class FusedOutput(NamedTuple):
    loss: torch.Tensor       # mean cross-entropy over the batch (nn.CrossEntropyLoss(), model_utils.py:556)
    loss_g: torch.Tensor     # MagFace regulariser (0 for the other heads); caller adds lambda_g * loss_g
    acc1: torch.Tensor       # top-1 accuracy in percent on the pre-margin logits (metrics.py:3-16)
    acc5: torch.Tensor
    norms: torch.Tensor      # [B, 1], as returned by the reference head


class _MarginHeadBase(nn.Module):
    family = ""
    layout = "CD"
    param_name = "weight"

    def _init_engine(self, num_classes: int, mode: str = "tc", **cfg):
        self._engine = HeadEngine(self.family, self.layout, num_classes, cfg, mode=mode)
        self._num_classes = num_classes
        # device-side state vector shared with the kernels (see MH_STATE_FLOATS in margin_head.h)
        self.register_buffer("_mh_state", torch.zeros(L.STATE_FLOATS), persistent=False)
        self._margins_override: Optional[torch.Tensor] = None

    # precision / path selection: "tc" (bf16 tensor cores) or "exact" (fp32 SIMT, small C)
    @property
    def mode(self) -> str:
        return self._engine.mode

    @mode.setter
    def mode(self, v: str):
        assert v in ("tc", "exact")
        self._engine.mode = v

    # backward of the tensor-core path: "auto" (stash when the head allows it), "stash" or "recompute"
    @property
    def backward_mode(self) -> str:
        return self._engine.backward_mode

    @backward_mode.setter
    def backward_mode(self, v: str):
        assert v in ("auto", "stash", "recompute")
        self._engine.backward_mode = v

    def _param(self) -> torch.Tensor:
        return getattr(self, self.param_name)

    # what optim.HeadSGD needs: the class-centre parameter and the engine whose workspace holds its w_hat
    def head_parameter(self) -> torch.Tensor:
        return self._param()

    def head_engine(self):
        return self._engine

    # hooks for heads with state ------------------------------------------------------------------
    def _push_state(self):
        pass

    def _pull_state(self):
        pass

    def _pre_forward(self, feats):
        pass

    def _sample_margins(self, feats, labels):
        return None

    # ------------------------------------------------------------------------------------------------
    def _check(self, feats, labels):
        if labels is None:
            raise ValueError("labels are required in training mode")
        if feats.dim() != 2 or feats.shape[1] != L.D:
            raise ValueError(f"expected feats of shape [B, {L.D}], got {tuple(feats.shape)}")
        if labels.shape[0] != feats.shape[0]:
            raise ValueError("feats / labels batch mismatch")

    def prefetch(self) -> None:
        """Enqueue the W prologue of the NEXT forward now, on the current stream.  It depends on the class centres only, not
        on the batch, so a training loop can call this before the batch's host-to-device copy or before / beside the
        backbone forward (on a side stream: the prologue is HBM-bound, the backbone tensor-bound) and the head's forward
        then starts with its GEMM.  Good for exactly one forward; any write to the parameter in between invalidates it
        (storage + version counter), and the forward then simply runs its own prologue."""
        self._engine.prefetch_w(self._param())

    def fused_loss(self, feats: torch.Tensor, labels: torch.Tensor) -> FusedOutput:
        self._check(feats, labels)
        self._pre_forward(feats)
        margins = self._sample_margins(feats, labels)
        self._push_state()
        out = FusedMarginLossFn.apply(feats, self._param(), labels, self._engine, self._mh_state, margins,
                                      self.training or True, torch.is_grad_enabled())
        self._pull_state()
        return FusedOutput(*out)

    def forward(self, feats: torch.Tensor, labels: torch.Tensor):
        self._check(feats, labels)
        self._pre_forward(feats)
        margins = self._sample_margins(feats, labels)
        self._push_state()
        pre, logits, norms, loss_g = DenseMarginLogitsFn.apply(feats, self._param(), labels, self._engine,
                                                               self._mh_state, margins, True)
        self._pull_state()
        one_hot = torch.zeros_like(logits)
        one_hot.scatter_(1, labels.view(-1, 1), 1.0)
        if self.family != "magface":
            loss_g = 0                      # the reference returns the int 0 (criterion.py:107,195,299,449,587,907,1021)
        return [pre, logits], norms, loss_g, one_hot

    def get_proxy(self, labels: torch.Tensor) -> torch.Tensor:
        """Raw class centres for the given labels, [D, N] (CosFace.get_proxy, criterion.py:157-159)."""
        W = self._param()
        return (W[labels].t() if self.layout == "CD" else W[:, labels]).clone().detach()


def _insightface_init(p: torch.Tensor):
    # criterion.py:152,367,833,1218
    p.data.uniform_(-1, 1).renorm_(2, 1, 1e-5).mul_(1e5)


class SphereFace(_MarginHeadBase):
    """criterion.py:12-107."""
    family, layout, param_name = "sphereface", "CD", "weight"

    def __init__(self, in_features: int, out_features: int, device_id=None, m: int = 4):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.in_features, self.out_features, self.m, self.device_id = in_features, out_features, m, device_id
        self.base, self.gamma, self.power, self.LambdaMin, self.iter = 1000.0, 0.12, 1, 5.0, 0
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)
        self._init_engine(out_features, sphere_m=int(m))

    def _pre_forward(self, feats):
        self.iter += 1                                                   # criterion.py:58-60
        self.lamb = max(self.LambdaMin, self.base * (1 + self.gamma * self.iter) ** (-self.power))
        self._engine.cfg.sphere_lambda = self.lamb


class CosFace(_MarginHeadBase):
    """criterion.py:137-197."""
    family, layout, param_name = "cosface", "DC", "kernel"

    def __init__(self, embedding_size=512, classnum=51332, s: float = 64.0, m: float = 0.4):
        super().__init__()
        self.classnum, self.s, self.m, self.eps = classnum, s, m, 1e-4
        self.kernel = nn.Parameter(torch.empty(embedding_size, classnum))
        _insightface_init(self.kernel)
        self._init_engine(classnum, s=s, m=m)


class ArcFace(_MarginHeadBase):
    """criterion.py:232-301."""
    family, layout, param_name = "arcface", "CD", "weight"

    def __init__(self, embed_size, num_classes, device_id=None, s=64.0, m=0.50, easy_margin=True):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.in_features, self.out_features, self.device_id = embed_size, num_classes, device_id
        self.s, self.m, self.easy_margin = s, m, easy_margin
        self.weight = nn.Parameter(torch.empty(num_classes, embed_size))
        nn.init.xavier_uniform_(self.weight)
        self.cos_m, self.sin_m = math.cos(m), math.sin(m)
        self.th, self.mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
        self._init_engine(num_classes, s=s, m=m, easy_margin=int(bool(easy_margin)))


class MV_Softmax(_MarginHeadBase):
    """criterion.py:327-461."""
    layout, param_name = "CD", "weight"

    def __init__(self, feat_dim: int, num_class: int, margin: float = 0.35, mv_weight: float = 1.12,
                 s: float = 32.0, margin_type: str = "arc", device_id=None):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.margin, self.mv_weight, self.s = feat_dim, num_class, margin, mv_weight, s
        self.margin_type = margin_type.lower()
        self.device_id = device_id
        assert self.margin_type in ("am", "arc"), "margin_type must be 'am' or 'arc'"
        self.weight = nn.Parameter(torch.empty(num_class, feat_dim))
        _insightface_init(self.weight)
        if self.margin_type == "arc":
            self.cos_m, self.sin_m = math.cos(margin), math.sin(margin)
            self.th, self.mm = math.cos(math.pi - margin), math.sin(margin) * margin
        self.family = "mv_am" if self.margin_type == "am" else "mv_arc"
        self._init_engine(num_class, s=s, m=margin, mv_weight=mv_weight)

    def _pre_forward(self, feats):
        # evaluate_models.py:50,53 flips margin_type after construction; honour it
        fam = "mv_am" if self.margin_type == "am" else "mv_arc"
        if fam != self._engine.family:
            self.family = fam
            self._engine.family = fam
            self._engine.cfg.family = L.FAMILY[fam]


class CurricularFace(_MarginHeadBase):
    """criterion.py:491-587."""
    family, layout, param_name = "curricularface", "DC", "kernel"

    def __init__(self, feat_dim: int, num_class: int, m: float = 0.5, s: float = 64.0, momentum: float = 0.01):
        super().__init__()
        self.m, self.s, self.momentum = m, s, momentum
        self.cos_m, self.sin_m = math.cos(m), math.sin(m)
        self.threshold, self.mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
        self.kernel = nn.Parameter(torch.empty(feat_dim, num_class))
        nn.init.normal_(self.kernel, std=0.01)
        self.register_buffer("t", torch.zeros(1))
        self._init_engine(num_class, s=s, m=m, momentum=momentum)

    def _push_state(self):
        self._mh_state[0:1].copy_(self.t.to(torch.float32))

    def _pull_state(self):
        # in place (the reference rebinds self.t, criterion.py:572): the buffer keeps its address, so a CUDA-graph replay
        # of the step reads and writes the live state
        with torch.no_grad():
            self.t.copy_(self._mh_state[0:1])


class AdaFace(_MarginHeadBase):
    """criterion.py:795-918."""
    family, layout, param_name = "adaface", "DC", "kernel"

    def __init__(self, feat_dim: int, num_class: int, m: float = 0.4, h: float = 0.333, s: float = 64.0,
                 t_alpha: float = 1.0, device_id=None):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.m, self.h, self.s, self.t_alpha = feat_dim, num_class, m, h, s, t_alpha
        self.device_id, self.eps = device_id, 1e-3
        self.kernel = nn.Parameter(torch.empty(feat_dim, num_class))
        _insightface_init(self.kernel)
        self.register_buffer("t", torch.zeros(1))
        self.register_buffer("batch_mean", torch.ones(1) * 20)
        self.register_buffer("batch_std", torch.ones(1) * 100)
        self._init_engine(num_class, s=s, m=m, h=h, t_alpha=t_alpha)

    def _push_state(self):
        self._mh_state[1:2].copy_(self.batch_mean.to(torch.float32))
        self._mh_state[2:3].copy_(self.batch_std.to(torch.float32))

    def _pull_state(self):
        with torch.no_grad():                          # in place, see CurricularFace._pull_state
            self.batch_mean.copy_(self._mh_state[1:2])
            self.batch_std.copy_(self._mh_state[2:3])


class _ElasticBase(_MarginHeadBase):
    layout, param_name = "DC", "kernel"

    def __init__(self, feat_dim: int, num_class: int, s: float, m: float, std: float, plus: bool, device_id):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.s, self.m, self.std, self.plus = feat_dim, num_class, s, m, std, plus
        self.device_id = device_id
        self.kernel = nn.Parameter(torch.empty(feat_dim, num_class))
        nn.init.normal_(self.kernel, std=0.01)
        self._init_engine(num_class, s=s, m=m, plus=int(bool(plus)))

    def _sample_margins(self, feats, labels):
        if bool((labels < 0).any()):
            raise ValueError("labels == -1 (criterion.py:997) are not supported: CrossEntropyLoss rejects them")
        if self._margins_override is not None:
            return self._margins_override
        # same RNG consumption as criterion.py:1003-1005 / 1116-1118
        mg = torch.normal(mean=self.m, std=self.std, size=(labels.shape[0], 1), device=feats.device)
        return mg.clamp(self.m - self.std, self.m + self.std).squeeze(1)


class ElasticCosFace(_ElasticBase):
    """criterion.py:951-1030."""
    family = "elastic_cos"

    def __init__(self, feat_dim: int, num_class: int, s: float = 64.0, m: float = 0.35, std: float = 0.0125,
                 plus: bool = False, device_id=None):
        super().__init__(feat_dim, num_class, s, m, std, plus, device_id)


class ElasticArcFace(_ElasticBase):
    """criterion.py:1054-1154."""
    family = "elastic_arc"

    def __init__(self, feat_dim: int, num_class: int, s: float = 64.0, m: float = 0.50, std: float = 0.0125,
                 plus: bool = False, device_id=None):
        super().__init__(feat_dim, num_class, s, m, std, plus, device_id)


class MagFace(_MarginHeadBase):
    """criterion.py:1178-1301."""
    family, layout, param_name = "magface", "DC", "kernel"

    def __init__(self, feat_dim: int, num_class: int, s: float = 64.0, easy_margin: bool = True,
                 l_margin: float = 0.45, u_margin: float = 0.8, l_a: float = 10.0, u_a: float = 110.0,
                 device_id=None):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.s, self.easy_margin = feat_dim, num_class, s, easy_margin
        self.l_margin, self.u_margin, self.l_a, self.u_a, self.device_id = l_margin, u_margin, l_a, u_a, device_id
        self.kernel = nn.Parameter(torch.empty(feat_dim, num_class))
        _insightface_init(self.kernel)
        self._init_engine(num_class, s=s, easy_margin=int(bool(easy_margin)), l_margin=l_margin, u_margin=u_margin,
                          l_a=l_a, u_a=u_a)


class VPLArcFace(_MarginHeadBase):
    """criterion.py:619-762 (VPL-ArcFace, virtual proxies from a per-class feature memory).

    Every non-target cosine of the reference, (1 - a_j) x^.w^_j + a_j x^.m^_j with a_j = lamda * 1[life_j > 0], is the
    cosine against the mixed class vector v_j = (1 - a_j) w^_j + a_j m^_j, so the head runs the same fused tensor-core
    pipeline on v (``mh_vpl_mix``); the target column, (1 - a_y) x^.w^_y + a_y, is a per-row term.  Supported through
    ``fused_loss`` on the tensor-core path in stash mode (s <= 69); the materialised 4-tuple is not built for this head.
    """
    family, layout, param_name = "vpl_arcface", "CD", "weight"

    def __init__(self, feat_dim: int, num_class: int, s: float = 64.0, m: float = 0.50, easy_margin: bool = True,
                 lamda: float = 0.15, delta: int = 100, device_id=None):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.s, self.m = feat_dim, num_class, s, m
        self.easy_margin, self.lamda, self.delta, self.device_id = easy_margin, lamda, delta, device_id
        self.weight = nn.Parameter(torch.empty(num_class, feat_dim))
        nn.init.xavier_uniform_(self.weight)
        self.register_buffer("mem", torch.zeros(num_class, feat_dim))
        self.register_buffer("life", torch.zeros(num_class))
        self.register_buffer("cos_m", torch.tensor(math.cos(m), dtype=torch.float32))
        self.register_buffer("sin_m", torch.tensor(math.sin(m), dtype=torch.float32))
        self.register_buffer("th", torch.tensor(math.cos(math.pi - m), dtype=torch.float32))
        self.register_buffer("mm", torch.tensor(math.sin(math.pi - m) * m, dtype=torch.float32))
        self.norm_training_flag = True
        self._init_engine(num_class, s=s, m=m, easy_margin=int(bool(easy_margin)))

    def change_training_mode(self, flag: bool):
        """Toggle memory-based proxy learning (criterion.py:677-679)."""
        self.norm_training_flag = flag

    @torch.no_grad()
    def _update_memory(self, feats: torch.Tensor, labels: torch.Tensor):
        """criterion.py:703-717, without the Python loop over classes: mem[c] = mean of the batch's raw features of class c,
        life[c] = delta for the classes present, then every lifetime decays by one."""
        uniq, inv = torch.unique(labels, return_inverse=True)
        sums = torch.zeros(uniq.numel(), feats.shape[1], dtype=torch.float32, device=feats.device)
        sums.index_add_(0, inv, feats.float())
        cnt = torch.bincount(inv, minlength=uniq.numel()).clamp_min(1).unsqueeze(1)
        mean = sums / cnt
        if feats.dtype != torch.float32:
            mean = mean.to(feats.dtype).float()              # the reference takes the mean in the features' dtype
        self.mem[uniq] = mean
        self.life[uniq] = float(self.delta)
        self.life.sub_(1.0)

    def prefetch(self) -> None:
        """Enqueue the W prologue of the NEXT forward now, on the current stream.  It depends on the class centres only, not
        on the batch, so a training loop can call this before the batch's host-to-device copy or before / beside the
        backbone forward (on a side stream: the prologue is HBM-bound, the backbone tensor-bound) and the head's forward
        then starts with its GEMM.  Good for exactly one forward; any write to the parameter in between invalidates it
        (storage + version counter), and the forward then simply runs its own prologue."""
        self._engine.prefetch_w(self._param())

    def fused_loss(self, feats: torch.Tensor, labels: torch.Tensor) -> FusedOutput:
        self._check(feats, labels)
        if self.norm_training_flag:
            self._update_memory(feats.detach(), labels)
            self._engine.vpl = dict(mem=self.mem, life=self.life, lamda=float(self.lamda))
        else:
            self._engine.vpl = None
        out = FusedMarginLossFn.apply(feats, self._param(), labels, self._engine, self._mh_state, None, True,
                                      torch.is_grad_enabled())
        return FusedOutput(*out)

    def forward(self, feats: torch.Tensor, labels: torch.Tensor):
        raise NotImplementedError("VPLArcFace is available through fused_loss(feats, labels); the materialised "
                                  "[pre, logits] 4-tuple of criterion.py:762 is not built for this head")


class QAFace(_MarginHeadBase):
    """criterion.py:1331-1520 (QAFace: quality-aware injection memory).

    Same skeleton as VPLArcFace with a BINARY class mask: a class whose memory is alive (``life > 0``) is represented
    by its normalised memory vector instead of its centre in every non-target column (criterion.py:1480-1484), i.e. the
    fused GEMM runs on the mixed class vectors of ``mh_vpl_mix`` with lamda = 1.  The target column is
    ``x^ . normalise(w_y + injection)`` with the RAW centre and the quality-gated, magnitude-normalised ``minput`` row
    (1459-1461, 1487-1490): an O(B d) term formed here in differentiable PyTorch and handed to the kernels as an external
    target cosine; the backward returns ``d loss / d t_i``, so gradients reach ``feats``, ``weight`` and ``minput``
    exactly as the reference's autograd routes them.  The ``muy`` / ``std`` statistics are carried detached between
    steps (the reference keeps their graph, so its own second backward fails when ``minput`` requires grad).
    ``fused_loss(feats, minput, labels)`` only: the materialised 4-tuple of criterion.py:1520 is not built.
    """
    family, layout, param_name = "vpl_arcface", "CD", "weight"

    def __init__(self, feat_dim: int, num_class: int, s: float = 64.0, m: float = 0.50, easy_margin: bool = True,
                 delta: int = 1000, tto: float = 2.0, alpha: float = 0.99, device_id=None):
        super().__init__()
        if device_id is not None:
            raise NotImplementedError("device_id model-parallel is replaced by ShardedMarginHead")
        self.feat_dim, self.num_class, self.s, self.m, self.easy_margin = feat_dim, num_class, s, m, easy_margin
        self.delta, self.tto, self.alpha, self.device_id = delta, tto, alpha, device_id
        self.weight = nn.Parameter(torch.empty(num_class, feat_dim))
        nn.init.xavier_uniform_(self.weight)
        self.register_buffer("mem", torch.zeros(num_class, feat_dim))
        self.register_buffer("life", torch.zeros(num_class))
        self.register_buffer("muy", torch.tensor(0.0))
        self.register_buffer("std", torch.tensor(1.0))
        self.register_buffer("cos_m", torch.tensor(math.cos(m), dtype=torch.float32))
        self.register_buffer("sin_m", torch.tensor(math.sin(m), dtype=torch.float32))
        self.register_buffer("th", torch.tensor(math.cos(math.pi - m), dtype=torch.float32))
        self.register_buffer("mm", torch.tensor(math.sin(math.pi - m) * m, dtype=torch.float32))
        self.norm_training_flag = True
        self._init_engine(num_class, s=s, m=m, easy_margin=int(bool(easy_margin)))

    def change_training_mode(self, flag: bool):
        """Toggle quality-aware memory injection (criterion.py:1399-1401)."""
        self.norm_training_flag = flag

    def injection_cal(self, norm_mag_minput: torch.Tensor) -> torch.Tensor:
        """exp(-z) where |z| < tto, else 0 (criterion.py:1409-1413)."""
        f = torch.exp(-norm_mag_minput)
        return torch.where(torch.abs(norm_mag_minput) < self.tto, f, torch.zeros_like(f))

    def fused_loss(self, feats: torch.Tensor, minput: torch.Tensor, labels: torch.Tensor) -> FusedOutput:
        self._check(feats, labels)
        if not self.norm_training_flag:
            self._engine.vpl = None                                        # plain clamped ArcFace on the class centres
            out = FusedMarginLossFn.apply(feats, self.weight, labels, self._engine, self._mh_state, None, True,
                                          torch.is_grad_enabled())
            return FusedOutput(*out)
        if minput.shape != feats.shape:
            raise ValueError("feats / minput shape mismatch")
        mi = minput.float()
        mag = torch.norm(mi, p=2, dim=1, keepdim=True)                       # criterion.py:1447
        mean, sd = mag.mean(), mag.std()
        first = self.muy == 0.0                                              # 1451: decided on the device, no host sync
        muy = torch.where(first, mean, self.alpha * self.muy + (1 - self.alpha) * mean)
        std = torch.where(first, sd, self.alpha * self.std + (1 - self.alpha) * sd)
        self.muy, self.std = muy.detach(), std.detach()
        z = ((mag - muy) / (std + 1e-6)).squeeze(1)                          # 1459
        injection = self.injection_cal(z).unsqueeze(1) * mi / (mag + 1e-6)   # 1460-1461
        with torch.no_grad():                                                # 1464-1477 without the loop over classes
            uniq, inv = torch.unique(labels, return_inverse=True)
            sums = torch.zeros(uniq.numel(), mi.shape[1], dtype=torch.float32, device=mi.device)
            sums.index_add_(0, inv, injection.detach())
            cnt = torch.bincount(inv, minlength=uniq.numel()).clamp_min(1).unsqueeze(1)
            self.mem[uniq] = sums / cnt
            self.life[uniq] = float(self.delta)
            self.life.sub_(1.0)
        xh = torch.nn.functional.normalize(feats.float(), dim=1)
        tw = torch.nn.functional.normalize(self.weight[labels] + injection, dim=1)     # 1487-1488: raw centre + injection
        t_ext = (xh * tw).sum(dim=1)                                                    # 1489
        self._engine.vpl = dict(mem=self.mem, life=self.life, lamda=1.0)               # a_j = 1[life_j > 0]
        out = FusedMarginLossFn.apply(feats, self.weight, labels, self._engine, self._mh_state, None, True,
                                      torch.is_grad_enabled(), t_ext)
        return FusedOutput(*out)

    def forward(self, feats: torch.Tensor, minput: torch.Tensor, labels: torch.Tensor):
        raise NotImplementedError("QAFace is available through fused_loss(feats, minput, labels); the materialised "
                                  "[pre, logits] 4-tuple of criterion.py:1520 is not built for this head")


HEAD_CLASSES = dict(
    arcface=ArcFace, cosface=CosFace, sphereface=SphereFace, mv_am=MV_Softmax, mv_arc=MV_Softmax,
    curricularface=CurricularFace, adaface=AdaFace, elastic_cos=ElasticCosFace, elastic_arc=ElasticArcFace,
    magface=MagFace, vpl_arcface=VPLArcFace,
)
# QAFace takes a second feature tensor (fused_loss(feats, minput, labels)) and is therefore not in HEAD_CLASSES
