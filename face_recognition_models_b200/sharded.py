"""Class-sharded (Partial-FC style) margin head: one process per GPU, NCCL over NVLink.

The class-centre matrix is split along C across the ranks of ``group`` (the reference's dead
``device_id`` path chunks W the same way, criterion.py:271, but copies tensors every step and has no
backward-side collective).  Per step and rank:

  1. all-gather  x, labels                      -> every rank sees the global batch  B_g = R * B
  2. all-reduce  (SUM) of the target cosines    -> owner rank contributes, others 0 (thresholds/EMA need all rows)
  3. local fused forward on [B_g x C/R]         -> per-row (max, sum-exp, rank-count, e*u)
  4. all-gather  of the row statistics + merge  -> global log-sum-exp; loss = mean over B_g
  5. local fused backward                       -> dW for the local shard (no collective: W is model-parallel)
  6. reduce-scatter (SUM) of dx^ [B_g, 512]     -> each rank's own B rows, then normalise-backward

Scale convention: the loss is the mean over the GLOBAL batch.  dW uses 1/B_g.  The returned dx is
d(global-mean loss)/dx_local multiplied by ``dx_scale`` (default 1).  When the backbone is wrapped in
DDP (which averages parameter gradients over ranks) pass ``dx_scale=world_size`` so the backbone
gradient equals the single-process gradient on the concatenated batch (InsightFace Partial-FC convention).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L
from .functional import HeadEngine, ShardInfo
from .heads import HEAD_CLASSES, FusedOutput


def shard_range(num_classes: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the classes owned by ``rank``: equal chunks of ceil(C/R), last one ragged."""
    per = (num_classes + world - 1) // world
    b = min(num_classes, rank * per)
    e = min(num_classes, b + per)
    return b, e


class ShardComm:
    """The four exchanges of the sharded head, backend-agnostic (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._backend = dist.get_backend(group) if dist.is_initialized() else "none"

    def gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """all-gather along dim 0 (x [B,512] -> [R*B,512]; labels [B] -> [R*B])."""
        if self.world == 1:
            return t
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather_stats(self, stats: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[planes, B_pad] per rank -> [R, planes, B_pad]."""
        if self.world == 1:
            return stats.unsqueeze(0)
        if out is None:
            out = torch.empty((self.world,) + tuple(stats.shape), dtype=stats.dtype, device=stats.device)
        # concatenation along dim 0 is what every backend accepts: view [R, planes, B] as [R*planes, B]
        dist.all_gather_into_tensor(out.view(-1, stats.shape[-1]), stats.contiguous(), group=self.group)
        return out

    def reduce_scatter_rows(self, full: torch.Tensor) -> torch.Tensor:
        """SUM over ranks of [R*B, 512], returning this rank's [B, 512] slice."""
        if self.world == 1:
            return full
        Bl = full.shape[0] // self.world
        if self._backend == "gloo":           # gloo has no reduce_scatter: all-reduce then slice
            f = full.clone()
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)
            return f[self.rank * Bl:(self.rank + 1) * Bl].contiguous()
        out = torch.empty((Bl,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
        return out


class _ShardedFusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, W_local, labels_local, head, margins_local, grad_enabled=True, y_g=None, engine=None):
        # y_g / engine: Partial-FC sampling hands in the gathered labels (already remapped to the sampled class ids) and
        # the engine sized for the sampled sub-matrix; W_local is then the gathered sub-matrix W[index]
        comm: ShardComm = head.comm
        engine = engine or head.engine
        engine.prefetch_w(W_local)                   # the GPU normalises the shard while the host issues the gathers
        x_g = comm.gather_rows(x_local.contiguous())
        if y_g is None:
            y_g = comm.gather_rows(labels_local.contiguous().to(torch.int64))
        margins_g = comm.gather_rows(margins_local.contiguous()) if margins_local is not None else None
        c = engine.forward(x_g, W_local, y_g, head._mh_state, margins_g, update_state=True,
                           want_grad=bool(grad_enabled and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])))
        ctx.engine = engine
        ctx.head = head
        ctx.c = c
        ctx.set_materialize_grads(False)        # None instead of zero-filled tensors for the outputs nobody differentiates
        ctx.B_local = x_local.shape[0]
        sc = c["scalars"]
        r0 = comm.rank * ctx.B_local
        norms = c["rowp"][L.RP["NORMS"], r0:r0 + ctx.B_local].clone().unsqueeze(1)
        loss, loss_g, acc1, acc5 = sc[0].clone(), head._mh_state[3].clone(), sc[1].clone(), sc[2].clone()
        ctx.mark_non_differentiable(norms, acc1, acc5)
        return loss, loss_g, acc1, acc5, norms

    @staticmethod
    def backward(ctx, g_loss, g_lossg, _a1, _a5, _n):
        head = ctx.head
        dx, dW = ctx.engine.backward(ctx.c, g_loss, g_lossg, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        if dx is not None and head.dx_scale != 1.0:
            dx = dx * head.dx_scale
        return dx, dW, None, None, None, None, None, None


class ShardedMarginHead(nn.Module):
    """Any of the reference head families with the class dimension sharded over ``group``.

    ``ctor_kwargs`` are the reference constructor's keyword arguments other than the two size
    arguments (e.g. ``s=64.0, m=0.5, easy_margin=False`` for ArcFace).  The local parameter keeps the
    reference's name and layout (``weight [C_local, D]`` / ``kernel [D, C_local]``).
    """

    def __init__(self, family: str, num_classes: int, group=None, mode: str = "tc", dx_scale: float = 1.0,
                 sample_rate: float = 1.0, **ctor_kwargs):
        super().__init__()
        self.family = family
        self.num_classes = num_classes
        self.comm = ShardComm(group)
        self.dx_scale = float(dx_scale)
        if not 0.0 < sample_rate <= 1.0:
            raise ValueError("sample_rate must be in (0, 1]")
        self.sample_rate = float(sample_rate)
        b, e = shard_range(num_classes, self.comm.world, self.comm.rank)
        smallest = min(e_ - b_ for b_, e_ in (shard_range(num_classes, self.comm.world, r) for r in range(self.comm.world)))
        if smallest < 2:            # the same verdict on every rank (a rank raising alone would hang the others)
            raise ValueError(f"ShardedMarginHead: {num_classes} classes in chunks of ceil(C/R) over {self.comm.world} ranks "
                             f"leave a rank with {smallest} classes; every rank needs at least 2 - use fewer ranks")
        self.c_begin, self.c_end = b, e
        cls = HEAD_CLASSES[family]
        if family in ("mv_am", "mv_arc"):
            ctor_kwargs = dict(ctor_kwargs, margin_type="am" if family == "mv_am" else "arc")
        # build the local shard with the reference initialisation, then rebind its engine to the shard
        self.local = cls(512, e - b, **ctor_kwargs)
        self.local._engine.shard = ShardInfo(comm=self.comm, rank=self.comm.rank, world=self.comm.world, c_offset=b,
                                             c_total=num_classes)
        self.local._engine.mode = mode
        self.engine: HeadEngine = self.local._engine
        # Partial-FC negative sampling: every step each rank keeps the classes of its shard that
        # occur in the global batch plus random negatives, num_sample = sample_rate * ceil(C / R) per rank, and the whole
        # fused pipeline runs on the gathered sub-matrix through a second engine of that size
        self.num_sample = 0
        self.sub_engine = None
        self.last_index = None
        if self.sample_rate < 1.0:
            per = (num_classes + self.comm.world - 1) // self.comm.world
            self.num_sample = max(2, int(self.sample_rate * per))
            if self.num_sample > smallest:
                raise ValueError(f"sample_rate {sample_rate} asks for {self.num_sample} classes per rank but the smallest shard "
                                 f"holds {smallest}")
            self.sub_engine = HeadEngine(self.engine.family, self.local.layout, self.num_sample, {}, mode=mode,
                                         shard=ShardInfo(comm=self.comm, rank=self.comm.rank, world=self.comm.world,
                                                         c_offset=self.comm.rank * self.num_sample,
                                                         c_total=self.comm.world * self.num_sample))
            self.sub_engine.cfg = self.engine.cfg              # same hyper-parameter block (updated in place by the head)

    @property
    def _mh_state(self):
        return self.local._mh_state

    def shard_parameter(self) -> torch.Tensor:
        return self.local._param()

    def head_parameter(self) -> torch.Tensor:      # optim.HeadSGD protocol
        return self.shard_parameter()

    def head_engine(self) -> HeadEngine:
        return self.engine

    def fused_loss(self, feats: torch.Tensor, labels: torch.Tensor) -> FusedOutput:
        self.local._check(feats, labels)
        self.local._pre_forward(feats)
        margins = self.local._sample_margins(feats, labels)
        self.local._push_state()
        if self.family == "vpl_arcface":
            self._vpl_prepare(feats, labels)
        if self.sub_engine is None:
            out = _ShardedFusedLossFn.apply(feats, self.local._param(), labels, self, margins, torch.is_grad_enabled())
        else:
            self.sub_engine.backward_mode = self.engine.backward_mode
            y_g = self.comm.gather_rows(labels.contiguous().to(torch.int64))
            index, y_sub = self.sample_classes(y_g)
            W = self.local._param()
            W_sub = W.index_select(0 if self.local.layout == "CD" else 1, index)      # autograd scatters dW_sub back into dW
            out = _ShardedFusedLossFn.apply(feats, W_sub, labels, self, margins, torch.is_grad_enabled(), y_sub,
                                            self.sub_engine)
        self.local._pull_state()
        return FusedOutput(*out)

    @torch.no_grad()
    def _vpl_prepare(self, feats: torch.Tensor, labels: torch.Tensor):
        """VPL-ArcFace on class shards (criterion.py:703-717): the memory bank and the lifetimes are sharded like the class
        centres; every rank refreshes the entries of ITS classes from the gathered batch, then decays its lifetimes."""
        if self.sub_engine is not None:
            raise L.MarginHeadError("VPLArcFace does not combine with Partial-FC sampling")
        loc = self.local
        if not loc.norm_training_flag:
            self.engine.vpl = None
            return
        x_g = self.comm.gather_rows(feats.detach().contiguous())
        y_g = self.comm.gather_rows(labels.contiguous().to(torch.int64)) - self.c_begin
        n_local = self.c_end - self.c_begin
        owned = (y_g >= 0) & (y_g < n_local)
        x_o, y_o = x_g[owned], y_g[owned]                        # this shard's rows of the global batch (may be empty)
        if y_o.numel() > 0:
            uniq, inv = torch.unique(y_o, return_inverse=True)
            sums = torch.zeros(uniq.numel(), x_o.shape[1], dtype=torch.float32, device=x_o.device)
            sums.index_add_(0, inv, x_o.float())
            mean = sums / torch.bincount(inv, minlength=uniq.numel()).clamp_min(1).unsqueeze(1)
            if x_o.dtype != torch.float32:
                mean = mean.to(x_o.dtype).float()                # the reference takes the mean in the features' dtype
            loc.mem[uniq] = mean
            loc.life[uniq] = float(loc.delta)
        loc.life.sub_(1.0)
        self.engine.vpl = dict(mem=loc.mem, life=loc.life, lamda=float(loc.lamda))

    @torch.no_grad()
    def sample_classes(self, y_g: torch.Tensor):
        """Partial-FC sampling on the device (no host sync): returns (index, y_sub).

        index [num_sample], sorted local class ids of this rank's shard: every class of the shard that occurs in the
        global batch y_g plus uniformly random negatives.  y_sub [B_g]: the labels in the sampled id space of the whole
        head, rank r owning [r * num_sample, (r + 1) * num_sample) - what the engine of the sub-matrix expects.  More
        distinct positives than num_sample in one shard poisons the loss (NaN) instead of dropping targets silently."""
        n_local = self.c_end - self.c_begin
        k = self.num_sample
        dev = y_g.device
        loc = y_g - self.c_begin
        owned = (loc >= 0) & (loc < n_local)
        perm = torch.rand(n_local, device=dev)
        # positives first (insightface partial_fc.sample: perm[positive] = 2.0), as a scatter-max so that rows of other
        # shards (clamped index, value 0) change nothing and no boolean indexing forces a host sync
        perm.scatter_reduce_(0, loc.clamp(0, n_local - 1), torch.where(owned, 2.0, 0.0).to(perm.dtype), reduce="amax")
        index = torch.topk(perm, k)[1].sort()[0]
        pos = torch.searchsorted(index, loc.clamp(0, n_local - 1))
        hit = owned & (index[pos.clamp(max=k - 1)] == loc)        # a positive that did not fit is not "hit"
        # rows owned by this rank carry their new id, all others 0; an owned positive that was dropped carries an id outside
        # [0, world * k) so that the prologue poisons the row (NaN loss) on every rank
        new_id = torch.where(hit, pos + self.comm.rank * k, torch.zeros_like(pos))
        new_id = torch.where(owned & ~hit, torch.full_like(pos, self.comm.world * k), new_id)
        bad = (y_g < 0) | (y_g >= self.num_classes)               # never owned by anyone: poison as well
        new_id = torch.where(bad & (self.comm.rank == 0), torch.full_like(pos, self.comm.world * k), new_id)
        self.comm.allreduce_sum_(new_id)
        self.last_index = index
        return index, new_id

    forward = fused_loss

    # ---- checkpoint interchange with the unsharded reference head ---------------------------
    def _class_major(self, w: torch.Tensor) -> torch.Tensor:
        """View a parameter (or shard) with the class index first: [C, D] as is, [D, C] transposed."""
        return w if self.local.layout == "CD" else w.t()

    @torch.no_grad()
    def load_full_parameter(self, full: torch.Tensor):
        """Scatter a full reference parameter (``weight [C, D]`` / ``kernel [D, C]``, e.g. ``checkpoint['model_state_dict']
        ['arcface.weight']`` saved by model_utils.py:43-60) into this rank's shard.  Every rank passes the same tensor."""
        C_ = self.num_classes
        want = (C_, L.D) if self.local.layout == "CD" else (L.D, C_)
        if tuple(full.shape) != want:
            raise ValueError(f"expected the full parameter of shape {want}, got {tuple(full.shape)}")
        shard = self._class_major(full)[self.c_begin:self.c_end]
        self._class_major(self.shard_parameter().data).copy_(shard.to(self.shard_parameter().device))

    @torch.no_grad()
    def gather_full_parameter(self) -> torch.Tensor:
        """All-gather the shards back into the reference parameter layout (on every rank): what the unsharded head's
        ``state_dict()`` holds under ``weight`` / ``kernel``, so a checkpoint written from it loads into the reference."""
        w = self._class_major(self.shard_parameter().data).contiguous()
        if self.comm.world == 1:
            full = w.clone()
        else:
            per = (self.num_classes + self.comm.world - 1) // self.comm.world
            pad = torch.zeros((per, L.D), dtype=w.dtype, device=w.device)
            pad[:w.shape[0]] = w                                   # the last shard may be ragged
            out = torch.empty((self.comm.world * per, L.D), dtype=w.dtype, device=w.device)
            dist.all_gather_into_tensor(out, pad, group=self.comm.group)
            full = out[:self.num_classes]
        return full if self.local.layout == "CD" else full.t().contiguous()

    def full_state_dict(self):
        """``state_dict`` of the equivalent unsharded reference head: the gathered parameter plus the head's buffers
        (CurricularFace ``t``, AdaFace ``batch_mean`` / ``batch_std``; identical on every rank)."""
        sd = {self.local.param_name: self.gather_full_parameter()}
        for k, v in self.local.state_dict().items():
            if k != self.local.param_name:
                sd[k] = v.clone()
        return sd

    def load_full_state_dict(self, sd):
        self.load_full_parameter(sd[self.local.param_name])
        with torch.no_grad():
            for k, v in sd.items():
                if k != self.local.param_name:
                    getattr(self.local, k).copy_(v)
