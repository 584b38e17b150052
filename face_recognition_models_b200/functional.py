"""Host-side engine of the margin head: workspaces, kernel sequencing, autograd glue.

Two execution paths, both CUDA only (there is no CPU / PyTorch fallback):

* ``tc``    - bf16 tensor-core path (tcgen05 + TMEM + TMA).  A forward that needs no gradient never
              materialises B x C.  For training there are two backward modes:
              ``recompute`` - backward recomputes the logit tiles and keeps only G = (P-Y) dz/dcos in bf16;
              ``stash``     - the forward additionally writes the unnormalised probabilities
                              E' = exp2(z - ref) du/dcos in bf16 (fixed softmax reference), and the backward
                              runs only the two gradient GEMMs (3 instead of 4 GEMM passes per step).
              ``auto`` (default) = stash whenever ``mh_tc_stash_ok`` holds for the head; else, at scale, the GUARDED
                              stash (``mh_tc_stash_guarded_ok``: CurricularFace, SphereFace, s > 69 -- the same
                              forward + stash run speculatively, a device-side check of the row sums, and the
                              general path behind it as gated launches); else recompute.
* ``exact`` - fp32 SIMT path that materialises the cosine matrix (small C: fp32-tolerance mode,
              and the compat mode returning the reference's 4-tuple).

The sequence mirrors the reference forward (criterion.py, any head) + ``nn.CrossEntropyLoss`` +
``accuracy`` of model_utils.py:176-182, and autograd's backward of all of it.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    # the raw cudaStream_t of the current stream of the current device (the engine methods run under the data's device);
    # torch.cuda.current_stream() builds a Stream object through several Python layers: ~20 us per call on the hot path
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _on_device_of(argpos: int, same_device=(0, 1, 2, 3)):
    """Run the method with the device of its ``argpos``-th positional tensor as the current CUDA device, so the stream
    handed to the C ABI, the workspaces and every launch belong to the GPU that holds the data (a head on cuda:1 works
    while cuda:0 is current, as the reference's modules do).  The tensor arguments at positions ``same_device`` (x / W /
    labels / state, or W / grad / momentum) must live on that device."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(self, *args, **kw):
            t = args[argpos]
            if isinstance(t, dict):
                t = t["x_hat"]
            if not t.is_cuda:
                raise L.MarginHeadError("margin head tensors must be CUDA tensors (no CPU fallback)")
            for a in [args[i] for i in same_device if i < len(args)]:
                if isinstance(a, torch.Tensor) and a.device != t.device:
                    raise L.MarginHeadError(f"margin head tensors must share one device: {a.device} vs {t.device}")
            if torch.cuda.current_device() == t.device.index:
                return fn(self, *args, **kw)
            with torch.cuda.device(t.device):
                return fn(self, *args, **kw)
        return wrapped
    return deco


def _selfproj() -> bool:
    """MH_DW_SELFPROJ=1 selects the self-projecting dW kernel (mh_tc_backward_dw_proj) and a dx GEMM without side pass."""
    return os.environ.get("MH_DW_SELFPROJ", "0") == "1"


def _merged_bwd() -> bool:
    """The merged dx + dW kernel (mh_tc_backward_dxdw) runs wherever the shape is eligible (>= 8 class tiles per CTA pair)
    and both gradients are wanted: 0.4 ms of 7.6 per cfg4 step, 4.7 GB less DRAM traffic.  MH_BWD_MERGED=0 restores the
    two back-to-back kernels (no cross-CTA waits at all)."""
    return os.environ.get("MH_BWD_MERGED", "1") != "0"


def _merged_fwd() -> bool:
    """MH_FWD_MERGED=1 selects the merged W prologue + forward kernel (mh_tc_forward_pw) where head and shape are eligible."""
    return os.environ.get("MH_FWD_MERGED", "0") == "1"


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


_DT = {torch.float32: L.DT_F32, torch.bfloat16: L.DT_BF16, torch.float16: L.DT_F16}
GUARDED_MIN_BC = 1 << 25        # smallest B_pad * C_pad at which backward_mode = 'auto' takes the guarded stash
_DEVICE_OK = set()


def sphere_lambda(it: int) -> float:
    """SphereFace annealing, criterion.py:60 (base 1000, gamma 0.12, power 1, LambdaMin 5)."""
    return max(5.0, 1000.0 * (1 + 0.12 * it) ** (-1))


@dataclass
class ShardInfo:
    """Class-dimension sharding (Partial-FC style): this rank owns [c_offset, c_offset + C_local).

    ``comm`` is a ``sharded.ShardComm`` (the collective plumbing; None when world == 1)."""
    comm: object = None
    rank: int = 0
    world: int = 1
    c_offset: int = 0
    c_total: int = 0              # class count of the whole head (0: unsharded, = the engine's C)


class HeadEngine:
    """Owns the device workspaces of one head and runs the kernel pipeline through the C ABI."""

    def __init__(self, family: str, layout: str, num_classes_local: int, cfg_kwargs: Dict,
                 mode: str = "tc", shard: Optional[ShardInfo] = None):
        assert family in L.FAMILY, family
        assert layout in ("CD", "DC")
        assert mode in ("tc", "exact")
        self.family = family
        self.layout = L.LAYOUT_CD if layout == "CD" else L.LAYOUT_DC
        self.C = int(num_classes_local)
        self.mode = mode
        self.shard = shard or ShardInfo()
        self.cfg = L.MhConfig()
        self.cfg.family = L.FAMILY[family]
        self.cfg.s = 64.0
        for k, v in cfg_kwargs.items():
            setattr(self.cfg, k, v)
        self._ws: Dict[str, torch.Tensor] = {}
        self._gen = 0
        self._shadow = None             # (W.data_ptr(), W._version, w_hat.data_ptr()) when sgd_step left a valid w_hat behind
        self._shadow_once = False       # set by prefetch_w: good for the next forward only (never a cross-step cache)
        self.vpl = None                 # VPLArcFace: dict(mem, life, lamda) set by the head before each forward
        self._graph_cache = None        # handle of mh_step_cache_create (CUDA-graph replay of the two phases)
        self._step_key = None           # cached mh_step_ws descriptor of the whole-phase entry points (see _forward_step)
        self._step_ws = self._step_T = None
        self._stash_ok_key = None
        # "auto" | "stash" | "recompute"; MH_BACKWARD in the environment overrides (A/B measurements)
        self.backward_mode = os.environ.get("MH_BACKWARD", "auto")

    # copy.deepcopy(head) (EMA copies, checkpoints of the module object) and pickling: the copy keeps the hyper-parameters
    # and modes but none of the device workspaces, raw-pointer descriptors or the CUDA-graph cache of the original (they
    # would alias the original's buffers, and the cache handle would be destroyed twice); it rebuilds its own on first use
    def __getstate__(self):
        d = self.__dict__.copy()
        d.update(_ws={}, _shadow=None, _shadow_once=False, _step_key=None, _step_ws=None, _step_T=None, _graph_cache=None,
                 _stash_ok_key=None, vpl=None, _gen=0)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)

    def stash_ok(self) -> bool:
        """True when the forward may stash E' for the backward (fixed-reference softmax applies)."""
        if self.mode != "tc" or self.backward_mode == "recompute":
            return False
        ok = bool(L.load().mh_tc_stash_ok(C.byref(self.cfg), self.C))
        if self.backward_mode == "stash" and not ok and not L.load().mh_tc_stash_guarded_ok(C.byref(self.cfg), self.C):
            raise L.MarginHeadError("backward_mode='stash' needs a head that passes mh_tc_stash_ok (fixed scale with "
                                    "s*log2(e)*2 <= 200, no hard-negative re-weighting) or mh_tc_stash_guarded_ok")
        return ok

    # -- workspace ------------------------------------------------------------------------------
    def _buf(self, name: str, shape, dtype, device, zero: bool = False) -> torch.Tensor:
        key = name
        t = self._ws.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self._ws[key] = t
        return t

    def release_workspaces(self):
        self._ws.clear()
        self._shadow = None
        self._step_key = self._step_ws = self._step_T = None
        self._drop_graph_cache()

    def _graph_cache_handle(self):
        """CUDA-graph cache of the two phases (mh_step_cache_create): the whole-phase entry points replay a phase whose
        arguments they have seen before with one cudaGraphLaunch.  MH_STEP_GRAPH=0 keeps plain launches."""
        if os.environ.get("MH_STEP_GRAPH", "1") == "0":
            return None
        if self._graph_cache is None:
            h = C.c_void_p(0)
            L.call("mh_step_cache_create", C.byref(h))
            self._graph_cache = h
        return self._graph_cache.value

    def graph_stats(self):
        """(replayed, captured, plain) phase counts of this engine's CUDA-graph cache, or None without one."""
        if self._graph_cache is None:
            return None
        out = (C.c_int64 * 3)()
        L.call("mh_step_cache_stats", self._graph_cache, out)
        return tuple(int(v) for v in out)

    def _drop_graph_cache(self):
        h, self._graph_cache = getattr(self, "_graph_cache", None), None
        if h is not None and h.value:
            try:
                L.call("mh_step_cache_destroy", h)
            except Exception:                                   # interpreter shutdown: the library may already be gone
                pass

    def __del__(self):
        self._drop_graph_cache()

    def _stash_ok_cached(self) -> bool:
        """stash_ok() depends only on the hyper-parameters, the shard size and backward_mode: asked once per change."""
        key = (self.backward_mode, self.mode, self.family, float(self.cfg.s), float(self.cfg.mv_weight), self.C)
        if self._stash_ok_key != key:
            self._stash_ok_val = self.stash_ok()
            self._stash_guard_val = bool(self.mode == "tc" and self.backward_mode != "recompute" and not self._stash_ok_val
                                         and L.load().mh_tc_stash_guarded_ok(C.byref(self.cfg), self.C))
            self._stash_ok_key = key
        return self._stash_ok_val

    def _stash_kind(self, B_pad: int, C_pad: int) -> int:
        """0 recompute, 1 the proven stash (mh_tc_stash_ok), 2 the guarded stash of the step API (CurricularFace, SphereFace,
        s > 69: fixed-reference forward + stash run speculatively, a device-side check of the row sums decides whether the
        general path has to re-run; mh_tc_stash_guarded_ok).  'auto' takes the guarded stash where a whole GEMM pass costs more
        than its handful of gated launches (B_pad * C_pad >= 2^25); backward_mode = 'stash' takes it at any size."""
        if self._stash_ok_cached():
            return 1
        if self.family == "vpl_arcface":           # the memory-bank heads are built on the proven stash only
            return 0
        if self._stash_guard_val and os.environ.get("MH_STASH_GUARDED", "1") != "0":
            if self.backward_mode == "stash" or B_pad * C_pad >= GUARDED_MIN_BC:
                return 2
        return 0

    def invalidate_shadow(self):
        """Forget the w_hat left by sgd_step().  Only needed after writing the parameter behind autograd's back
        (`W.data.<op>_()`, raw pointers): such writes do not move W's version counter."""
        self._shadow = None

    # -- fused optimizer step ---------------------------------------------------------------------
    @_on_device_of(0)
    def sgd_step(self, W: torch.Tensor, grad: torch.Tensor, momentum_buf: torch.Tensor, lr: float, momentum: float,
                 weight_decay: float, grad_scale: Optional[torch.Tensor] = None,
                 found_inf: Optional[torch.Tensor] = None) -> None:
        """SGD-momentum update of the class-centre parameter (model_utils.py:557, 186) in one pass that also writes the
        next step's w_hat / inv_norm, so the next tensor-core forward of this head skips mh_prologue_w.  The shadow is
        tied to W's storage and version counter: any other in-place write to W invalidates it."""
        if not W.is_cuda:
            raise L.MarginHeadError("margin head parameters must be CUDA tensors (no CPU fallback)")
        assert W.dtype == torch.float32 and W.is_contiguous()
        assert grad.dtype == torch.float32 and grad.is_contiguous() and grad.shape == W.shape, "grad must match W"
        assert momentum_buf.dtype == torch.float32 and momentum_buf.is_contiguous() and momentum_buf.shape == W.shape
        for t in (grad_scale, found_inf):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.numel() == 1)
        dev = W.device
        Cn = self.C
        C_pad = _round_up(Cn, L.NTILE)
        w_hat = self._buf("w_hat", (C_pad, L.D), torch.bfloat16, dev)
        inv_norm = self._buf("inv_norm", (Cn,), torch.float32, dev)
        self._gen += 1                  # a forward context that has not run its backward yet loses its w_hat
        L.call("mh_sgd_step_w", _ptr(W), self.layout, Cn, W.shape[1], _ptr(grad), _ptr(momentum_buf), C.c_float(lr),
               C.c_float(momentum), C.c_float(weight_decay), _ptr(grad_scale), _ptr(found_inf), _ptr(w_hat), C_pad,
               _ptr(inv_norm), _stream())
        torch.autograd.graph.increment_version(W)
        self._shadow = (W.data_ptr(), W._version, w_hat.data_ptr())
        self._shadow_once = False

    @_on_device_of(0)
    def prefetch_w(self, W: torch.Tensor) -> None:
        """Launch the W prologue ahead of the rest of the forward: it does not depend on the batch, so the sharded head
        calls this before its all-gathers and the GPU has 0.1-0.9 ms of work while the host issues the collectives.
        forward() then finds w_hat valid for this W (same mechanism as the shadow left by sgd_step)."""
        if self.mode == "exact" or W.dtype != torch.float32 or not W.is_contiguous():
            return
        dev = W.device
        Cn = self.C
        C_pad = _round_up(Cn, L.NTILE)
        w_hat = self._buf("w_hat", (C_pad, L.D), torch.bfloat16, dev)
        inv_norm = self._buf("inv_norm", (Cn,), torch.float32, dev)
        if self._shadow == (W.data_ptr(), W._version, w_hat.data_ptr()):
            return
        L.call("mh_prologue_w", _ptr(W), self.layout, Cn, W.shape[1], _ptr(w_hat), C_pad, _ptr(None), _ptr(inv_norm), _stream())
        self._shadow = (W.data_ptr(), W._version, w_hat.data_ptr())
        self._shadow_once = True

    # -- forward ----------------------------------------------------------------------------------
    @_on_device_of(0)
    def forward(self, x: torch.Tensor, W: torch.Tensor, labels: torch.Tensor, state: torch.Tensor,
                margins: Optional[torch.Tensor], update_state: bool = True, want_dense: bool = False,
                want_grad: bool = False, t_ext: Optional[torch.Tensor] = None):
        """Runs prologues + fused forward.  x is the (already gathered) global batch [B, 512].
        want_grad: a backward will follow, so the forward may stash E' (see the module docstring).

        t_ext: externally computed target cosines [B] (QAFace: the target column is x^ . normalise(w_y + injection),
        criterion.py:1487-1490, formed differentiably by the caller); they replace the prologue's <x^, w^_y> and the
        backward hands d loss / d t_ext back instead of applying the target column to dx / dW.

        Returns a context dict with every tensor the backward needs plus the user-visible outputs.
        """
        lib = L.load()
        if not x.is_cuda:
            raise L.MarginHeadError("margin head inputs must be CUDA tensors (no CPU fallback)")
        if x.device.index not in _DEVICE_OK:
            L.check(lib.mh_device_check(), "mh_device_check")        # once per device: compute capability 10.x
            _DEVICE_OK.add(x.device.index)
        assert x.dim() == 2 and x.shape[1] == L.D, "embedding dimension must be 512"
        assert x.dtype in _DT, x.dtype
        assert W.dtype == torch.float32 and W.is_contiguous()
        x = x.contiguous()
        labels = labels.contiguous().to(torch.int64)
        dev = x.device
        B = x.shape[0]
        Cn = self.C
        B_pad = _round_up(B, 2 * L.TILE)     # 256: one cta_group::2 pair tile of rows
        C_pad = _round_up(Cn, L.NTILE)
        ld = W.shape[1]
        if self.layout == L.LAYOUT_CD:
            assert tuple(W.shape) == (Cn, L.D), (W.shape, Cn)
        else:
            assert tuple(W.shape) == (L.D, Cn), (W.shape, Cn)
        st = _stream()
        exact = self.mode == "exact" or want_dense
        self._gen += 1

        w_hat = self._buf("w_hat", (C_pad, L.D), torch.bfloat16, dev)
        inv_norm = self._buf("inv_norm", (Cn,), torch.float32, dev)
        w_hat32 = self._buf("w_hat32", (Cn, L.D), torch.float32, dev) if exact else None
        run_pw = True
        if self._shadow_once and not exact and self._shadow == (W.data_ptr(), W._version, w_hat.data_ptr()):
            self._shadow = None         # prefetch_w ran the prologue of THIS forward; the next one runs its own
            self._shadow_once = False
            run_pw = False
        elif exact or self._shadow != (W.data_ptr(), W._version, w_hat.data_ptr()):
            self._shadow = None
            self._shadow_once = False
        else:
            run_pw = False              # sgd_step() wrote w_hat / inv_norm from this very W (same storage, same version)

        # ---- one C call for the whole forward (single GPU, tensor-core path): see csrc/step.cu -------------------------
        plus = self.family in ("elastic_cos", "elastic_arc") and bool(self.cfg.plus)
        if (not exact and self.shard.world == 1 and self.family != "vpl_arcface" and not plus
                and os.environ.get("MH_STEP_API", "1") != "0"):
            return self._forward_step(x, W, labels, state, margins, update_state, want_grad, run_pw, B, B_pad, C_pad, ld)
        if run_pw:
            L.call("mh_prologue_w", _ptr(W), self.layout, Cn, ld, _ptr(w_hat), C_pad, _ptr(w_hat32), _ptr(inv_norm), st)

        x_hat = self._buf("x_hat", (B_pad, L.D), torch.bfloat16, dev)
        x_hat32 = self._buf("x_hat32", (B, L.D), torch.float32, dev)
        xnorm = self._buf("xnorm", (B,), torch.float32, dev)
        t_raw = self._buf("t_raw", (B,), torch.float32, dev)
        label_local = self._buf("label_local", (B_pad,), torch.int32, dev)
        L.call("mh_prologue_x", _ptr(x), _DT[x.dtype], B, B_pad, _ptr(labels), _ptr(W), self.layout, Cn, ld,
               self.shard.c_offset, _ptr(inv_norm), _ptr(x_hat), _ptr(x_hat32), _ptr(xnorm), _ptr(t_raw),
               _ptr(label_local), self.shard.c_total or Cn, st)
        if self.shard.world > 1:
            # every rank needs every row's target cosine (thresholds, EMA); only the owner computed it
            self.shard.comm.allreduce_sum_(t_raw)
        if t_ext is not None:
            if self.family != "vpl_arcface" or self.shard.world > 1:
                raise L.MarginHeadError("external target cosines are supported by the memory-bank heads on one GPU only")
            t_raw.copy_(t_ext.detach().to(torch.float32))

        w_gemm = w_hat
        alpha = None
        if self.family == "vpl_arcface":
            # VPL-ArcFace (criterion.py:716-725): the GEMM runs against the mixed class vectors v_j; the per-row
            # interpolation weight a_{y_i} enters the target logit through row_params' `margins` slot
            if exact:
                raise L.MarginHeadError("VPLArcFace runs on the tensor-core path only")
            alpha = self._buf("vpl_alpha", (Cn,), torch.float32, dev)
            if self.vpl is not None:
                w_gemm = self._buf("vpl_v", (C_pad, L.D), torch.bfloat16, dev)
                L.call("mh_vpl_mix", _ptr(w_hat), _ptr(self.vpl["mem"]), _ptr(self.vpl["life"]), C.c_float(self.vpl["lamda"]),
                       Cn, C_pad, _ptr(w_gemm), _ptr(alpha), st)
            else:
                alpha.zero_()
            # the row's interpolation weight of the target column; an external target cosine is already final
            if t_ext is not None:
                margins = torch.zeros(B, dtype=torch.float32, device=dev)
            elif self.shard.world == 1:
                margins = alpha[labels].contiguous()
            else:
                # class-sharded: only the owner of y_i holds a_{y_i}; label_local is -1 elsewhere -> 0, then all-reduce(SUM)
                ll = label_local[:B].to(torch.int64)
                margins = torch.where(ll >= 0, alpha[ll.clamp_min(0)], torch.zeros((), dtype=torch.float32, device=dev))
                self.shard.comm.allreduce_sum_(margins)
        elif self.family in ("elastic_cos", "elastic_arc"):
            assert margins is not None and margins.numel() == B
            margins = margins.to(device=dev, dtype=torch.float32).contiguous()
            if self.cfg.plus:
                # criterion.py:1007-1012 exactly as written: sorted margins indexed by the permutation
                lo, hi = -1 + 1e-7, 1 - 1e-7
                rank = torch.sort(t_raw.clamp(lo, hi), descending=True)[1]
                margins = torch.sort(margins)[0][rank].contiguous()
        else:
            margins = None

        rowp = self._buf("rowp", (L.RP_PLANES, B_pad), torch.float32, dev)
        L.call("mh_row_params", C.byref(self.cfg), B, _ptr(xnorm), _ptr(t_raw), _ptr(margins), _ptr(state),
               1 if update_state else 0, _ptr(rowp), B_pad, st)

        stats = self._buf("stats", (L.ST_PLANES, B_pad), torch.float32, dev)
        S = pre = logits = stash = guard = None
        if exact:
            S = self._buf("S", (B, Cn), torch.float32, dev)
            # S = x_hat32 [B,512] . w_hat32^T  (w_hat32 is [C,512]: B operand strides (1, 512))
            L.call("mh_sgemm_strided", B, Cn, L.D, _ptr(x_hat32), L.D, 1, _ptr(w_hat32), 1, L.D, _ptr(S), Cn, st)
            if want_dense:
                pre = torch.empty((B, Cn), dtype=torch.float32, device=dev)
                logits = torch.empty((B, Cn), dtype=torch.float32, device=dev)
            L.call("mh_dense_forward", C.byref(self.cfg), _ptr(S), Cn, B, B_pad, Cn, _ptr(rowp), B_pad,
                   _ptr(label_local), _ptr(state), _ptr(stats), _ptr(pre), _ptr(logits), st)
        else:
            n_tiles = int(lib.mh_fwd_num_tiles(C_pad))
            stats_tiles = self._buf("stats_tiles", (n_tiles, L.ST_PLANES, B_pad), torch.float32, dev)
            scratch = self._buf("merge_scratch", (L.MERGE_BLOCKS, L.ST_PLANES, B_pad), torch.float32, dev)
            kind = self._stash_kind(B_pad, C_pad) if want_grad else 0     # 0 none / recompute, 1 stash, 2 guarded stash
            if self.family == "vpl_arcface" and want_grad:
                self.stash_ok()                                           # raises on a forced but ineligible stash
            if kind:
                stash = self._buf("G", (B_pad, C_pad), torch.bfloat16, dev)
            if kind == 2:
                guard = self._buf("stash_guard", (1,), torch.int32, dev, zero=True)
            L.call("mh_tc_forward_ex", C.byref(self.cfg), _ptr(x_hat), B, B_pad, _ptr(w_gemm), Cn, C_pad, _ptr(rowp), B_pad,
                   _ptr(label_local), _ptr(state), _ptr(stats_tiles), _ptr(stash), kind, _ptr(None), 0, st)
            L.call("mh_merge_stats", _ptr(stats_tiles), n_tiles, B, B_pad, _ptr(scratch), _ptr(stats), st)

        if self.shard.world > 1:
            all_stats = self._buf("all_stats", (self.shard.world, L.ST_PLANES, B_pad), torch.float32, dev)
            self.shard.comm.allgather_stats(stats, out=all_stats)
            scratch = self._buf("merge_scratch", (L.MERGE_BLOCKS, L.ST_PLANES, B_pad), torch.float32, dev)
            L.call("mh_merge_stats", _ptr(all_stats), self.shard.world, B, B_pad, _ptr(scratch), _ptr(stats), st)

        rowout = self._buf("rowout", (L.RO_PLANES, B_pad), torch.float32, dev)
        scalars = torch.empty(4, dtype=torch.float32, device=dev)      # loss, acc@1, acc@5, loss_g (fresh: returned to the user)
        sphere = 1 if self.family == "sphereface" else 0
        if guard is None:
            L.call("mh_finalize_rows", _ptr(stats), B_pad, _ptr(rowp), B_pad, B, B, sphere, _ptr(rowout), B_pad, _ptr(scalars),
                   _ptr(state), st)
        else:
            # Guarded stash, sequenced here as csrc/step.cu does for one GPU: the finaliser checks the (globally merged) row
            # sums and sets the flag -- identical on every rank, they all merged the same statistics -- and the general
            # forward follows as launches gated on it.  The collective in between cannot be gated: it runs either way (with
            # the gate closed it gathers the untouched statistics and the gated merge ignores the result).
            c_tot = float(self.shard.c_total or Cn)
            L.call("mh_finalize_rows_ex", _ptr(stats), B_pad, _ptr(rowp), B_pad, B, B, sphere, _ptr(rowout), B_pad, _ptr(scalars),
                   _ptr(state), C.c_float(math.ldexp(c_tot, -102)), _ptr(guard), _ptr(None), 0, st)
            L.call("mh_tc_forward_ex", C.byref(self.cfg), _ptr(x_hat), B, B_pad, _ptr(w_gemm), Cn, C_pad, _ptr(rowp), B_pad,
                   _ptr(label_local), _ptr(state), _ptr(stats_tiles), _ptr(None), 0, _ptr(guard), 1, st)
            L.call("mh_merge_stats_ex", _ptr(stats_tiles), n_tiles, B, B_pad, _ptr(scratch), _ptr(stats), _ptr(guard), 1, st)
            if self.shard.world > 1:
                self.shard.comm.allgather_stats(stats, out=all_stats)
                L.call("mh_merge_stats_ex", _ptr(all_stats), self.shard.world, B, B_pad, _ptr(scratch), _ptr(stats), _ptr(guard), 1, st)
            L.call("mh_finalize_rows_ex", _ptr(stats), B_pad, _ptr(rowp), B_pad, B, B, sphere, _ptr(rowout), B_pad, _ptr(scalars),
                   _ptr(state), C.c_float(0.0), _ptr(None), _ptr(guard), 1, st)
        return dict(guard=guard, B=B, B_pad=B_pad, C_pad=C_pad, x_dtype=x.dtype, w_hat=w_hat, w_hat32=w_hat32, inv_norm=inv_norm,
                    x_hat=x_hat, x_hat32=x_hat32, xnorm=xnorm, label_local=label_local, rowp=rowp, rowout=rowout,
                    scalars=scalars, S=S, pre=pre, logits=logits, exact=exact, gen=self._gen, state=state,
                    W_shape=tuple(W.shape), ld=ld, stash=stash, w_gemm=w_gemm, vpl_alpha=alpha, t_ext=t_ext is not None)

    def _forward_step(self, x, W, labels, state, margins, update_state, want_grad, run_pw, B, B_pad, C_pad, ld):
        """The forward through mh_step_forward: same kernels, same workspaces, one FFI call (host overhead of the
        launch-bound configs).  The descriptor is rebuilt only when a shape, dtype or workspace pointer changes."""
        dev = x.device
        Cn = self.C
        stash = self._stash_kind(B_pad, C_pad) if want_grad else 0        # 0 none / recompute, 1 stash, 2 guarded stash
        guarded = stash == 2
        key = (B, x.dtype, dev, ld, bool(want_grad), stash, _selfproj(), _merged_bwd(), _merged_fwd(),
               os.environ.get("MH_STEP_GRAPH", "1"))
        if self._step_key != key:
            lib = L.load()
            n_tiles = int(lib.mh_fwd_num_tiles(C_pad))
            b = self._buf
            T = dict(
                w_hat=b("w_hat", (C_pad, L.D), torch.bfloat16, dev), inv_norm=b("inv_norm", (Cn,), torch.float32, dev),
                x_hat=b("x_hat", (B_pad, L.D), torch.bfloat16, dev), x_hat32=b("x_hat32", (B, L.D), torch.float32, dev),
                xnorm=b("xnorm", (B,), torch.float32, dev), t_raw=b("t_raw", (B,), torch.float32, dev),
                label_local=b("label_local", (B_pad,), torch.int32, dev),
                rowp=b("rowp", (L.RP_PLANES, B_pad), torch.float32, dev),
                stats_tiles=b("stats_tiles", (n_tiles, L.ST_PLANES, B_pad), torch.float32, dev),
                merge_scratch=b("merge_scratch", (L.MERGE_BLOCKS, L.ST_PLANES, B_pad), torch.float32, dev),
                stats=b("stats", (L.ST_PLANES, B_pad), torch.float32, dev),
                rowout=b("rowout", (L.RO_PLANES, B_pad), torch.float32, dev))
            if _merged_fwd() and self.layout == L.LAYOUT_CD and not guarded:
                T.update(pw_ready=b("pw_ready", (C_pad // L.NTILE + 1,), torch.int32, dev))
            part_splits = 0
            if want_grad:
                ns = C.c_int(0)
                L.call("mh_tc_backward_dx", _ptr(None), B_pad, C_pad, _ptr(None), _ptr(None), C.byref(ns), _ptr(None), _stream())
                part_splits = ns.value
                T.update(bc=b("G", (B_pad, C_pad), torch.bfloat16, dev),
                         dxhat_part=b("dxhat_part", (part_splits, B_pad, L.D), torch.float32, dev),
                         gscal=b("gscal", (2,), torch.float32, dev), dx_sync=b("dx_sync", (L.DX_SYNC_INTS,), torch.int32, dev))
                if _selfproj() or _merged_bwd() or guarded:
                    T.update(rpart=b("dw_rpart", (4, C_pad), torch.float32, dev),
                             rflag=b("dw_rflag", (C_pad // L.TILE,), torch.int32, dev))
                if guarded:      # the guarded stash always projects inside the dW kernel (its stash does not give cos back)
                    T.update(guard=b("stash_guard", (1,), torch.int32, dev, zero=True))
                elif not _selfproj():
                    T.update(r_colsum=b("r_colsum", (B_pad // L.TILE if stash else 1, C_pad), torch.float32, dev))
                if _merged_bwd():
                    T.update(prog=b("bwd_prog", (2,), torch.int32, dev))
                    ms = C.c_int(0)
                    L.call("mh_tc_backward_dxdw", _ptr(None), B_pad, Cn, C_pad, _ptr(None), _ptr(None), _ptr(None), _ptr(None),
                           self.layout, _ptr(None), ld, _ptr(None), C.byref(ms), _ptr(None), _ptr(None), _ptr(None), _stream())
                    if ms.value > part_splits:
                        part_splits = ms.value
                        T["dxhat_part"] = b("dxhat_part", (part_splits, B_pad, L.D), torch.float32, dev)
                if stash:
                    T.update(xs=b("xs", (B_pad, L.D), torch.bfloat16, dev), rho=b("rho", (B_pad,), torch.float32, dev),
                             gty=b("gty", (B_pad,), torch.float32, dev),
                             dxhat_full=b("dxhat_full", (1, B_pad, L.D), torch.float32, dev))
            ws = L.MhStepWs()
            ws.graph_cache = self._graph_cache_handle()
            ws.B, ws.B_pad, ws.C, ws.C_pad, ws.ld = B, B_pad, Cn, C_pad, ld
            ws.layout, ws.x_dtype = self.layout, _DT[x.dtype]
            ws.n_tiles, ws.part_splits = n_tiles, part_splits
            for k, t in T.items():
                setattr(ws, k, t.data_ptr())
            # the descriptor holds raw pointers: T keeps the tensors alive for as long as the descriptor is cached
            self._step_key, self._step_ws, self._step_T = key, ws, T
        T = self._step_T
        ws = self._step_ws
        if self.family in ("elastic_cos", "elastic_arc"):
            assert margins is not None and margins.numel() == B
            margins = margins.to(device=dev, dtype=torch.float32).contiguous()
        else:
            margins = None
        scalars = torch.empty(4, dtype=torch.float32, device=dev)      # loss, acc@1, acc@5, loss_g (fresh: returned to the user)
        L.call("mh_step_forward", C.byref(self.cfg), C.byref(ws), _ptr(x), _ptr(labels), _ptr(W), _ptr(margins), _ptr(state),
               1 if update_state else 0, 1 if run_pw else 0, stash, _ptr(scalars), _stream())
        return dict(B=B, B_pad=B_pad, C_pad=C_pad, x_dtype=x.dtype, w_hat=T["w_hat"], w_hat32=None, inv_norm=T["inv_norm"],
                    x_hat=T["x_hat"], x_hat32=T["x_hat32"], xnorm=T["xnorm"], label_local=T["label_local"], rowp=T["rowp"],
                    rowout=T["rowout"], scalars=scalars, S=None, pre=None, logits=None, exact=False, gen=self._gen,
                    state=state, W_shape=tuple(W.shape), ld=ld, stash=T.get("bc") if stash else None, w_gemm=T["w_hat"],
                    vpl_alpha=None, step_ws=ws if want_grad else None, step_stash=stash, margins=margins)

    # -- backward of the fused loss ------------------------------------------------------------------
    @_on_device_of(0)
    def backward(self, ctx: Dict, g_loss: torch.Tensor, g_lossg: Optional[torch.Tensor],
                 need_dx: bool = True, need_dw: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Returns (dxhat_partials_or_dx, dW).  For world == 1 the first element is dx [B,512] in x's dtype."""
        if ctx["gen"] != self._gen:
            raise L.MarginHeadError("margin head workspaces were overwritten by a later forward; "
                                    "run backward before the next forward of the same head")
        dev = ctx["x_hat"].device
        B, B_pad, C_pad, Cn = ctx["B"], ctx["B_pad"], ctx["C_pad"], self.C
        st = _stream()
        if ctx.get("step_ws") is not None:
            # one C call for the whole backward (csrc/step.cu); the forward of this step built the descriptor
            if g_loss is not None and g_loss.dtype != torch.float32:
                g_loss = g_loss.float()
            if g_lossg is not None and g_lossg.dtype != torch.float32:
                g_lossg = g_lossg.float()
            dx = torch.empty((B, L.D), dtype=ctx["x_dtype"], device=dev) if need_dx else None
            dW = torch.empty(ctx["W_shape"], dtype=torch.float32, device=dev) if need_dw else None
            L.call("mh_step_backward", C.byref(self.cfg), C.byref(ctx["step_ws"]), int(ctx["step_stash"]),
                   _ptr(ctx["state"]), _ptr(g_loss), _ptr(g_lossg), _ptr(dx), _ptr(dW), st)
            return dx, dW
        gscal = self._gscal(g_loss, g_lossg, B, dev)
        rowp, rowout, state = ctx["rowp"], ctx["rowout"], ctx["state"]
        lse2 = rowout[L.RO["LSE2"]]
        if ctx["exact"]:
            S = ctx["S"]
            L.call("mh_dense_backward_dc", C.byref(self.cfg), _ptr(S), Cn, B, Cn, _ptr(rowp), B_pad,
                   _ptr(ctx["label_local"]), _ptr(state), _ptr(lse2), _ptr(None), _ptr(None), _ptr(None), st)
            return self._exact_grads(ctx, S, gscal, rowout[L.RO["AUX0"]], rowout[L.RO["AUX1"]], need_dx, need_dw)
        w_hat, x_hat, label_local = ctx["w_hat"], ctx["x_hat"], ctx["label_local"]
        stash = ctx.get("stash")
        if self.family == "vpl_arcface":
            return self._backward_vpl(ctx, gscal, need_dx, need_dw)
        # dW projection r_j = w^_j . dw^_j: produced beforehand by the dx side pass / the backward-G column sums (default),
        # or (MH_DW_SELFPROJ=1) taken from the dW accumulators themselves - measured break-even, see DESIGN.md 4.2
        guard = ctx.get("guard")                 # guarded stash: the dW GEMM always projects from its own accumulators
        selfp = _selfproj() or guard is not None
        r_parts = B_pad // L.TILE if stash is not None else 1      # stash: one partial per 128-row block (no atomics)
        rsum = None if selfp else self._buf("r_colsum", (r_parts, C_pad), torch.float32, dev)
        ns = C.c_int(0)
        L.call("mh_tc_backward_dx", _ptr(None), B_pad, C_pad, _ptr(None), _ptr(None), C.byref(ns), _ptr(None), st)
        n_split = ns.value
        split_stride = B_pad * L.D
        dxsync = self._buf("dx_sync", (L.DX_SYNC_INTS,), torch.int32, dev)
        dx = dW = None
        if _merged_bwd() and need_dx and need_dw:
            ms = C.c_int(0)
            L.call("mh_tc_backward_dxdw", _ptr(None), B_pad, Cn, C_pad, _ptr(None), _ptr(None), _ptr(None), _ptr(None),
                   self.layout, _ptr(None), ctx["ld"], _ptr(None), C.byref(ms), _ptr(None), _ptr(None), _ptr(None), st)
            if ms.value > 0:
                return self._backward_merged(ctx, gscal, ms.value)
        if stash is not None:
            # stash mode: G_ij = rho_i E'_ij off the target column; the target column is a sparse fp32 term.
            G = stash
            xs = self._buf("xs", (B_pad, L.D), torch.bfloat16, dev)
            rho = self._buf("rho", (B_pad,), torch.float32, dev)
            gty = self._buf("gty", (B_pad,), torch.float32, dev)
            self._stash_prep(ctx, xs, rho, gty, st)
            if need_dx or not selfp:       # without self-projection the dx kernel's side pass also produces r_colsum
                part = self._buf("dxhat_part", (n_split, B_pad, L.D), torch.float32, dev)
                if selfp:
                    L.call("mh_tc_backward_dx", _ptr(G), B_pad, C_pad, _ptr(w_hat), _ptr(part), C.byref(ns), _ptr(dxsync), st)
                else:
                    L.call("mh_tc_backward_dx_stash", C.byref(self.cfg), _ptr(G), B_pad, Cn, C_pad, _ptr(w_hat), _ptr(rho),
                           _ptr(rowp), B_pad, _ptr(part), _ptr(rsum), C.byref(ns), _ptr(dxsync), st)
            if need_dx:
                full = self._buf("dxhat_full", (1, B_pad, L.D), torch.float32, dev)
                L.call("mh_stash_dx_combine", _ptr(part), n_split, split_stride, _ptr(rho), _ptr(gty), _ptr(label_local),
                       _ptr(w_hat), B, _ptr(full), st)
                dx = self._finish_dx(ctx, full, 1, split_stride, gscal, rowout[L.RO["AUX0"]], rowout[L.RO["AUX1"]])
        else:
            G = self._buf("G", (B_pad, C_pad), torch.bfloat16, dev)
            L.call("mh_tc_backward_g", C.byref(self.cfg), _ptr(x_hat), B, B_pad, _ptr(w_hat), Cn, C_pad,
                   _ptr(rowp), B_pad, _ptr(label_local), _ptr(state), _ptr(lse2), _ptr(G),
                   _ptr(rsum if (need_dw and not selfp) else None), st)
            xs = x_hat
            if need_dx:
                part = self._buf("dxhat_part", (n_split, B_pad, L.D), torch.float32, dev)
                L.call("mh_tc_backward_dx", _ptr(G), B_pad, C_pad, _ptr(w_hat), _ptr(part), C.byref(ns), _ptr(dxsync), st)
                dx = self._finish_dx(ctx, part, n_split, split_stride, gscal, rowout[L.RO["AUX0"]], rowout[L.RO["AUX1"]])
        if need_dw:
            dW = torch.empty(ctx["W_shape"], dtype=torch.float32, device=dev)
            if selfp:
                rpart = self._buf("dw_rpart", (4, C_pad), torch.float32, dev)
                rflag = self._buf("dw_rflag", (C_pad // L.TILE,), torch.int32, dev)
                L.call("mh_tc_backward_dw_proj", _ptr(G), B_pad, Cn, C_pad, _ptr(xs), _ptr(w_hat), _ptr(ctx["inv_norm"]),
                       _ptr(gscal), self.layout, _ptr(dW), ctx["ld"], _ptr(rpart), _ptr(rflag), st)
            else:
                L.call("mh_tc_backward_dw_fused", _ptr(G), B_pad, Cn, C_pad, _ptr(xs), _ptr(w_hat), _ptr(ctx["inv_norm"]),
                       _ptr(rsum), r_parts, _ptr(gscal), self.layout, _ptr(dW), ctx["ld"], st)
            if stash is not None:
                L.call("mh_stash_dw_target", _ptr(gty), _ptr(label_local), _ptr(ctx["x_hat32"]), _ptr(w_hat),
                       _ptr(ctx["inv_norm"]), _ptr(gscal), B, self.layout, _ptr(dW), ctx["ld"], st)
        return dx, dW

    def _backward_merged(self, ctx, gscal, n_split):
        """Both backward GEMMs in one persistent kernel (mh_tc_backward_dxdw, MH_BWD_MERGED=1): the dx role and the
        self-projecting dW role share the stash / G and w^ through L2.  Stash or recompute mode, single GPU or sharded."""
        dev = ctx["x_hat"].device
        B, B_pad, C_pad, Cn = ctx["B"], ctx["B_pad"], ctx["C_pad"], self.C
        st = _stream()
        rowp, rowout, state = ctx["rowp"], ctx["rowout"], ctx["state"]
        w_hat, x_hat, label_local = ctx["w_hat"], ctx["x_hat"], ctx["label_local"]
        stash = ctx.get("stash")
        split_stride = B_pad * L.D
        if stash is not None:
            G = stash
            xs = self._buf("xs", (B_pad, L.D), torch.bfloat16, dev)
            rho = self._buf("rho", (B_pad,), torch.float32, dev)
            gty = self._buf("gty", (B_pad,), torch.float32, dev)
            self._stash_prep(ctx, xs, rho, gty, st)
        else:
            G = self._buf("G", (B_pad, C_pad), torch.bfloat16, dev)
            L.call("mh_tc_backward_g", C.byref(self.cfg), _ptr(x_hat), B, B_pad, _ptr(w_hat), Cn, C_pad, _ptr(rowp), B_pad,
                   _ptr(label_local), _ptr(state), _ptr(rowout[L.RO["LSE2"]]), _ptr(G), _ptr(None), st)
            xs = x_hat
        part = self._buf("dxhat_part", (n_split, B_pad, L.D), torch.float32, dev)
        rpart = self._buf("dw_rpart", (4, C_pad), torch.float32, dev)
        rflag = self._buf("dw_rflag", (C_pad // L.TILE,), torch.int32, dev)
        prog = self._buf("bwd_prog", (2,), torch.int32, dev)
        dW = torch.empty(ctx["W_shape"], dtype=torch.float32, device=dev)
        ns = C.c_int(0)
        L.call("mh_tc_backward_dxdw", _ptr(G), B_pad, Cn, C_pad, _ptr(w_hat), _ptr(xs), _ptr(ctx["inv_norm"]), _ptr(gscal),
               self.layout, _ptr(dW), ctx["ld"], _ptr(part), C.byref(ns), _ptr(rpart), _ptr(rflag), _ptr(prog), st)
        aux0, aux1 = rowout[L.RO["AUX0"]], rowout[L.RO["AUX1"]]
        if stash is not None:
            full = self._buf("dxhat_full", (1, B_pad, L.D), torch.float32, dev)
            L.call("mh_stash_dx_combine", _ptr(part), n_split, split_stride, _ptr(rho), _ptr(gty), _ptr(label_local),
                   _ptr(w_hat), B, _ptr(full), st)
            dx = self._finish_dx(ctx, full, 1, split_stride, gscal, aux0, aux1)
            L.call("mh_stash_dw_target", _ptr(gty), _ptr(label_local), _ptr(ctx["x_hat32"]), _ptr(w_hat),
                   _ptr(ctx["inv_norm"]), _ptr(gscal), B, self.layout, _ptr(dW), ctx["ld"], st)
        else:
            dx = self._finish_dx(ctx, part, n_split, split_stride, gscal, aux0, aux1)
        return dx, dW

    def _stash_prep(self, ctx, xs, rho, gty, st):
        """rho_i, gty_i and the scaled rows of the stash backward.  Guarded stash: a gated backward-G launch first rewrites the
        stash with the recomputed G when the forward raised the flag, and the prep kernel then uses rho = 1 / gty = 0."""
        B, B_pad, C_pad, Cn = ctx["B"], ctx["B_pad"], ctx["C_pad"], self.C
        rowp, rowout, guard = ctx["rowp"], ctx["rowout"], ctx.get("guard")
        if guard is not None:
            L.call("mh_tc_backward_g_ex", C.byref(self.cfg), _ptr(ctx["x_hat"]), B, B_pad, _ptr(ctx["w_hat"]), Cn, C_pad,
                   _ptr(rowp), B_pad, _ptr(ctx["label_local"]), _ptr(ctx["state"]), _ptr(rowout[L.RO["LSE2"]]),
                   _ptr(ctx["stash"]), _ptr(None), _ptr(guard), 1, st)
        L.call("mh_stash_prep_ex", C.byref(self.cfg), _ptr(rowp), B_pad, _ptr(rowout), B_pad, _ptr(ctx["x_hat32"]), B, B_pad,
               _ptr(xs), _ptr(rho), _ptr(gty), _ptr(guard), st)

    def _backward_vpl(self, ctx, gscal, need_dx, need_dw):
        """VPL-ArcFace backward on the stash: dx^ = diag(rho) E'.v + G_iy (1 - a_y) w^_y;  dw^_j = (1 - a_j) E'^T.(rho x^)
        + the target term; the memory bank carries no gradient (oracle/vpl_oracle.py has the derivation)."""
        stash = ctx.get("stash")
        if stash is None:
            raise L.MarginHeadError("VPLArcFace needs the stash backward (fixed logit scale with s*log2(e)*2 <= 200)")
        dev = stash.device
        B, B_pad, C_pad, Cn = ctx["B"], ctx["B_pad"], ctx["C_pad"], self.C
        st = _stream()
        rowp, rowout, w_hat, label_local = ctx["rowp"], ctx["rowout"], ctx["w_hat"], ctx["label_local"]
        xs = self._buf("xs", (B_pad, L.D), torch.bfloat16, dev)
        rho = self._buf("rho", (B_pad,), torch.float32, dev)
        gty = self._buf("gty", (B_pad,), torch.float32, dev)
        L.call("mh_stash_prep", C.byref(self.cfg), _ptr(rowp), B_pad, _ptr(rowout), B_pad, _ptr(ctx["x_hat32"]), B, B_pad,
               _ptr(xs), _ptr(rho), _ptr(gty), st)
        ext = bool(ctx.get("t_ext"))
        # external target cosine (QAFace): the target column does not touch dx / dW here; d loss / d t_i = gscal * gty_i goes
        # back to the caller's autograd graph instead
        gty_used = self._buf("gty_zero", (B_pad,), torch.float32, dev, zero=True) if ext else gty
        ctx["dt_ext"] = (gty[:B] * gscal[0]) if ext else None
        dx = dW = None
        if need_dx:
            ns = C.c_int(0)
            L.call("mh_tc_backward_dx", _ptr(None), B_pad, C_pad, _ptr(None), _ptr(None), C.byref(ns), _ptr(None), st)
            part = self._buf("dxhat_part", (ns.value, B_pad, L.D), torch.float32, dev)
            L.call("mh_tc_backward_dx", _ptr(stash), B_pad, C_pad, _ptr(ctx["w_gemm"]), _ptr(part), C.byref(ns), _ptr(None), st)
            full = self._buf("dxhat_full", (1, B_pad, L.D), torch.float32, dev)
            # gty already carries (1 - a_y): the target column reaches x^ through w^_y only
            L.call("mh_stash_dx_combine", _ptr(part), ns.value, B_pad * L.D, _ptr(rho), _ptr(gty_used), _ptr(label_local),
                   _ptr(w_hat), B, _ptr(full), st)
            dx = self._finish_dx(ctx, full, 1, B_pad * L.D, gscal, rowout[L.RO["AUX0"]], rowout[L.RO["AUX1"]])
        if need_dw:
            dwh = self._buf("dw_hat_raw", (C_pad, L.D), torch.float32, dev)
            L.call("mh_tc_backward_dw", _ptr(stash), B_pad, C_pad, _ptr(xs), _ptr(dwh), st)
            dW = torch.empty(ctx["W_shape"], dtype=torch.float32, device=dev)
            beta = self._buf("vpl_beta", (Cn,), torch.float32, dev)
            torch.sub(1.0, ctx["vpl_alpha"], out=beta)               # 1 - a_j; the normalise-backward is linear in dw^
            L.call("mh_norm_backward_w", _ptr(dwh), _ptr(w_hat), _ptr(None), _ptr(ctx["inv_norm"]), _ptr(gscal), _ptr(beta),
                   Cn, self.layout, _ptr(dW), ctx["ld"], st)
            if not ext:
                L.call("mh_stash_dw_target", _ptr(gty), _ptr(label_local), _ptr(ctx["x_hat32"]), _ptr(w_hat),
                       _ptr(ctx["inv_norm"]), _ptr(gscal), B, self.layout, _ptr(dW), ctx["ld"], st)
        return dx, dW

    def _gscal(self, g_loss, g_lossg, B_total, dev) -> torch.Tensor:
        """Device scalars {g_loss / B_total, g_lossg} in one launch (no host sync under a GradScaler)."""
        gscal = torch.empty(2, dtype=torch.float32, device=dev)
        if g_loss is not None and g_loss.dtype != torch.float32:
            g_loss = g_loss.float()
        if g_lossg is not None and g_lossg.dtype != torch.float32:
            g_lossg = g_lossg.float()
        L.call("mh_make_gscal", _ptr(g_loss), _ptr(g_lossg), B_total, _ptr(gscal), _stream())
        return gscal

    def _finish_dx(self, ctx, part, n_split, split_stride, gscal, aux0, aux1):
        """Sum split partials (+ cross-rank reduce-scatter when sharded) and apply normalise-backward."""
        B = ctx["B"]
        dev = part.device
        st = _stream()
        if self.shard.world > 1:
            R = self.shard.world
            assert B % R == 0
            Bl = B // R
            if n_split > 1:               # recompute mode: sum the split-K partials with the combine kernel (rho = NULL)
                full = self._buf("dxhat_full", (1, ctx["B_pad"], L.D), torch.float32, dev)
                L.call("mh_stash_dx_combine", _ptr(part), n_split, split_stride, _ptr(None), _ptr(None), _ptr(None),
                       _ptr(None), B, _ptr(full), st)
                part = full
            mine = self.shard.comm.reduce_scatter_rows(part[0, :B])    # a contiguous view: no copy
            r0 = self.shard.rank * Bl
            dx = torch.empty((Bl, L.D), dtype=ctx["x_dtype"], device=dev)
            rowp_off = ctx["rowp"][:, r0:]
            # rowp planes stay at pitch B_pad; offsetting the base pointer by r0 rows keeps the pitch valid
            L.call("mh_norm_backward_x", _ptr(mine), 1, 0, _ptr(ctx["x_hat32"][r0:]), _ptr(ctx["xnorm"][r0:]),
                   C.c_void_p(rowp_off.data_ptr()), ctx["B_pad"], C.c_void_p(aux0[r0:].data_ptr()),
                   C.c_void_p(aux1[r0:].data_ptr()), _ptr(gscal), Bl, _ptr(dx), _DT[ctx["x_dtype"]], st)
            return dx
        dx = torch.empty((B, L.D), dtype=ctx["x_dtype"], device=dev)
        L.call("mh_norm_backward_x", _ptr(part), n_split, split_stride, _ptr(ctx["x_hat32"]), _ptr(ctx["xnorm"]),
               _ptr(ctx["rowp"]), ctx["B_pad"], _ptr(aux0), _ptr(aux1), _ptr(gscal), B, _ptr(dx), _DT[ctx["x_dtype"]], st)
        return dx

    def _exact_grads(self, ctx, dc, gscal, aux0, aux1, need_dx, need_dw):
        """dx^ = dc . w^ ; dw^ = dc^T . x^ with the fp32 SIMT GEMM, then the normalise-backward kernels."""
        dev = dc.device
        B, Cn = ctx["B"], self.C
        st = _stream()
        dx = dW = None
        if need_dx:
            dxh = self._buf("dxhat_exact", (1, B, L.D), torch.float32, dev)
            L.call("mh_sgemm_strided", B, L.D, Cn, _ptr(dc), Cn, 1, _ptr(ctx["w_hat32"]), L.D, 1, _ptr(dxh), L.D, st)
            dx = self._finish_dx(ctx, dxh, 1, B * L.D, gscal, aux0, aux1)
        if need_dw:
            dwh = self._buf("dw_hat_exact", (Cn, L.D), torch.float32, dev)
            # dw^ [C,512] = dc^T [C,B] . x^ [B,512] : A strides (1, C)
            L.call("mh_sgemm_strided", Cn, L.D, B, _ptr(dc), 1, Cn, _ptr(ctx["x_hat32"]), L.D, 1, _ptr(dwh), L.D, st)
            dW = torch.empty(ctx["W_shape"], dtype=torch.float32, device=dev)
            L.call("mh_norm_backward_w", _ptr(dwh), _ptr(None), _ptr(ctx["w_hat32"]), _ptr(ctx["inv_norm"]), _ptr(gscal),
                   _ptr(None), Cn, self.layout, _ptr(dW), ctx["ld"], st)
        return dx, dW

    # -- backward of the compat (materialised logits) outputs ---------------------------------------
    @_on_device_of(0)
    def backward_dense(self, ctx: Dict, dlogits: Optional[torch.Tensor], dpre: Optional[torch.Tensor],
                       g_lossg: Optional[torch.Tensor], need_dx=True, need_dw=True):
        if ctx["gen"] != self._gen:
            raise L.MarginHeadError("margin head workspaces were overwritten by a later forward")
        dev = ctx["x_hat"].device
        B, B_pad, Cn = ctx["B"], ctx["B_pad"], self.C
        st = _stream()
        if dlogits is None:
            dlogits = torch.zeros((B, Cn), dtype=torch.float32, device=dev)
        dlogits = dlogits.to(torch.float32).contiguous()
        dpre = dpre.to(torch.float32).contiguous() if dpre is not None else None
        rowaux = torch.zeros((2, B), dtype=torch.float32, device=dev)
        S = ctx["S"]
        L.call("mh_dense_backward_dc", C.byref(self.cfg), _ptr(S), Cn, B, Cn, _ptr(ctx["rowp"]), B_pad,
               _ptr(ctx["label_local"]), _ptr(ctx["state"]), _ptr(None), _ptr(dlogits), _ptr(dpre), _ptr(rowaux), st)
        gscal = self._gscal(torch.ones((), dtype=torch.float32, device=dev), g_lossg, 1, dev)
        return self._exact_grads(ctx, S, gscal, rowaux[0], rowaux[1], need_dx, need_dw)


class FusedMarginLossFn(torch.autograd.Function):
    """loss_id, loss_g, acc1, acc5, norms = f(x, W, labels).  Only loss_id / loss_g are differentiable."""

    @staticmethod
    def forward(ctx, x, W, labels, engine: HeadEngine, state, margins, update_state, grad_enabled=True, t_ext=None):
        # grad_enabled = torch.is_grad_enabled() at the call site (always False inside Function.forward)
        needs = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (t_ext is not None and ctx.needs_input_grad[8])
        c = engine.forward(x, W, labels, state, margins, update_state=update_state,
                           want_grad=bool(grad_enabled and needs), t_ext=t_ext)
        ctx.engine = engine
        ctx.c = c
        ctx.x_dtype = x.dtype
        # undefined upstream gradients arrive as None instead of freshly filled zero tensors (four fill launches per step
        # for loss_g / acc1 / acc5 / norms when only `loss` is differentiated); the C ABI takes NULL for "zero"
        ctx.set_materialize_grads(False)
        loss, acc1, acc5, loss_g = c["scalars"].unbind(0)           # views of a tensor created by this forward
        norms = c["rowp"][L.RP["NORMS"], :c["B"]].clone().unsqueeze(1)
        ctx.mark_non_differentiable(norms, acc1, acc5)
        return loss, loss_g, acc1, acc5, norms

    @staticmethod
    def backward(ctx, g_loss, g_lossg, _ga1, _ga5, _gn):
        eng = ctx.engine
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx, dW = eng.backward(ctx.c, g_loss, g_lossg, need_dx, need_dw)       # None = zero upstream gradient (NULL in the C ABI)
        return dx, dW, None, None, None, None, None, None, ctx.c.get("dt_ext")


class DenseMarginLogitsFn(torch.autograd.Function):
    """Compat mode: materialise the reference's ``[pre_margin_logits, logits]`` (criterion.py:301 etc.)."""

    @staticmethod
    def forward(ctx, x, W, labels, engine: HeadEngine, state, margins, update_state):
        c = engine.forward(x, W, labels, state, margins, update_state=update_state, want_dense=True)
        ctx.engine = engine
        ctx.c = c
        ctx.set_materialize_grads(False)
        norms = c["rowp"][L.RP["NORMS"], :c["B"]].clone().unsqueeze(1)
        loss_g = state[3].clone()
        ctx.mark_non_differentiable(norms)
        return c["pre"], c["logits"], norms, loss_g

    @staticmethod
    def backward(ctx, dpre, dlogits, _gn, g_lossg):
        eng = ctx.engine
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # an all-zero dpre (the caller only differentiated `logits`) is handled by the kernel: no host sync here
        dx, dW = eng.backward_dense(ctx.c, dlogits, dpre, g_lossg, need_dx, need_dw)
        return dx, dW, None, None, None, None, None
