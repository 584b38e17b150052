"""Pair verification on embeddings: the reference's LFW 10-fold protocol with the per-pair cosine on the GPU.

Mirrors main_code/utils/model_utils.py:320-474 (``compute_auc``, ``evaluate``, ``tune_threshold_roc``,
``cross_validate_kfold``) with the same names, argument meaning and return values, except that the functions take the
per-pair cosines (or the embedding pairs) instead of a model and an image dataset: in the reference every one of them
spends its time in ``F.normalize(model(img1)) . F.normalize(model(img2))``; that step is ``pair_cosine`` here
(``mh_pair_cosine``, one CUDA kernel, no PyTorch fallback), the backbone forward stays the caller's.  The statistics on
the ``[N_pairs]`` cosine vector are host-side scikit-learn calls exactly as in the reference, with its missing
``roc_auc_score`` import (model_utils.py:352 vs :14) supplied.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _lib as L
from .functional import _DT, _ptr, _stream


def pair_cosine(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    """cos_i = <e1_i, e2_i> / (|e1_i| |e2_i|), fp32 [N]; e1, e2 are CUDA [N, d] in fp32 / bf16 / fp16."""
    if not (e1.is_cuda and e2.is_cuda):
        raise L.MarginHeadError("pair_cosine inputs must be CUDA tensors (no CPU fallback)")
    if e1.dim() != 2 or e1.shape != e2.shape or e1.dtype != e2.dtype or e1.dtype not in _DT:
        raise ValueError("pair_cosine expects two [N, d] tensors of the same shape and dtype (fp32 / bf16 / fp16)")
    if e1.device != e2.device:
        raise L.MarginHeadError(f"pair_cosine inputs must share one device: {e1.device} vs {e2.device}")
    with torch.cuda.device(e1.device):               # launch on the GPU that holds the data, whatever is current
        return _pair_cosine(e1, e2)


def _pair_cosine(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    L.check(L.load().mh_device_check(), "mh_device_check")
    if e1.stride(1) != 1:
        e1 = e1.contiguous()
    if e2.stride(1) != 1:
        e2 = e2.contiguous()
    out = torch.empty(e1.shape[0], dtype=torch.float32, device=e1.device)
    if e1.shape[0] == 0:
        return out
    L.call("mh_pair_cosine", _ptr(e1), _ptr(e2), _DT[e1.dtype], e1.shape[0], e1.shape[1], e1.stride(0), e2.stride(0),
           _ptr(out), _stream())
    return out


def _np(cos, labels) -> Tuple[np.ndarray, np.ndarray]:
    if isinstance(cos, torch.Tensor):
        cos = cos.detach().float().cpu().numpy()
    if isinstance(labels, torch.Tensor):
        labels = labels.detach().cpu().numpy()
    return np.asarray(cos), np.asarray(labels).astype(np.int64)


def tune_threshold_roc(cos, labels) -> Tuple[float, float]:
    """Youden-index threshold (max TPR - FPR) and the accuracy at it (model_utils.py:379-414)."""
    from sklearn.metrics import roc_curve
    cos, labels = _np(cos, labels)
    fpr, tpr, thresholds = roc_curve(labels, cos)
    best_thresh = thresholds[int(np.argmax(tpr - fpr))]
    predictions = (cos > best_thresh).astype(int)
    best_acc = 100.0 * (predictions == labels).sum() / len(labels)
    return float(best_thresh), float(best_acc)


def evaluate(cos, labels, threshold: float = 0.33) -> float:
    """Accuracy in percent of ``cos > threshold`` (model_utils.py:354-377)."""
    cos, labels = _np(cos, labels)
    total = len(labels)
    return 100.0 * float(((cos > threshold).astype(np.int64) == labels).sum()) / total if total > 0 else 0.0


def compute_auc(cos, labels) -> float:
    """ROC AUC; 0.0 when only one class is present (model_utils.py:320-352)."""
    from sklearn.metrics import roc_auc_score
    cos, labels = _np(cos, labels)
    if len(np.unique(labels)) < 2:
        return 0.0
    return float(roc_auc_score(labels, cos))


def cross_validate_kfold(e1: torch.Tensor, e2: torch.Tensor, labels, k_fold: int = 10, verbose: bool = False):
    """10-fold evaluation of embedding pairs (model_utils.py:416-474): StratifiedKFold(k, shuffle, seed 42); the threshold
    is tuned on the held-out fold and accuracy / AUC are scored on the other k-1 folds.
    Returns (mean_acc, std_acc, mean_auc, std_auc) like the reference."""
    from sklearn.model_selection import StratifiedKFold
    cos, labels = _np(pair_cosine(e1, e2), labels)        # one kernel for all pairs; folds only index the result
    skf = StratifiedKFold(n_splits=k_fold, shuffle=True, random_state=42)
    fold_accuracies, fold_aucs = [], []
    for fold, (train_idx, val_idx) in enumerate(skf.split(np.zeros((len(labels), 1)), labels), 1):
        best_thresh, _ = tune_threshold_roc(cos[val_idx], labels[val_idx])
        acc = evaluate(cos[train_idx], labels[train_idx], best_thresh)
        auc = compute_auc(cos[train_idx], labels[train_idx])
        fold_accuracies.append(acc)
        fold_aucs.append(auc)
        if verbose:
            print(f"=== Fold {fold}/{k_fold} === threshold {best_thresh:.4f}  accuracy {acc:.3f}%  AUC {auc:.4f}")
    return (float(np.mean(fold_accuracies)), float(np.std(fold_accuracies)), float(np.mean(fold_aucs)),
            float(np.std(fold_aucs)))
