"""CPU oracle for the pair-verification row: TEST INFRASTRUCTURE ONLY.

Restates, on precomputed embeddings, what the reference's LFW evaluator computes per pair and per fold:
  * cosine of L2-normalised embedding pairs            main_code/utils/model_utils.py:333-335, 367-369, 392-394
  * Youden-index threshold on the tuning fold          model_utils.py:407-414 (tune_threshold_roc)
  * accuracy at that threshold on the other k-1 folds  model_utils.py:370-374 (evaluate: pred = cos > thr)
  * AUC on the other k-1 folds                         model_utils.py:346-352 (compute_auc)
  * StratifiedKFold(k, shuffle=True, random_state=42)  model_utils.py:439-444 (cross_validate_kfold)

roc_curve / roc_auc_score / StratifiedKFold live in scikit-learn (third-party, not vendored by the reference, no version
pinned in requirement.txt; 1.9 in this image); the reference's own call sites above are what is mirrored.  Parity is
pinned by tests/golden/lfw_synth_*.npz, produced by oracle/make_golden_lfw.py from the reference's own functions.
Only tests/ and the golden generator import this module.
"""
from __future__ import annotations

import numpy as np


def synthetic_pairs(n_pairs: int = 6000, d: int = 512, noise: float = 0.8, seed: int = 5):
    """n_pairs/2 same-identity and n_pairs/2 different-identity pairs, e = id_centre + noise * unit noise
    (SURVEY.md section 8d, cfg5).  Returns (e1, e2, same) as float32 / float32 / int64 numpy arrays (unnormalised)."""
    rng = np.random.default_rng(seed)
    half = n_pairs // 2

    def unit(n):
        v = rng.standard_normal((n, d))
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    ca, cb = unit(half), unit(half)                      # identity centres of the "different" pairs
    cs = unit(half)                                      # shared identity centre of the "same" pairs
    scale = rng.uniform(5.0, 40.0, size=(n_pairs, 2))    # raw embedding norms: the evaluator must normalise them away
    e1 = np.concatenate([cs + noise * unit(half), ca + noise * unit(half)]) * scale[:, :1]
    e2 = np.concatenate([cs + noise * unit(half), cb + noise * unit(half)]) * scale[:, 1:]
    same = np.concatenate([np.ones(half, dtype=np.int64), np.zeros(half, dtype=np.int64)])
    perm = rng.permutation(n_pairs)
    return e1[perm].astype(np.float32), e2[perm].astype(np.float32), same[perm]


def pair_cosine(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """F.normalize(., dim=1) on both sides (eps 1e-12) and the row-wise dot, in float64."""
    a, b = e1.astype(np.float64), e2.astype(np.float64)
    a = a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-12)
    b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-12)
    return (a * b).sum(axis=1)


def tune_threshold_roc(cos: np.ndarray, same: np.ndarray):
    from sklearn.metrics import roc_curve
    fpr, tpr, thr = roc_curve(same, cos)
    best = thr[int(np.argmax(tpr - fpr))]
    acc = 100.0 * ((cos > best).astype(int) == same).sum() / len(same)
    return float(best), float(acc)


def evaluate(cos: np.ndarray, same: np.ndarray, threshold: float) -> float:
    return 100.0 * float(((cos > threshold).astype(np.int64) == same).sum()) / len(same) if len(same) else 0.0


def compute_auc(cos: np.ndarray, same: np.ndarray) -> float:
    from sklearn.metrics import roc_auc_score
    if len(np.unique(same)) < 2:
        return 0.0
    return float(roc_auc_score(same, cos))


def cross_validate_kfold(cos: np.ndarray, same: np.ndarray, k_fold: int = 10):
    from sklearn.model_selection import StratifiedKFold
    skf = StratifiedKFold(n_splits=k_fold, shuffle=True, random_state=42)
    accs, aucs, thrs = [], [], []
    for train_idx, val_idx in skf.split(np.zeros((len(same), 1)), same):
        thr, _ = tune_threshold_roc(cos[val_idx], same[val_idx])
        thrs.append(thr)
        accs.append(evaluate(cos[train_idx], same[train_idx], thr))
        aucs.append(compute_auc(cos[train_idx], same[train_idx]))
    return dict(mean_acc=float(np.mean(accs)), std_acc=float(np.std(accs)), mean_auc=float(np.mean(aucs)),
                std_auc=float(np.std(aucs)), thresholds=np.array(thrs), accs=np.array(accs), aucs=np.array(aucs))
