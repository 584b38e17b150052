"""Generate tests/golden/lfw_synth_*.npz from the UNMODIFIED reference evaluator (build container only).

    python oracle/make_golden_lfw.py            # needs /root/reference

The reference's tune_threshold_roc / evaluate / compute_auc (main_code/utils/model_utils.py:320-414) take a model and a
dataset of image pairs.  Here the "model" is nn.Identity() and the dataset a TensorDataset of synthetic embedding
pairs, so the functions run unmodified on CPU; the k-fold driver cross_validate_kfold (model_utils.py:416-474) reads
image files, so its protocol (StratifiedKFold(10, shuffle=True, random_state=42); tune on the held-out fold, score the
other nine) is replayed around the reference's per-fold functions.  roc_auc_score is never imported by the reference
(model_utils.py:352 vs the import at :14), so it is injected into the module namespace here -- not into the reference.
`alive_progress` (imported by utils/dataset.py, absent in this image) is stubbed.  The oracle restatement is asserted
against these outputs before they are stored.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import verification_oracle as vo  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def load_reference():
    stub = types.ModuleType("alive_progress")
    stub.alive_bar = lambda *a, **k: None
    sys.modules.setdefault("alive_progress", stub)
    sys.path.insert(0, os.path.join(REF, "main_code"))
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        from main_code.utils import model_utils as mu  # type: ignore
    from sklearn.metrics import roc_auc_score
    mu.roc_auc_score = roc_auc_score
    return mu


def main():
    mu = load_reference()
    from sklearn.model_selection import StratifiedKFold
    from torch.utils.data import TensorDataset
    model = torch.nn.Identity()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, noise, n_pairs, seed in (("lfw_synth_easy", 0.8, 6000, 5), ("lfw_synth_hard", 2.5, 6000, 6),
                                       ("lfw_synth_small", 1.5, 600, 7)):
        e1, e2, same = vo.synthetic_pairs(n_pairs, 512, noise, seed)
        t1, t2, ts = torch.from_numpy(e1), torch.from_numpy(e2), torch.from_numpy(same)
        skf = StratifiedKFold(n_splits=10, shuffle=True, random_state=42)
        accs, aucs, thrs = [], [], []
        with contextlib.redirect_stdout(io.StringIO()):
            for train_idx, val_idx in skf.split(np.zeros((n_pairs, 1)), same):
                val = TensorDataset(t1[val_idx], t2[val_idx], ts[val_idx])
                train = TensorDataset(t1[train_idx], t2[train_idx], ts[train_idx])
                thr, _ = mu.tune_threshold_roc(model, val, 64, "cpu")
                thrs.append(float(thr))
                accs.append(mu.evaluate(model, train, 64, "cpu", thr))
                aucs.append(mu.compute_auc(model, train, 64, "cpu"))
        ref = dict(mean_acc=np.mean(accs), std_acc=np.std(accs), mean_auc=np.mean(aucs), std_auc=np.std(aucs))
        mine = vo.cross_validate_kfold(vo.pair_cosine(e1, e2), same, 10)
        # the reference works on float32 cosines; a pair whose cosine sits within 1e-6 of a threshold may flip
        assert np.allclose(mine["thresholds"], thrs, atol=2e-6), (mine["thresholds"], thrs)
        assert np.allclose(mine["accs"], accs, atol=100.0 * 2 / (0.9 * n_pairs)), (mine["accs"], accs)
        assert np.allclose(mine["aucs"], aucs, atol=1e-6)
        np.savez(os.path.join(out_dir, name + ".npz"), n_pairs=n_pairs, noise=noise, seed=seed, d=512,
                 thresholds=np.array(thrs), accs=np.array(accs), aucs=np.array(aucs),
                 e1_sum=float(e1.astype(np.float64).sum()), e2_sum=float(e2.astype(np.float64).sum()), **ref)
        print(f"{name}: acc {ref['mean_acc']:.3f} +- {ref['std_acc']:.3f}  auc {ref['mean_auc']:.5f} +- {ref['std_auc']:.5f}  "
              f"thr[0] {thrs[0]:.4f}")


if __name__ == "__main__":
    main()
