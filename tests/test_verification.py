"""Pair-verification row: the reference's LFW 10-fold protocol on synthetic embedding pairs.

CPU: the oracle restatement and the product's host-side protocol reproduce the goldens generated from the reference's own
tune_threshold_roc / evaluate / compute_auc (oracle/make_golden_lfw.py).  GPU: mh_pair_cosine against the oracle, and the
whole pipeline against the goldens.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import verification_oracle as vo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lfw_synth_*.npz")))


def load(path):
    z = np.load(path)
    e1, e2, same = vo.synthetic_pairs(int(z["n_pairs"]), int(z["d"]), float(z["noise"]), int(z["seed"]))
    # the regenerated pairs must be the ones the reference saw
    assert abs(float(e1.astype(np.float64).sum()) - float(z["e1_sum"])) < 1e-6 * max(1.0, abs(float(z["e1_sum"])))
    assert abs(float(e2.astype(np.float64).sum()) - float(z["e2_sum"])) < 1e-6 * max(1.0, abs(float(z["e2_sum"])))
    return z, e1, e2, same


def check_against_golden(z, mean_acc, std_acc, mean_auc, std_auc):
    n_scored = 0.9 * int(z["n_pairs"])
    flip = 100.0 * 2 / n_scored                       # a pair within float rounding of a threshold may flip
    assert abs(mean_acc - float(z["mean_acc"])) <= flip and abs(std_acc - float(z["std_acc"])) <= flip
    assert abs(mean_auc - float(z["mean_auc"])) <= 1e-5 and abs(std_auc - float(z["std_auc"])) <= 1e-5


def test_goldens_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_golden(path):
    z, e1, e2, same = load(path)
    r = vo.cross_validate_kfold(vo.pair_cosine(e1, e2), same, 10)
    assert np.allclose(r["thresholds"], z["thresholds"], atol=2e-6)
    assert np.allclose(r["aucs"], z["aucs"], atol=1e-6)
    check_against_golden(z, r["mean_acc"], r["std_acc"], r["mean_auc"], r["std_auc"])


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_host_protocol_matches_reference_golden(path):
    """The product's tune_threshold_roc / evaluate / compute_auc on oracle cosines (no GPU): same folds, same numbers."""
    from sklearn.model_selection import StratifiedKFold
    from face_recognition_models_b200 import verification as V
    z, e1, e2, same = load(path)
    cos = vo.pair_cosine(e1, e2).astype(np.float32)
    accs, aucs, thrs = [], [], []
    for tr, va in StratifiedKFold(n_splits=10, shuffle=True, random_state=42).split(np.zeros((len(same), 1)), same):
        thr, tune_acc = V.tune_threshold_roc(cos[va], same[va])
        assert 0.0 <= tune_acc <= 100.0
        thrs.append(thr)
        accs.append(V.evaluate(torch.from_numpy(cos[tr]), torch.from_numpy(same[tr]), thr))
        aucs.append(V.compute_auc(cos[tr], same[tr]))
    assert np.allclose(thrs, z["thresholds"], atol=2e-6)
    check_against_golden(z, float(np.mean(accs)), float(np.std(accs)), float(np.mean(aucs)), float(np.std(aucs)))


def test_edge_cases_host():
    from face_recognition_models_b200 import verification as V
    assert V.compute_auc(np.array([0.1, 0.2]), np.array([1, 1])) == 0.0          # one class only (model_utils.py:349-350)
    assert V.evaluate(np.array([]), np.array([])) == 0.0                          # empty (model_utils.py:377)
    assert V.evaluate(np.array([0.5, 0.1]), np.array([1, 0]), 0.33) == 100.0      # default threshold of the reference
    with pytest.raises(Exception):
        V.pair_cosine(torch.zeros(2, 4), torch.zeros(2, 4))                       # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,d,tol", [(torch.float32, 512, 2e-6), (torch.float32, 77, 2e-6), (torch.bfloat16, 512, 1e-6),
                                          (torch.float16, 512, 1e-6)])
def test_pair_cosine_kernel(dtype, d, tol):
    from face_recognition_models_b200 import verification as V
    e1, e2, _ = vo.synthetic_pairs(1000, d, 1.5, 9)
    a, b = torch.from_numpy(e1).cuda().to(dtype), torch.from_numpy(e2).cuda().to(dtype)
    a[3].zero_()                                                                   # zero embedding: cos = 0, not NaN
    got = V.pair_cosine(a, b).cpu().numpy()
    ref = vo.pair_cosine(a.float().cpu().numpy(), b.float().cpu().numpy())         # oracle on the same (rounded) inputs
    assert got.dtype == np.float32 and np.isfinite(got).all() and got[3] == 0.0
    assert np.abs(got - ref).max() <= tol + 1e-6
    # non-contiguous rows (a column slice of a wider buffer) go through the row pitch
    wide = torch.randn(1000, d + 8, device="cuda").to(dtype)
    wide[:, :d] = a
    assert torch.equal(V.pair_cosine(wide[:, :d], b), V.pair_cosine(a, b))
    assert V.pair_cosine(a[:0], b[:0]).numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_kfold_matches_reference_golden(path):
    from face_recognition_models_b200 import verification as V
    z, e1, e2, same = load(path)
    r = V.cross_validate_kfold(torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda(), torch.from_numpy(same), 10)
    check_against_golden(z, *r)
