"""CPU: the oracle restatement reproduces the golden vectors generated from the reference itself."""
import pytest
import torch

from oracle import margin_oracle as mo
from tests.helpers import golden_files, load_golden, rel


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_golden(path):
    g = load_golden(path)
    out = mo.loss_and_grads(g["cfg"], g["state_in"], g["x"], g["W"], g["labels"], margins=g["margins"],
                            lambda_g=g["lambda_g"], grad_scale=g["grad_scale"])
    assert abs(float(out["loss"]) - g["loss"]) <= 1e-7 * abs(g["loss"])          # fp64 vs fp64 (stored fp64)
    assert abs(float(out["loss_g"]) - g["loss_g"]) <= 1e-9
    assert abs(float(out["acc1"]) - g["acc1"]) < 1e-9 and abs(float(out["acc5"]) - g["acc5"]) < 1e-9
    assert rel(out["norms"], g["norms"]) < 1e-12
    assert rel(out["dx"], g["dx"]) < 1e-6                                       # golden grads stored as fp32
    assert rel(out["dW"], g["dW"]) < 1e-6
    ns, so = out["new_state"], g["state_out"]
    assert ns.sphere_iter == so.sphere_iter
    assert abs(ns.t_buf - so.t_buf) < 1e-12
    assert abs(ns.batch_mean - so.batch_mean) < 1e-9 and abs(ns.batch_std - so.batch_std) < 1e-9


def test_golden_set_covers_every_family():
    fams = {load_golden(p)["family"] for p in golden_files()}
    assert fams == set(mo.FAMILIES)


@pytest.mark.parametrize("fam", mo.FAMILIES)
def test_autograd_leg_matches_closed_form(fam):
    """The materialising autograd formulation (CPU-baseline leg) equals the closed-form backward."""
    cfg = mo.HeadConfig.default(fam)
    x, W, labels = mo.make_inputs(fam, 12, 77, 512, seed=5)
    margins = None
    if fam.startswith("elastic"):
        torch.manual_seed(3)
        margins = mo.sample_elastic_margins(cfg, 12)
    lg = 35.0 if fam == "magface" else 0.0
    a = mo.autograd_step(cfg, mo.HeadState(), x, W, labels, margins=margins, lambda_g=lg, dtype=torch.float64)
    b = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, margins=margins, lambda_g=lg)
    assert abs(float(a["loss"]) - float(b["loss"])) < 1e-10 * abs(float(b["loss"]))
    assert rel(a["dx"], b["dx"]) < 1e-9 and rel(a["dW"], b["dW"]) < 1e-9
    assert abs(float(a["acc1"]) - float(b["acc1"])) < 1e-4      # the autograd leg reports fp32 percentages


def test_survey_known_answer_tripwire():
    """SURVEY.md section 4 known-answer row for ArcFace (generated from the reference, torch CPU RNG)."""
    torch.manual_seed(1234)
    W = torch.empty(10575, 512)
    torch.nn.init.xavier_uniform_(W)
    x = torch.randn(64, 512)
    y = torch.randint(0, 10575, (64,))
    out = mo.loss_and_grads(mo.HeadConfig.default("arcface"), mo.HeadState(), x, W, y)
    assert abs(float(out["loss"]) - 43.764915) < 5e-4
    assert abs(float(out["dx"].norm()) - 3.158748e-01) < 1e-5
    assert abs(float(out["dW"].norm()) - 2.362826e+01) < 2e-2   # table value came from an fp32 run


def test_linearity_in_grad_scale():
    cfg = mo.HeadConfig.default("cosface")
    x, W, labels = mo.make_inputs("cosface", 6, 33, 512, seed=9)
    a = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, grad_scale=1.0)
    b = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, grad_scale=1024.0)
    assert rel(b["dx"], a["dx"] * 1024.0) < 1e-12 and rel(b["dW"], a["dW"] * 1024.0) < 1e-12


@pytest.mark.parametrize("name", ["arcface", "cosface", "curricularface", "curricularface_t05", "sphereface_m2",
                                  "sphereface_m4_iter"])
def test_chunked_reference_matches_golden(name):
    """Pins the chunked fp32 checker to the reference: it must reproduce the committed golden (outputs of the
    unmodified reference head + CrossEntropyLoss + autograd) at a small size, evaluated in 3 ragged chunks."""
    import os
    from oracle.chunked_fp32 import chunked_reference, cosine
    from tests.helpers import GOLDEN_DIR, load_golden
    g = load_golden(os.path.join(GOLDEN_DIR, name + ".npz"))
    fam = g["family"]
    assert g["grad_scale"] == 1.0 and not g["cfg"].easy_margin
    x, W, y = g["x"].float(), g["W"].float(), g["labels"]
    cd = mo.LAYOUT[fam] == "CD"
    Wc = W if cd else W.t().contiguous()
    kw = {}
    if fam == "curricularface":
        kw = dict(t_buf=g["state_in"].t_buf, momentum=g["cfg"].momentum)
    if fam == "sphereface":
        kw = dict(sphere_lambda=mo.sphere_lambda(g["state_in"].sphere_iter + 1))
    m = g["cfg"].sphere_m if fam == "sphereface" else g["cfg"].m
    res = chunked_reference(x, Wc, y, family=fam, s=g["cfg"].s, m=m, chunk=(Wc.shape[0] + 2) // 3, **kw)
    loss, dx, dWc = res[:3]
    dW = dWc if cd else dWc.t()
    assert abs(float(loss) - g["loss_id"]) <= 1e-5 * abs(g["loss_id"]), (float(loss), g["loss_id"])
    assert cosine(dx.cpu(), g["dx"]) > 1 - 1e-8 and cosine(dW.cpu(), g["dW"]) > 1 - 1e-8
    assert float((dx.cpu().double() - g["dx"].double()).norm() / g["dx"].double().norm()) < 2e-4
    assert float((dW.cpu().double() - g["dW"].double()).norm() / g["dW"].double().norm()) < 2e-4
    if fam == "curricularface":
        assert abs(res[3] - g["state_out"].t_buf) < 1e-6
