"""GPU: the guarded stash of the step API (mh_step_forward / mh_step_backward with stash == 2, csrc/step.cu).

CurricularFace (criterion.py:491-587) and SphereFace (criterion.py:12-107) fail the static proof behind the forward stash
(mh_tc_stash_ok), so they used to pay a fourth GEMM pass (backward-G recompute).  The guarded stash runs the
fixed-reference forward + stash speculatively, checks on the device that no row sum can have lost anything to underflow,
and re-runs the general path through gated launches when the check fails.  Both outcomes must meet the parity bar against
the oracle; the fallback must actually trigger on inputs built to underflow, and never on ordinary ones."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(fam, B, Cn, seed, adversarial):
    from oracle import margin_oracle as mo
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=seed)
    if adversarial and fam == "curricularface":
        # Every logit of every row far below the reference (s*2): all rows carry label 0, whose centre is orthogonal to
        # the samples (target cos ~ 0 -> cos(theta + m) ~ -0.48, z2 = -44), every other centre points away from them
        # (cos ~ -0.66 <= the hard-negative threshold ~ -0.5, so they stay "easy": z2 = -61).  With ref2 = 82.7 every term
        # is below 2^-126 and flushes, while the loss itself is O(1) (the many negatives weigh about as much as the
        # target).  (A single ordinary class per row would already make the row safe - and it is.)
        g = torch.Generator().manual_seed(seed + 1)
        v = torch.nn.functional.normalize(torch.randn(512, generator=g), dim=0)
        x = 20.0 * torch.nn.functional.normalize(v + 0.3 * torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1), dim=1)
        Wc = 0.01 * torch.nn.functional.normalize(-0.69 * v + 0.724 * torch.nn.functional.normalize(torch.randn(Cn, 512, generator=g), dim=1), dim=1)
        w0 = torch.randn(512, generator=g)
        Wc[0] = 0.01 * torch.nn.functional.normalize(w0 - (w0 @ v) * v, dim=0)
        labels = torch.zeros_like(labels)
        W = Wc if W.shape == Wc.shape else Wc.t().contiguous()
    if adversarial and fam == "sphereface":
        x = x * 6.0                  # |x| up to ~680: logits span ~2000 binades, ordinary cosines sit ~700 binades below the reference
    return x, W, labels


def _run(fam, bmode, B, Cn, seed=33, adversarial=False, step_api="1", monkeypatch=None):
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import build_head, prime_head
    if monkeypatch is not None:
        monkeypatch.setenv("MH_STEP_API", step_api)
    cfg = mo.HeadConfig.default(fam)
    state = mo.HeadState(t_buf=0.3) if fam == "curricularface" else mo.HeadState()
    x, W, labels = _inputs(fam, B, Cn, seed, adversarial)
    head = prime_head(build_head(pkg, fam, cfg, Cn).cuda(), fam, W, state, None)
    head.backward_mode = bmode
    xg = x.cuda().requires_grad_(True)
    out = head.fused_loss(xg, labels.cuda())
    out.loss.backward()
    torch.cuda.synchronize()
    T = (head._engine._step_T or {}) if step_api == "1" else {k[len("stash_"):]: v for k, v in head._engine._ws.items() if k == "stash_guard"}
    guard = int(T["guard"].item()) if "guard" in T else None
    return dict(loss=out.loss.detach().clone(), acc1=out.acc1.clone(), acc5=out.acc5.clone(), dx=xg.grad.clone(),
                dW=head._param().grad.clone(), guard=guard, cfg=cfg, state=state, x=x, W=W, labels=labels)


def _check_against_oracle(r, rel_tol=1e-2):
    from oracle import margin_oracle as mo
    from tests.helpers import cosim, rel
    ref = mo.loss_and_grads(r["cfg"], r["state"], r["x"], r["W"], r["labels"])
    assert abs(float(r["loss"]) - float(ref["loss_id"])) <= 2e-3 * abs(float(ref["loss_id"]))
    assert abs(float(r["acc1"]) - float(ref["acc1"])) < 0.6 and abs(float(r["acc5"]) - float(ref["acc5"])) < 0.6
    assert cosim(r["dx"], ref["dx"]) > 0.9995 and cosim(r["dW"], ref["dW"]) > 0.9995
    assert rel(r["dx"], ref["dx"]) < rel_tol and rel(r["dW"], ref["dW"]) < rel_tol


@pytest.mark.parametrize("fam", ["curricularface", "sphereface"])
@pytest.mark.parametrize("B,Cn", [(300, 4097), (700, 160_001)])
def test_guarded_stash_matches_oracle(fam, B, Cn, monkeypatch):
    """backward_mode='stash' takes the guarded stash at any size; (700, 160001) also takes it under 'auto' and is eligible
    for the merged dx + dW kernel.  Ordinary inputs: the guard must stay down."""
    r = _run(fam, "stash", B, Cn, monkeypatch=monkeypatch)
    assert r["guard"] == 0
    _check_against_oracle(r)
    again = _run(fam, "stash", B, Cn, monkeypatch=monkeypatch)
    assert torch.equal(r["loss"], again["loss"]) and torch.equal(r["dx"], again["dx"]) and torch.equal(r["dW"], again["dW"])
    # same step in recompute mode: same loss statistics up to the summation order, gradients within bf16 rounding of G
    from tests.helpers import rel
    rc = _run(fam, "recompute", B, Cn, monkeypatch=monkeypatch)
    assert rc["guard"] is None
    assert abs(float(r["loss"]) - float(rc["loss"])) <= 1e-5 * abs(float(rc["loss"]))
    assert torch.equal(r["acc1"], rc["acc1"]) and torch.equal(r["acc5"], rc["acc5"])
    assert rel(r["dx"], rc["dx"]) < 5e-3 and rel(r["dW"], rc["dW"]) < 5e-3


def test_auto_mode_takes_the_guarded_stash_only_at_scale(monkeypatch):
    small = _run("curricularface", "auto", 300, 4097, monkeypatch=monkeypatch)
    assert small["guard"] is None                       # B_pad * C_pad < 2^25: recompute, as before
    big = _run("curricularface", "auto", 700, 160_001, monkeypatch=monkeypatch)
    assert big["guard"] == 0
    monkeypatch.setenv("MH_STASH_GUARDED", "0")
    off = _run("curricularface", "auto", 700, 160_001, monkeypatch=monkeypatch)
    assert off["guard"] is None


@pytest.mark.parametrize("fam", ["curricularface", "sphereface"])
@pytest.mark.parametrize("B,Cn,step_api", [(300, 4097, "1"), (520, 160_001, "1"), (300, 4097, "0")])
def test_guarded_stash_falls_back_on_underflow(fam, B, Cn, step_api, monkeypatch):
    """Inputs whose fixed-reference terms all flush to zero: the device-side guard must rise, the gated general forward and
    the gated backward-G must take over, and the results must still meet the parity bar (both drivers: the whole-phase
    entry points and the per-kernel sequence the class-sharded head uses)."""
    r = _run(fam, "stash", B, Cn, adversarial=True, step_api=step_api, monkeypatch=monkeypatch)
    assert r["guard"] == 1
    assert torch.isfinite(r["loss"]) and torch.isfinite(r["dx"]).all() and torch.isfinite(r["dW"]).all()
    # SphereFace at |x| ~ 700: the softmax is saturated, a bf16 cosine error of 5e-5 is 0.03 nats on the few classes that
    # share the probability mass -- the same figure in recompute mode (compared below), hence the wider relative bound
    _check_against_oracle(r, rel_tol=2e-2 if fam == "sphereface" else 1e-2)
    # the fallback is the recompute path: same loss bits, gradients equal up to the order of the projection sums
    from tests.helpers import rel
    rc = _run(fam, "recompute", B, Cn, adversarial=True, step_api=step_api, monkeypatch=monkeypatch)
    assert torch.equal(r["loss"], rc["loss"]) and torch.equal(r["acc1"], rc["acc1"])
    assert rel(r["dx"], rc["dx"]) < 1e-5 and rel(r["dW"], rc["dW"]) < 2e-3


def test_guarded_stash_single_gradient(monkeypatch):
    """Only dx (frozen class centres) or only dW (detached embeddings) wanted: the non-merged guarded path."""
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import build_head, cosim, prime_head
    monkeypatch.setenv("MH_STEP_API", "1")
    fam, B, Cn = "curricularface", 300, 70_001
    cfg = mo.HeadConfig.default(fam)
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=5)
    ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels)
    for want in ("dx", "dW"):
        head = prime_head(build_head(pkg, fam, cfg, Cn).cuda(), fam, W, mo.HeadState(), None)
        head.backward_mode = "stash"
        head._param().requires_grad_(want == "dW")
        xg = x.cuda().requires_grad_(want == "dx")
        out = head.fused_loss(xg, labels.cuda())
        out.loss.backward()
        torch.cuda.synchronize()
        assert int(head._engine._step_T["guard"].item()) == 0
        got = xg.grad if want == "dx" else head._param().grad
        assert cosim(got, ref[want]) > 0.9995
