"""GPU: HeadSGD (mh_sgd_step_w) against torch.optim.SGD as the reference builds it (model_utils.py:557, 186), and
the w_hat shadow it leaves for the next forward against mh_prologue_w on the updated parameter."""
import copy
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(kind, Cn, seed=0):
    import face_recognition_models_b200 as pkg
    torch.manual_seed(seed)
    if kind == "arcface":                       # weight [C, D]
        a = pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False).cuda()
    elif kind == "cosface":                     # kernel [D, C]
        a = pkg.CosFace(512, Cn, s=64.0, m=0.35).cuda()
    else:                                       # kernel [D, C], in-place state
        a = pkg.AdaFace(512, Cn).cuda()
    b = copy.deepcopy(a)
    return a, b


def _prologue(head):
    """w_hat / inv_norm of the head's current parameter through mh_prologue_w, into fresh buffers."""
    from face_recognition_models_b200 import _lib as L
    eng = head.head_engine()
    W = head.head_parameter().detach()
    Cn = eng.C
    C_pad = (Cn + L.NTILE - 1) // L.NTILE * L.NTILE
    w_hat = torch.empty(C_pad, 512, dtype=torch.bfloat16, device="cuda")
    inv = torch.empty(Cn, dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.call("mh_prologue_w", C.c_void_p(W.data_ptr()), eng.layout, Cn, W.shape[1], C.c_void_p(w_hat.data_ptr()), C_pad,
           C.c_void_p(0), C.c_void_p(inv.data_ptr()), st)
    return w_hat, inv


def _called(fn):
    from face_recognition_models_b200 import _lib as L
    L.PROFILE = []
    try:
        fn()
        torch.cuda.synchronize()
        names = [r[0] for r in L.PROFILE]
        # the whole-phase entry point runs the W prologue as its first kernel unless told to skip it: 7 launches vs 6
        names += ["mh_prologue_w" for r in L.PROFILE if r[0] == "mh_step_forward" and r[3] == 7]
        return names
    finally:
        L.PROFILE = None


@pytest.mark.parametrize("kind,Cn", [("arcface", 1000), ("arcface", 1003), ("cosface", 1000), ("cosface", 1003),
                                     ("adaface", 2052)])
def test_trajectory_equals_torch_sgd(kind, Cn):
    """Same init, same batches: W, the momentum buffer and the loss follow torch.optim.SGD step for step, and each
    forward after a step runs on the shadow (no mh_prologue_w) whose bits equal mh_prologue_w(W_new)."""
    import face_recognition_models_b200 as pkg
    ref, fus = _pair(kind, Cn)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4, foreach=False)
    opt_fus = pkg.HeadSGD([fus], lr=0.05, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator(device="cuda").manual_seed(7)
    for step in range(4):
        x = torch.randn(96, 512, device="cuda", generator=g)
        y = torch.randint(0, Cn, (96,), device="cuda", generator=g)
        opt_ref.zero_grad(set_to_none=True)
        opt_fus.zero_grad(set_to_none=True)
        lr_ = ref.fused_loss(x, y).loss
        lr_.backward()
        box = {}

        def run():
            box["loss"] = fus.fused_loss(x, y).loss
            box["loss"].backward()
        names = _called(run)
        lf = box["loss"]
        assert ("mh_prologue_w" in names) == (step == 0), names       # steps 1.. run on the shadow
        assert abs(lr_.item() - lf.item()) <= 1e-6 * abs(lr_.item()), (step, lr_.item(), lf.item())
        opt_ref.step()
        opt_fus.step()
        Wr, Wf = ref.head_parameter().detach(), fus.head_parameter().detach()
        mr = opt_ref.state[ref.head_parameter()]["momentum_buffer"]
        mf = opt_fus.state[fus.head_parameter()]["momentum_buffer"]
        # same operation order and roundings as torch's kernels: expected bit-equal; the bound allows one fp32 ulp
        assert (Wr - Wf).abs().max().item() <= 2.0 ** -23 * Wr.abs().max().item(), step
        assert (mr - mf).abs().max().item() <= 2.0 ** -23 * mr.abs().max().item(), step
        w_hat, inv = _prologue(fus)
        eng = fus.head_engine()
        assert torch.equal(eng._ws["w_hat"], w_hat), step
        assert torch.equal(eng._ws["inv_norm"], inv), step


def test_shadow_is_dropped_when_w_changes_elsewhere():
    import face_recognition_models_b200 as pkg
    _, head = _pair("arcface", 512)
    opt = pkg.HeadSGD([head], lr=0.1)
    x = torch.randn(32, 512, device="cuda")
    y = torch.randint(0, 512, (32,), device="cuda")
    head.fused_loss(x, y).loss.backward()
    opt.step()
    assert "mh_prologue_w" not in _called(lambda: head.fused_loss(x, y))
    with torch.no_grad():
        head.weight.mul_(0.5)                   # any other in-place write bumps the version counter
    assert "mh_prologue_w" in _called(lambda: head.fused_loss(x, y))
    opt.zero_grad()
    head.fused_loss(x, y).loss.backward()
    opt.step()
    head.mode = "exact"                         # the exact path needs the fp32 w_hat: prologue, shadow dropped
    assert "mh_prologue_w" in _called(lambda: head.fused_loss(x, y))
    head.mode = "tc"
    assert "mh_prologue_w" in _called(lambda: head.fused_loss(x, y))
    # a write through .data is invisible to the version counter: the documented escape hatch
    opt.zero_grad()
    head.fused_loss(x, y).loss.backward()
    opt.step()
    head.weight.data.mul_(2.0)
    head.head_engine().invalidate_shadow()
    assert "mh_prologue_w" in _called(lambda: head.fused_loss(x, y))


def test_gradscaler_protocol_and_inf_skip():
    """GradScaler.step hands grad_scale / found_inf to the optimizer: the update unscales in the kernel, and a
    non-finite gradient leaves W and the momentum buffer untouched while the shadow stays valid."""
    import face_recognition_models_b200 as pkg
    ref, fus = _pair("cosface", 1000, seed=3)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4, foreach=False)
    opt_fus = pkg.HeadSGD([fus], lr=0.05, momentum=0.9, weight_decay=5e-4)
    sc_ref = torch.amp.GradScaler("cuda", init_scale=4096.0)
    sc_fus = torch.amp.GradScaler("cuda", init_scale=4096.0)
    x = torch.randn(64, 512, device="cuda")
    y = torch.randint(0, 1000, (64,), device="cuda")
    for _ in range(3):
        for head, opt, sc in ((ref, opt_ref, sc_ref), (fus, opt_fus, sc_fus)):
            opt.zero_grad(set_to_none=True)
            sc.scale(head.fused_loss(x, y).loss).backward()
            sc.step(opt)
            sc.update()
    Wr, Wf = ref.kernel.detach(), fus.kernel.detach()
    assert (Wr - Wf).abs().max().item() <= 1e-6 * Wr.abs().max().item()
    # poison the gradient: the step must be skipped on the device, without a host sync
    before = fus.kernel.detach().clone()
    mom_before = opt_fus.state[fus.kernel]["momentum_buffer"].clone()
    opt_fus.zero_grad(set_to_none=True)
    sc_fus.scale(fus.fused_loss(x, y).loss).backward()
    fus.kernel.grad[3, 5] = float("inf")
    sc_fus.step(opt_fus)
    sc_fus.update()
    assert torch.equal(fus.kernel.detach(), before)
    assert torch.equal(opt_fus.state[fus.kernel]["momentum_buffer"], mom_before)
    assert sc_fus.get_scale() == 2048.0
    w_hat, inv = _prologue(fus)
    assert torch.equal(fus.head_engine()._ws["w_hat"], w_hat)
    assert "mh_prologue_w" not in _called(lambda: fus.fused_loss(x, y))


def test_state_dict_interchanges_with_torch_sgd_and_lr_scheduler():
    import face_recognition_models_b200 as pkg
    ref, fus = _pair("arcface", 640, seed=5)
    opt_fus = pkg.HeadSGD([fus], lr=0.1, momentum=0.9, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt_fus, step_size=1, gamma=0.1)
    x = torch.randn(32, 512, device="cuda")
    y = torch.randint(0, 640, (32,), device="cuda")
    fus.fused_loss(x, y).loss.backward()
    opt_fus.step()
    sched.step()
    assert abs(opt_fus.param_groups[0]["lr"] - 0.01) < 1e-12
    # continue the run in torch.optim.SGD from HeadSGD's state
    ref.load_state_dict(fus.state_dict())
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4, foreach=False)
    sd = opt_fus.state_dict()
    opt_ref.load_state_dict({"state": copy.deepcopy(sd["state"]), "param_groups": [dict(opt_ref.state_dict()["param_groups"][0],
                                                                            lr=sd["param_groups"][0]["lr"])]})
    for head, opt in ((ref, opt_ref), (fus, opt_fus)):
        opt.zero_grad(set_to_none=True)
        head.fused_loss(x, y).loss.backward()
        opt.step()
    assert (ref.weight - fus.weight).abs().max().item() <= 2.0 ** -23 * ref.weight.abs().max().item()


def test_error_paths():
    import face_recognition_models_b200 as pkg
    from face_recognition_models_b200 import _lib as L
    with pytest.raises(TypeError):
        pkg.HeadSGD([torch.nn.Linear(4, 4)], lr=0.1)
    with pytest.raises(ValueError):
        pkg.HeadSGD([], lr=0.1)
    cpu_head = pkg.ArcFace(512, 64)
    opt = pkg.HeadSGD([cpu_head], lr=0.1)
    cpu_head.weight.grad = torch.zeros_like(cpu_head.weight)
    with pytest.raises(L.MarginHeadError):
        opt.step()


def test_coupled_found_inf_skips_the_head_step_too():
    """GradScaler decides the inf-skip per optimizer.  With coupled=[opt_backbone] + couple_scaler(scaler) an
    overflow seen only in the BACKBONE gradients also skips the head update (what the reference's single optimizer does),
    without a host sync; without the coupling the head would step."""
    import face_recognition_models_b200 as pkg
    torch.manual_seed(5)
    lin = torch.nn.Linear(512, 512).cuda()                              # stands in for the backbone
    x = torch.randn(32, 512, device="cuda")
    y = torch.randint(0, 1000, (32,), device="cuda")
    for coupled in (True, False):
        head = pkg.ArcFace(512, 1000, s=64.0, m=0.5, easy_margin=False).cuda()
        opt_b = torch.optim.SGD(lin.parameters(), lr=0.01, momentum=0.9)
        opt_h = pkg.HeadSGD([head], lr=0.05, momentum=0.9, weight_decay=5e-4, coupled=[opt_b] if coupled else ())
        scaler = torch.amp.GradScaler("cuda", init_scale=256.0)
        if coupled:
            opt_h.couple_scaler(scaler)
        w0 = head.weight.detach().clone()
        opt_b.zero_grad(set_to_none=True)
        opt_h.zero_grad(set_to_none=True)
        out = head.fused_loss(lin(x), y)
        scaler.scale(out.loss).backward()
        lin.weight.grad[0, 0] = float("inf")                            # overflow in the backbone gradients only
        assert bool(torch.isfinite(head.weight.grad).all())
        scaler.unscale_(opt_b)                                           # records opt_b's found_inf (must precede the head step)
        scaler.step(opt_b)
        scaler.step(opt_h)
        scaler.update()
        moved = not torch.equal(head.weight.detach(), w0)
        assert moved == (not coupled), (coupled, moved)
        assert float(scaler.get_scale()) == 128.0                        # the overflow halves the scale either way
