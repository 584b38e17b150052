"""GPU: the whole-phase entry points (mh_step_forward / mh_step_backward, csrc/step.cu) run the same kernels as the
entry-point-by-entry-point driver, so the results must be bit-identical; the self-projecting dW kernel
(MH_DW_SELFPROJ=1, mh_tc_backward_dw_proj) must meet the same parity bar against the oracle and be bit-reproducible."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(fam, bmode, B, Cn, seed=21):
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import build_head, prime_head
    cfg = mo.HeadConfig.default(fam)
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=seed)
    margins = None
    if fam.startswith("elastic"):
        torch.manual_seed(99)
        margins = mo.sample_elastic_margins(cfg, B)
    head = prime_head(build_head(pkg, fam, cfg, Cn).cuda(), fam, W, mo.HeadState(), margins)
    head.backward_mode = bmode
    xg = x.cuda().requires_grad_(True)
    out = head.fused_loss(xg, labels.cuda())
    (out.loss + 0.5 * out.loss_g).backward()
    torch.cuda.synchronize()
    return (out.loss.detach().clone(), out.acc1.clone(), out.acc5.clone(), xg.grad.clone(), head._param().grad.clone(),
            (cfg, x, W, labels, margins))


@pytest.mark.parametrize("fam,bmode", [("arcface", "auto"), ("arcface", "recompute"), ("cosface", "auto"),
                                       ("curricularface", "auto"), ("sphereface", "auto"), ("magface", "auto"),
                                       ("mv_am", "auto"), ("elastic_arc", "auto"), ("adaface", "auto"),
                                       ("curricularface", "stash"), ("sphereface", "stash")])      # "stash" here = guarded stash
def test_step_api_is_bit_identical_to_the_per_kernel_driver(fam, bmode, monkeypatch):
    monkeypatch.setenv("MH_STEP_API", "1")
    a = _run(fam, bmode, 300, 4097)
    monkeypatch.setenv("MH_STEP_API", "0")
    b = _run(fam, bmode, 300, 4097)
    for u, v in zip(a[:4], b[:4]):                  # loss, acc@1, acc@5, dx: bit-identical in every mode
        assert torch.equal(u, v), fam
    recompute = bmode == "recompute" or (fam in ("curricularface", "sphereface") and bmode == "auto")   # small shape: auto = recompute
    if recompute:
        # the backward-G kernel accumulates the projection sums r_j with fp32 atomics: dW is order-dependent at ~1e-6
        from tests.helpers import rel
        assert rel(a[4], b[4]) < 1e-5, fam
    else:
        assert torch.equal(a[4], b[4]), fam


@pytest.mark.parametrize("fam,bmode", [("arcface", "auto"), ("arcface", "recompute"), ("cosface", "auto"),
                                       ("curricularface", "auto"), ("mv_am", "auto")])
@pytest.mark.parametrize("step_api", ["1", "0"])
def test_self_projecting_dw_matches_oracle(fam, bmode, step_api, monkeypatch):
    from oracle import margin_oracle as mo
    from tests.helpers import cosim, rel
    monkeypatch.setenv("MH_DW_SELFPROJ", "1")
    monkeypatch.setenv("MH_STEP_API", step_api)
    B, Cn = 300, 70_001                         # 274 class tiles: every CTA pair exchanges partials over several tiles
    loss, a1, a5, dx, dW, (cfg, x, W, labels, margins) = _run(fam, bmode, B, Cn)
    ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, margins=margins, lambda_g=0.5)
    assert abs(float(loss) - float(ref["loss_id"])) < 2e-3 * abs(float(ref["loss_id"]))
    assert cosim(dx, ref["dx"]) > 0.9995 and cosim(dW, ref["dW"]) > 0.9995
    assert rel(dW, ref["dW"]) < 1e-2
    # bit-reproducible in BOTH backward modes (fixed-order sum of four partial dots, no atomics)
    again = _run(fam, bmode, B, Cn)
    assert torch.equal(dW, again[4]) and torch.equal(dx, again[3])
    # and consistent with the default projection path
    monkeypatch.setenv("MH_DW_SELFPROJ", "0")
    base = _run(fam, bmode, B, Cn)
    assert rel(dW, base[4]) < 2e-3 and torch.equal(dx, base[3])


@pytest.mark.parametrize("fam,bmode,B", [("arcface", "auto", 300), ("arcface", "recompute", 300), ("cosface", "auto", 1024),
                                         ("curricularface", "auto", 700)])
def test_merged_dx_dw_kernel_matches_oracle(fam, bmode, B, monkeypatch):
    """MH_BWD_MERGED=1: the dx GEMM (interleaved class chunks) and the self-projecting dW GEMM as two roles of one
    persistent kernel that throttle each other.  Same parity bar, bit-reproducible, dx bit-equal to the split kernels'
    sum only up to the different split count (compared by tolerance)."""
    from oracle import margin_oracle as mo
    from tests.helpers import cosim, rel
    Cn = 160_001                                  # 626 class tiles >= 8 per CTA pair: eligible for the merged kernel
    monkeypatch.setenv("MH_BWD_MERGED", "1")
    loss, a1, a5, dx, dW, (cfg, x, W, labels, margins) = _run(fam, bmode, B, Cn)
    ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, margins=margins, lambda_g=0.5)
    assert abs(float(loss) - float(ref["loss_id"])) < 2e-3 * abs(float(ref["loss_id"]))
    assert cosim(dx, ref["dx"]) > 0.9995 and cosim(dW, ref["dW"]) > 0.9995
    assert rel(dW, ref["dW"]) < 1e-2 and rel(dx, ref["dx"]) < 1e-2
    again = _run(fam, bmode, B, Cn)
    assert torch.equal(dW, again[4]) and torch.equal(dx, again[3])
    monkeypatch.setenv("MH_BWD_MERGED", "0")
    base = _run(fam, bmode, B, Cn)
    assert rel(dW, base[4]) < 2e-3 and rel(dx, base[3]) < 1e-4


@pytest.mark.parametrize("fam,bmode,B", [("arcface", "auto", 300), ("arcface", "recompute", 1024), ("mv_am", "auto", 700),
                                         ("sphereface", "auto", 300)])
def test_merged_prologue_forward_kernel(fam, bmode, B, monkeypatch):
    """MH_FWD_MERGED=1: the W prologue as a role of the forward launch ([C, 512] heads).  w^ / inv_norm must carry the same
    bits as the stand-alone prologue, the step the same results as the two-kernel path, and the oracle's within the bar."""
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import cosim, rel
    Cn = 160_001
    monkeypatch.setenv("MH_FWD_MERGED", "1")
    loss, a1, a5, dx, dW, (cfg, x, W, labels, margins) = _run(fam, bmode, B, Cn)
    ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels, margins=margins, lambda_g=0.5)
    assert abs(float(loss) - float(ref["loss_id"])) < 2e-3 * abs(float(ref["loss_id"]))
    assert abs(float(a1) - float(ref["acc1"])) < 0.6
    assert cosim(dx, ref["dx"]) > 0.9995 and cosim(dW, ref["dW"]) > 0.9995
    again = _run(fam, bmode, B, Cn)
    assert torch.equal(loss, again[0]) and torch.equal(dx, again[3])
    monkeypatch.setenv("MH_FWD_MERGED", "0")
    base = _run(fam, bmode, B, Cn)
    assert abs(float(loss) - float(base[0])) <= 1e-6 * abs(float(base[0]))
    assert torch.equal(a1, base[1]) and torch.equal(a5, base[2])
    assert rel(dx, base[3]) < 1e-5 and rel(dW, base[4]) < 1e-5


@pytest.mark.parametrize("B", [2048, 4096, 8192])
def test_merged_backward_at_multi_gpu_row_counts(B):
    """The per-rank shapes of the class-sharded head at 2 / 4 / 8 GPUs (B_g = 2048 / 4096 / 8192 rows: 8 / 16 / 32 row tiles,
    dx role split 4 / 2 / 1) through the merged dx + dW kernel on one GPU, against the chunked fp32 restatement."""
    import face_recognition_models_b200 as pkg
    from oracle.chunked_fp32 import chunked_reference, cosine
    Cn = 160_001
    g = torch.Generator(device="cuda").manual_seed(B)
    head = pkg.ArcFace(512, Cn, s=64.0, m=0.5, easy_margin=False).cuda()
    with torch.no_grad():
        head.weight.normal_(0, 0.01, generator=g)
    y = torch.randint(0, Cn, (B,), device="cuda", generator=g)
    x = torch.randn(B, 512, device="cuda", generator=g)
    near = torch.arange(B, device="cuda") % 2 == 0
    with torch.no_grad():
        x[near] = 20.0 * torch.nn.functional.normalize(
            torch.nn.functional.normalize(head.weight[y[near]], dim=1) + torch.nn.functional.normalize(x[near], dim=1), dim=1)
    x.requires_grad_(True)
    out = head.fused_loss(x, y)
    out.loss.backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        loss, dx, dW = chunked_reference(x.detach(), head.weight.detach(), y, "arcface", 64.0, 0.5, chunk=20_000)
    assert abs(float(out.loss) - float(loss)) <= 2e-3 * abs(float(loss))
    assert cosine(x.grad, dx) >= 0.9995 and cosine(head.weight.grad, dW) >= 0.9995
    assert abs(float(x.grad.norm()) / float(dx.norm()) - 1.0) <= 2e-3
    assert abs(float(head.weight.grad.norm()) / float(dW.norm()) - 1.0) <= 2e-3


def test_prefetch_runs_the_prologue_ahead_and_never_serves_a_stale_one():
    """head.prefetch(): the W prologue of the next forward enqueued early (before the batch copy / beside the backbone).
    Same bits as without it; a parameter update between prefetch and forward invalidates it."""
    import face_recognition_models_b200 as pkg
    g = torch.Generator(device="cuda").manual_seed(3)
    head = pkg.ArcFace(512, 30_011).cuda()
    x = torch.randn(200, 512, device="cuda", generator=g)
    y = torch.randint(0, 30_011, (200,), device="cuda", generator=g)

    def step(prefetch, update=False):
        head.weight.grad = None
        if prefetch:
            head.prefetch()
        if update:
            with torch.no_grad():
                head.weight.mul_(1.0 + torch.linspace(0, 1, 30_011, device="cuda").unsqueeze(1))   # changes every direction's norm, and w_0 .. w_C differently
                head.weight[:, 0] += 0.05
        xg = x.detach().requires_grad_(True)
        out = head.fused_loss(xg, y)
        out.loss.backward()
        return out.loss.detach().clone(), xg.grad.clone(), head.weight.grad.clone()

    a = step(False)
    b = step(True)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    c = step(True, update=True)           # prologue prefetched for the OLD weights, then the weights change
    d = step(False)                       # reference: plain step on the new weights
    assert not torch.equal(a[0], c[0])
    for u, v in zip(c, d):
        assert torch.equal(u, v)


@pytest.mark.parametrize("fam,bmode,Cn", [("arcface", "auto", 30_011), ("arcface", "recompute", 30_011),
                                          ("curricularface", "stash", 30_011), ("cosface", "auto", 160_001)])
def test_phase_graph_replay_is_bit_identical(fam, bmode, Cn, monkeypatch):
    """The CUDA-graph cache of the whole-phase entry points (mh_step_cache_create): a phase whose arguments were seen before
    is replayed with one cudaGraphLaunch.  Same kernels and arguments, so every step must carry the bits of the plain
    launches -- with new DATA at the same addresses every step, and with input tensors that alternate between addresses."""
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import build_head, prime_head
    cfg = mo.HeadConfig.default(fam)
    B = 300
    x0, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=9)
    xs = [x0.cuda(), (x0 * 0.7).cuda().roll(3, 0), (x0 + 0.1).cuda()]

    def trajectory(graph):
        monkeypatch.setenv("MH_STEP_GRAPH", graph)
        monkeypatch.setenv("MH_STEP_API", "1")
        head = prime_head(build_head(pkg, fam, cfg, Cn).cuda(), fam, W, mo.HeadState(), None)
        head.backward_mode = bmode
        xbuf = [torch.empty_like(xs[0]) for _ in range(2)]
        y = labels.cuda()
        outs = []
        for step in range(8):
            xg = xbuf[step % 2]                                  # alternating input addresses: two keys per phase
            xg.requires_grad_(False).copy_(xs[step % 3])         # new data at an address seen before
            xg.requires_grad_(True)
            xg.grad = None
            head._param().grad = None
            out = head.fused_loss(xg, y)
            out.loss.backward()
            # results go to the host: a loop that kept every step's tensors on the device would get fresh addresses for the
            # step's outputs each time (new keys, no replay), which is not what a training loop does
            outs.append((out.loss.detach().cpu(), out.acc1.cpu(), xg.grad.cpu(), head._param().grad.cpu()))
            del out
        torch.cuda.synchronize()
        return outs, head._engine.graph_stats()

    ref, st0 = trajectory("0")
    got, st1 = trajectory("1")
    assert st0 is None
    assert st1[0] >= 6 and 2 <= st1[1] <= 10, st1                 # replays dominate; a handful of captures (address sets)
    from tests.helpers import rel
    for a, b in zip(ref, got):
        for k, (u, v) in enumerate(zip(a, b)):
            if k == 3 and bmode == "recompute":      # backward-G sums the projection terms with fp32 atomics: order-dependent
                assert rel(u, v) < 1e-5
            else:
                assert torch.equal(u, v)
