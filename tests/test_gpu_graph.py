"""GPU: a fused head step is CUDA-graph capturable (no host sync, no allocation outside torch's graph-safe pool, static
workspaces), which is how the launch-bound small configurations (BASELINE config 2: B=512, C=10,575 - ~15 launches of a
few microseconds each) should be driven."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fam,bmode", [("arcface", "auto"), ("arcface", "recompute"), ("curricularface", "auto")])
def test_fused_step_replays_from_a_cuda_graph(fam, bmode):
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import build_head, prime_head
    B, Cn = 512, 10575
    cfg = mo.HeadConfig.default(fam)
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=21)
    head = prime_head(build_head(pkg, fam, cfg, Cn).cuda(), fam, W, mo.HeadState(), None)
    head.backward_mode = bmode
    xs = x.cuda().requires_grad_(True)
    ys = labels.cuda()

    def step():
        xs.grad = None
        head._param().grad = None
        out = head.fused_loss(xs, ys)
        out.loss.backward()
        return out

    # eager reference + warm-up on a side stream (PyTorch's whole-network capture recipe: the leaves' gradient
    # accumulators must not be tied to the legacy default stream before the capture)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        if fam == "curricularface":
            head.t.zero_()
        e = step()
        side.synchronize()
        loss_e, dx_e, dW_e = float(e.loss.detach()), xs.grad.clone(), head._param().grad.clone()
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    # new inputs of the same shape, copied into the static buffers, then replay
    if fam == "curricularface":
        head.t.zero_()
    with torch.no_grad():
        xs.copy_(x.cuda())
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(out.loss.detach()) - loss_e) <= 1e-6 * abs(loss_e)
    assert torch.equal(xs.grad, dx_e)
    dW_g = head._param().grad
    if head._engine.stash_ok():
        assert torch.equal(dW_g, dW_e)                   # stash mode: no atomics anywhere, bit-reproducible
    else:
        # recompute mode: the projection sums r_j are accumulated with fp32 atomics (order varies): equal to ~1e-6
        assert float((dW_g - dW_e).norm() / dW_e.norm()) < 1e-5

    # launch-bound: replaying the graph must not be slower than the eager step
    def timeit(fn, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_graph, t_eager = timeit(g.replay), timeit(step)
    print(f"{fam}/{bmode}: eager {t_eager * 1e3:.0f} us, graph {t_graph * 1e3:.0f} us per fused step (B=512, C=10,575)")
    assert t_graph <= t_eager * 1.05
