"""GPU, 2 and 8 ranks, NCCL: the class-sharded head equals the oracle on the concatenated batch (loss, accuracy, dx after
the reduce-scatter, the shard's dW), in both backward modes; HeadSGD on the shards; bad labels; the device guard.
Run on a multi-GPU box: pytest -m gpu tests/test_gpu_sharded.py (a case is skipped when the box has fewer GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fam, bmode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        from oracle import margin_oracle as mo
        from tests.helpers import build_head, cosim, prime_head, rel
        Bl, Cn = 96, 5003                       # ragged last shard at 2 and at 8 ranks (2502/2501, 7 x 626 + 621)
        cfg = mo.HeadConfig.default(fam)
        # "stash" on CurricularFace / SphereFace = the guarded stash (every rank derives the same flag from the merged
        # statistics); "stash_adv": inputs built to underflow, so the gated general path must take over on every rank
        adversarial = bmode == "stash_adv"
        if adversarial:
            from tests.test_gpu_guarded_stash import _inputs
            x, W, labels = _inputs(fam, Bl * world, Cn, 77, True)
            bmode = "stash"
        else:
            x, W, labels = mo.make_inputs(fam, Bl * world, Cn, 512, seed=77)
        ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels)
        kw = dict(arcface=dict(s=cfg.s, m=cfg.m, easy_margin=False), curricularface=dict(m=cfg.m, s=cfg.s, momentum=cfg.momentum),
                  cosface=dict(s=cfg.s, m=cfg.m), mv_am=dict(margin=cfg.m, mv_weight=cfg.mv_weight, s=cfg.s),
                  sphereface=dict(m=cfg.sphere_m))[fam]
        head = pkg.ShardedMarginHead(fam, Cn, **kw).cuda()
        head.engine.backward_mode = bmode
        b, e = head.c_begin, head.c_end
        Wc = W if mo.LAYOUT[fam] == "CD" else W.t()
        shard = Wc[b:e] if mo.LAYOUT[fam] == "CD" else Wc[b:e].t()
        with torch.no_grad():
            head.shard_parameter().copy_(shard.contiguous().cuda())
        xl = x[rank * Bl:(rank + 1) * Bl].cuda().requires_grad_(True)
        yl = labels[rank * Bl:(rank + 1) * Bl].cuda()
        out = head.fused_loss(xl, yl)
        out.loss.backward()
        torch.cuda.synchronize()
        assert abs(float(out.loss) - float(ref["loss"])) < 2e-3 * abs(float(ref["loss"]))
        assert abs(float(out.acc1) - float(ref["acc1"])) < 0.6
        if bmode == "stash" and fam in ("curricularface", "sphereface"):
            assert int(head.engine._ws["stash_guard"].item()) == (1 if adversarial else 0)
        dx_ref = ref["dx"][rank * Bl:(rank + 1) * Bl]
        dW_ref = ref["dW"][b:e] if mo.LAYOUT[fam] == "CD" else ref["dW"][:, b:e]
        assert cosim(xl.grad, dx_ref) > 0.9995 and cosim(head.shard_parameter().grad, dW_ref) > 0.9995
        assert rel(xl.grad, dx_ref) < 1e-2 and rel(head.shard_parameter().grad, dW_ref) < 1e-2
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _bad_label_worker(rank, world, port, q):
    """A label outside [0, C) must poison the loss on EVERY rank (no silent drop of the target, no hang)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        head = pkg.ShardedMarginHead("arcface", 1000, s=64.0, m=0.5, easy_margin=False).cuda()
        g = torch.Generator(device="cuda").manual_seed(5 + rank)
        x = torch.randn(8, 512, device="cuda", generator=g)
        y = torch.randint(0, 1000, (8,), device="cuda", generator=g)
        ok = head.fused_loss(x, y)
        assert bool(torch.isfinite(ok.loss))
        if rank == world - 1:
            y[3] = 1000                          # one bad label on one rank
        bad = head.fused_loss(x, y)
        assert bool(torch.isnan(bad.loss)), float(bad.loss)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_bad_label_poisons_every_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bad_label_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_head_on_non_current_device():
    """a head on cuda:1 must work while cuda:0 is the current device (streams, workspaces and launches follow
    the data), like the reference's modules."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import cosim
    torch.cuda.set_device(0)
    x, W, labels = mo.make_inputs("cosface", 64, 1000, 512, seed=3)
    ref = mo.loss_and_grads(mo.HeadConfig.default("cosface"), mo.HeadState(), x, W, labels)
    head = pkg.CosFace(512, 1000, s=64.0, m=0.35).to("cuda:1")       # DC layout: > 48 KB dynamic smem prologue on device 1
    with torch.no_grad():
        head.kernel.copy_(W.to("cuda:1"))
    xg = x.to("cuda:1").requires_grad_(True)
    out = head.fused_loss(xg, labels.to("cuda:1"))
    out.loss.backward()
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert abs(float(out.loss) - float(ref["loss"])) < 2e-3 * abs(float(ref["loss"]))
    assert cosim(xg.grad, ref["dx"]) > 0.9995 and cosim(head.kernel.grad, ref["dW"]) > 0.9995
    with pytest.raises(pkg.MarginHeadError, match="share one device"):
        head.fused_loss(xg, labels.to("cuda:0"))


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("fam,bmode", [("arcface", "auto"), ("arcface", "recompute"), ("curricularface", "auto"),
                                       ("cosface", "auto"), ("mv_am", "auto"), ("curricularface", "stash"),
                                       ("curricularface", "stash_adv"), ("sphereface", "stash")])
def test_sharded_matches_oracle(fam, bmode, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fam, bmode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _worker_sgd(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        Bl, Cn = 64, 3001                       # ragged shards
        torch.manual_seed(11)
        heads = [pkg.ShardedMarginHead("arcface", Cn, s=64.0, m=0.5, easy_margin=False).cuda() for _ in range(2)]
        with torch.no_grad():
            heads[1].shard_parameter().copy_(heads[0].shard_parameter())
        opts = [torch.optim.SGD([heads[0].shard_parameter()], lr=0.05, momentum=0.9, weight_decay=5e-4, foreach=False),
                pkg.HeadSGD([heads[1]], lr=0.05, momentum=0.9, weight_decay=5e-4)]
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        for step in range(3):
            x = torch.randn(Bl, 512, device="cuda", generator=g)
            y = torch.randint(0, Cn, (Bl,), device="cuda", generator=g)
            losses = []
            for head, opt in zip(heads, opts):
                opt.zero_grad(set_to_none=True)
                out = head.fused_loss(x, y)
                out.loss.backward()
                opt.step()
                losses.append(float(out.loss))
            assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[0]), (step, losses)
            a, b = heads[0].shard_parameter().detach(), heads[1].shard_parameter().detach()
            assert (a - b).abs().max().item() <= 2.0 ** -23 * a.abs().max().item(), step
        assert heads[1].engine._shadow is not None
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_head_sgd_equals_torch_sgd():
    """HeadSGD on the class shards (each rank steps its own [C/R, 512] slice) follows torch.optim.SGD bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sgd, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
