"""GPU, 2 ranks, NCCL: the class-sharded head equals the single-GPU head on the concatenated batch.
Run on a multi-GPU box: pytest -m gpu tests/test_gpu_sharded.py (skipped when fewer than 2 GPUs)."""
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


def _worker(rank, world, port, fam, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        from oracle import margin_oracle as mo
        from tests.helpers import build_head, cosim, prime_head, rel
        Bl, Cn = 96, 5000
        cfg = mo.HeadConfig.default(fam)
        x, W, labels = mo.make_inputs(fam, Bl * world, Cn, 512, seed=77)
        ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels)
        kw = dict(arcface=dict(s=cfg.s, m=cfg.m, easy_margin=False), curricularface=dict(m=cfg.m, s=cfg.s, momentum=cfg.momentum),
                  cosface=dict(s=cfg.s, m=cfg.m))[fam]
        head = pkg.ShardedMarginHead(fam, Cn, **kw).cuda()
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


@pytest.mark.parametrize("fam", ["arcface", "curricularface", "cosface"])
def test_two_gpu_sharded_matches_oracle(fam):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fam, q)) for r in range(world)]
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
