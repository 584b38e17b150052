"""GPU: Partial-FC negative sampling (insightface partial_fc.sample is the published algorithm, the
reference has no line for it).  The sampled step must equal the margin oracle evaluated on the sampled sub-matrix
W[index] with the labels remapped, and classes that were not sampled must receive exactly zero gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fam,bmode", [("arcface", "auto"), ("arcface", "recompute"), ("cosface", "auto"), ("curricularface", "auto")])
def test_single_gpu_sampled_step_matches_oracle_on_submatrix(fam, bmode):
    import face_recognition_models_b200 as pkg
    from oracle import margin_oracle as mo
    from tests.helpers import cosim, rel
    torch.manual_seed(3)
    B, Cn = 96, 5003
    cfg = mo.HeadConfig.default(fam)
    x, W, labels = mo.make_inputs(fam, B, Cn, 512, seed=31)
    kw = dict(arcface=dict(s=cfg.s, m=cfg.m, easy_margin=False), cosface=dict(s=cfg.s, m=cfg.m),
              curricularface=dict(m=cfg.m, s=cfg.s, momentum=cfg.momentum))[fam]
    head = pkg.ShardedMarginHead(fam, Cn, sample_rate=0.2, **kw).cuda()
    head.engine.backward_mode = bmode
    cd = mo.LAYOUT[fam] == "CD"
    with torch.no_grad():
        head.shard_parameter().copy_(W.cuda())
    xg = x.cuda().requires_grad_(True)
    out = head.fused_loss(xg, labels.cuda())
    out.loss.backward()
    torch.cuda.synchronize()
    index = head.last_index.cpu()
    assert index.numel() == head.num_sample == 1000
    y_sub = torch.searchsorted(index, labels)
    assert torch.equal(index[y_sub], labels)
    W_sub = W[index] if cd else W[:, index]
    ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W_sub, y_sub)
    assert abs(float(out.loss) - float(ref["loss"])) < 2e-3 * abs(float(ref["loss"]))
    assert abs(float(out.acc1) - float(ref["acc1"])) < 0.6
    gW = head.shard_parameter().grad.cpu()
    gW_sub = gW[index] if cd else gW[:, index]
    assert cosim(xg.grad, ref["dx"]) > 0.9995 and cosim(gW_sub, ref["dW"]) > 0.9995
    assert rel(xg.grad, ref["dx"]) < 1e-2 and rel(gW_sub, ref["dW"]) < 1e-2
    mask = torch.ones(Cn, dtype=torch.bool)
    mask[index] = False
    rest = gW[mask] if cd else gW[:, mask]
    assert float(rest.abs().max()) == 0.0                          # unsampled classes: no gradient at all


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        from oracle import margin_oracle as mo
        from tests.helpers import cosim
        torch.manual_seed(100 + rank)
        Bl, Cn = 64, 4001                       # ragged shards 2001 / 2000
        cfg = mo.HeadConfig.default("arcface")
        x, W, labels = mo.make_inputs("arcface", Bl * world, Cn, 512, seed=41)
        head = pkg.ShardedMarginHead("arcface", Cn, sample_rate=0.25, s=cfg.s, m=cfg.m, easy_margin=False).cuda()
        b, e = head.c_begin, head.c_end
        with torch.no_grad():
            head.shard_parameter().copy_(W[b:e].cuda())
        xl = x[rank * Bl:(rank + 1) * Bl].cuda().requires_grad_(True)
        out = head.fused_loss(xl, labels[rank * Bl:(rank + 1) * Bl].cuda())
        out.loss.backward()
        torch.cuda.synchronize()
        k = head.num_sample
        idx_all = torch.empty(world * k, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(idx_all, head.last_index + b)          # global class ids of every rank's sample
        idx_all = idx_all.cpu()
        y_sub = torch.searchsorted(idx_all, labels)
        assert torch.equal(idx_all[y_sub], labels)                           # every positive was sampled by its owner
        ref = mo.loss_and_grads(cfg, mo.HeadState(), x, W[idx_all], y_sub)
        assert abs(float(out.loss) - float(ref["loss"])) < 2e-3 * abs(float(ref["loss"]))
        assert cosim(xl.grad, ref["dx"][rank * Bl:(rank + 1) * Bl]) > 0.9995
        mine = head.last_index.cpu()
        assert cosim(head.shard_parameter().grad.cpu()[mine], ref["dW"][rank * k:(rank + 1) * k]) > 0.9995
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sampled_step_matches_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
