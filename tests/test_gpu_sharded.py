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
