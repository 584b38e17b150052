"""CPU, world_size 2, gloo: the class-sharded head's collective plumbing (ShardComm, shard_range, scale
convention) reproduces the single-process result.  The per-shard arithmetic is emulated with the oracle
here (the product's compute is CUDA-only); what is under test is the N>1 host logic of sharded.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import margin_oracle as mo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fam, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from face_recognition_models_b200.sharded import ShardComm, shard_range
        comm = ShardComm()
        Bl, Cn = 4, 51
        cfg = mo.HeadConfig.default(fam)
        x, W, labels = mo.make_inputs(fam, Bl * world, Cn, 512, seed=42)
        Wc = W if mo.LAYOUT[fam] == "CD" else W.t().contiguous()
        full = mo.loss_and_grads(cfg, mo.HeadState(), x, W, labels)
        ff = mo.forward_logits(cfg, mo.HeadState(), x, W, labels)

        # ---- what each rank owns -------------------------------------------------------------------
        b, e = shard_range(Cn, world, rank)
        x_l, y_l = x[rank * Bl:(rank + 1) * Bl].double(), labels[rank * Bl:(rank + 1) * Bl]
        W_l = Wc[b:e].double()
        # 1. gather the batch
        x_g, y_g = comm.gather_rows(x_l), comm.gather_rows(y_l)
        assert torch.equal(y_g, labels) and torch.allclose(x_g, x.double())
        xn = x_g.norm(dim=1)
        xh = x_g / xn[:, None]
        wh = W_l / W_l.norm(dim=1, keepdim=True)
        raw = xh @ wh.t()                                            # [B_g, C_local]
        # 2. target cosine: owner contributes, all-reduce
        y_loc = y_g - b
        owned = (y_loc >= 0) & (y_loc < e - b)
        t_raw = torch.zeros(Bl * world, dtype=torch.float64)
        t_raw[owned] = raw[owned.nonzero().flatten(), y_loc[owned]]
        comm.allreduce_sum_(t_raw)
        rows = mo.row_terms(cfg, mo.HeadState(), xn, t_raw)
        # 3. local logits + statistics
        bounds = mo._clamp_bounds(fam)
        c = raw if bounds is None else raw.clamp(*bounds)
        inside = torch.ones_like(raw) if bounds is None else ((raw >= bounds[0]) & (raw <= bounds[1])).double()
        kind, ha, hb = rows["hard"]
        thr = rows["thr"][:, None]
        if kind == 2:
            hard = c > thr
            u, du = torch.where(hard, c * (ha + c), c), torch.where(hard, ha + 2 * c, torch.ones_like(c))
        else:
            u, du = c, torch.ones_like(c)
        z = rows["scale"][:, None] * u
        dzdc = rows["scale"][:, None] * du * inside
        idx = owned.nonzero().flatten()
        z[idx, y_loc[owned]] = rows["zt"][owned]
        dzdc[idx, y_loc[owned]] = rows["dzt"][owned]
        m = z.max(dim=1).values
        l = torch.exp(z - m[:, None]).sum(1)
        cnt = (c > rows["t"][:, None]).double()
        cnt[idx, y_loc[owned]] = 0
        stats = torch.stack([m, l, cnt.sum(1), torch.zeros_like(m)])   # [4, B_g]
        # 4. all-gather + merge
        allst = comm.allgather_stats(stats)
        M = allst[:, 0].max(dim=0).values
        Lsum = (allst[:, 1] * torch.exp(allst[:, 0] - M[None])).sum(0)
        lse = M + torch.log(Lsum)
        loss = (lse - rows["zt"]).mean()
        assert abs(float(loss) - float(full["loss_id"])) < 1e-10
        assert torch.equal(allst[:, 2].sum(0).long(), full["rank_count"])
        # 5./6. local backward, reduce-scatter of dx^
        G = torch.exp(z - lse[:, None])
        G[idx, y_loc[owned]] -= 1.0
        dc = G * dzdc / (Bl * world)
        dxh_part = dc @ wh
        mine = comm.reduce_scatter_rows(dxh_part)
        dxh_full = (ff["dz_dc"] * (torch.exp(ff["logits"] - full["lse"][:, None]) -
                                   torch.nn.functional.one_hot(labels, Cn)) / (Bl * world)) @ ff["wh"]
        assert torch.allclose(mine, dxh_full[rank * Bl:(rank + 1) * Bl], atol=1e-12)
        dwh = dc.t() @ xh
        dW_l = (dwh - wh * (wh * dwh).sum(1, keepdim=True)) / W_l.norm(dim=1, keepdim=True)
        dW_ref = full["dW"] if mo.LAYOUT[fam] == "CD" else full["dW"].t()
        assert torch.allclose(dW_l, dW_ref[b:e], atol=1e-12)
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fam", ["arcface", "curricularface", "cosface"])
def test_two_rank_sharded_plumbing(fam):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fam, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_range_partitions_classes():
    from face_recognition_models_b200.sharded import shard_range
    for Cn, R in ((2_000_000, 8), (10_575, 8), (85_742, 4), (10, 3), (7, 8)):
        ranges = [shard_range(Cn, R, r) for r in range(R)]
        assert ranges[0][0] == 0 and ranges[-1][1] == Cn
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0]
        assert sum(e - b for b, e in ranges) == Cn


def _ckpt_worker(rank, world, port, fam, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import face_recognition_models_b200 as pkg
        Cn = 37                                                  # ragged: shards of 19 and 18 classes
        kw = dict(arcface=dict(s=64.0, m=0.5, easy_margin=False), curricularface=dict(m=0.5, s=64.0, momentum=0.01))[fam]
        torch.manual_seed(123)                                   # same "reference checkpoint" on every rank
        ref = pkg.HEAD_CLASSES[fam](512, Cn, **kw)
        sd_ref = {k: v.clone() for k, v in ref.state_dict().items()}
        if fam == "curricularface":
            sd_ref["t"].fill_(0.37)
        head = pkg.ShardedMarginHead(fam, Cn, **kw)
        head.load_full_state_dict(sd_ref)
        b, e = head.c_begin, head.c_end
        full = sd_ref[head.local.param_name]
        mine = full[b:e] if head.local.layout == "CD" else full[:, b:e]
        assert torch.equal(head.shard_parameter().data, mine)
        back = head.full_state_dict()
        assert set(back) == set(sd_ref)
        for k in sd_ref:
            assert torch.equal(back[k], sd_ref[k]), k
        ref2 = pkg.HEAD_CLASSES[fam](512, Cn, **kw)
        ref2.load_state_dict(back)                               # the gathered dict loads into the unsharded head
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fam", ["arcface", "curricularface"])
def test_checkpoint_interchange_with_unsharded_head(fam):
    """a reference-style state_dict scatters into the class shards and gathers back unchanged."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ckpt_worker, args=(r, world, port, fam, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
