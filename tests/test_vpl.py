"""VPL-ArcFace row: reference criterion.py:619-762, three consecutive steps so that the
memory bank is populated, interpolated and (delta = 2) expires.  CPU: the oracle against the goldens produced from the
reference's own autograd (oracle/make_golden_vpl.py).  GPU: the CUDA path (mh_vpl_mix + the fused tensor-core
pipeline in stash mode) against the same goldens and against the oracle at a BASELINE-like shape."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vpl_oracle as vo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vpl_*.npz")))
LOSS_REL_TC, GRAD_COS_TC, GRAD_NORM_TC = 2e-3, 0.9995, 1e-2


def cfg_of(z):
    return vo.VplConfig(s=float(z["s"]), m=float(z["m"]), easy_margin=bool(z["easy_margin"]), lamda=float(z["lamda"]),
                        delta=int(z["delta"]))


def cosim(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().flatten().cpu(), b.double().flatten().cpu(), dim=0))


def test_goldens_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    cfg, B, Cn, seed, gs = cfg_of(z), int(z["B"]), int(z["C"]), int(z["seed"]), float(z["grad_scale"])
    mem, life = torch.zeros(Cn, 512, dtype=torch.float64), torch.zeros(Cn, dtype=torch.float64)
    for step in range(int(z["n_steps"])):
        x, W, labels = vo.make_inputs(B, Cn, 512, seed * 10 + step)
        r = vo.loss_and_grads(cfg, x, W, labels, mem, life, True, gs)
        mem, life = r["mem"], r["life"]
        assert abs(float(r["loss"]) - float(z[f"s{step}_loss"])) < 1e-10 * abs(float(z[f"s{step}_loss"]))
        assert abs(float(r["acc1"]) - float(z[f"s{step}_acc1"])) < 1e-9 and abs(float(r["acc5"]) - float(z[f"s{step}_acc5"])) < 1e-9
        assert np.allclose(r["dx"].numpy(), z[f"s{step}_dx"], rtol=1e-9, atol=1e-13)
        assert np.allclose(r["dW"].numpy(), z[f"s{step}_dW"], rtol=1e-9, atol=1e-13)
        assert int((life > 0).sum()) == int(z[f"s{step}_n_active"])
        assert abs(float(mem.sum()) - float(z[f"s{step}_mem_sum"])) < 1e-9 * max(1.0, abs(float(z[f"s{step}_mem_sum"])))


def test_training_flag_off_is_plain_arcface_with_eps():
    """norm_training_flag = False: no memory, cosine = cosine_weight (criterion.py:726-727)."""
    cfg = vo.VplConfig()
    x, W, labels = vo.make_inputs(8, 61, 512, 3)
    r = vo.loss_and_grads(cfg, x, W, labels, torch.zeros(61, 512), torch.zeros(61), training_flag=False)
    assert float(r["alpha"].abs().sum()) == 0.0 and float(r["life"].sum()) == 0.0 and torch.isfinite(r["loss"])


def test_module_contract_host():
    import face_recognition_models_b200 as pkg
    h = pkg.VPLArcFace(512, 50, s=64.0, m=0.5, easy_margin=True, lamda=0.15, delta=100)
    assert list(h.state_dict()) == ["weight", "mem", "life", "cos_m", "sin_m", "th", "mm"]      # criterion.py:656-668
    assert tuple(h.weight.shape) == (50, 512) and tuple(h.mem.shape) == (50, 512) and tuple(h.life.shape) == (50,)
    assert h.cos_m.dtype == torch.float32 and h.norm_training_flag is True
    h.change_training_mode(False)
    assert h.norm_training_flag is False
    with pytest.raises(NotImplementedError):
        h(torch.zeros(2, 512), torch.zeros(2, dtype=torch.long))


def run_cuda_steps(pkg, cfg, B, Cn, seed, n_steps, gs):
    head = pkg.VPLArcFace(512, Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin, lamda=cfg.lamda, delta=cfg.delta).cuda()
    out = []
    for step in range(n_steps):
        x, W, labels = vo.make_inputs(B, Cn, 512, seed * 10 + step)
        with torch.no_grad():
            head.weight.copy_(W.cuda())
        head.weight.grad = None
        xg = x.cuda().requires_grad_(True)
        o = head.fused_loss(xg, labels.cuda())
        (o.loss * gs).backward()
        torch.cuda.synchronize()
        out.append(dict(loss=float(o.loss), acc1=float(o.acc1), acc5=float(o.acc5), dx=xg.grad.cpu(), dW=head.weight.grad.cpu(),
                        mem=head.mem.cpu().clone(), life=head.life.cpu().clone(), inputs=(x, W, labels)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_matches_reference_golden(path):
    import face_recognition_models_b200 as pkg
    z = np.load(path)
    cfg, B, Cn, seed, gs = cfg_of(z), int(z["B"]), int(z["C"]), int(z["seed"]), float(z["grad_scale"])
    for step, r in enumerate(run_cuda_steps(pkg, cfg, B, Cn, seed, int(z["n_steps"]), gs)):
        ref_loss = float(z[f"s{step}_loss"])
        assert abs(r["loss"] - ref_loss) <= LOSS_REL_TC * abs(ref_loss)
        assert abs(r["acc1"] - float(z[f"s{step}_acc1"])) < 1e-3 and abs(r["acc5"] - float(z[f"s{step}_acc5"])) < 1e-3
        for got, ref in ((r["dx"], torch.from_numpy(z[f"s{step}_dx"])), (r["dW"], torch.from_numpy(z[f"s{step}_dW"]))):
            assert cosim(got, ref) >= GRAD_COS_TC
            assert abs(float(got.double().norm()) - float(ref.norm())) <= GRAD_NORM_TC * float(ref.norm())
        assert int((r["life"] > 0).sum()) == int(z[f"s{step}_n_active"])
        assert abs(float(r["mem"].double().sum()) - float(z[f"s{step}_mem_sum"])) < 1e-4 * max(1.0, abs(float(z[f"s{step}_mem_sum"])))


@pytest.mark.gpu
def test_cuda_matches_oracle_at_scale():
    """B = 512, C = 10,575 (BASELINE config 2 shape), three steps with a live memory bank; flag-off step afterwards."""
    import face_recognition_models_b200 as pkg
    cfg = vo.VplConfig(easy_margin=False, lamda=0.15, delta=100)
    B, Cn = 512, 10575
    res = run_cuda_steps(pkg, cfg, B, Cn, 4, 3, 1.0)
    mem, life = torch.zeros(Cn, 512, dtype=torch.float64), torch.zeros(Cn, dtype=torch.float64)
    for r in res:
        x, W, labels = r["inputs"]
        ref = vo.loss_and_grads(cfg, x, W, labels, mem, life, True, 1.0)
        mem, life = ref["mem"], ref["life"]
        assert abs(r["loss"] - float(ref["loss"])) <= LOSS_REL_TC * abs(float(ref["loss"]))
        assert abs(r["acc1"] - float(ref["acc1"])) < 0.5 and abs(r["acc5"] - float(ref["acc5"])) < 0.5
        assert cosim(r["dx"], ref["dx"]) >= GRAD_COS_TC and cosim(r["dW"], ref["dW"]) >= GRAD_COS_TC
        assert abs(float(r["dW"].double().norm()) - float(ref["dW"].norm())) <= 2e-3 * float(ref["dW"].norm())
        assert torch.allclose(r["mem"].double(), mem, atol=1e-5) and torch.equal(r["life"].double(), life)


def _sharded_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import face_recognition_models_b200 as pkg
        cfg = vo.VplConfig(easy_margin=False, lamda=0.15, delta=2)        # delta = 2: entries expire within the run
        Bl, Cn = 48, 2001                                                   # ragged shards
        head = pkg.ShardedMarginHead("vpl_arcface", Cn, s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin, lamda=cfg.lamda,
                                     delta=cfg.delta).cuda()
        b, e = head.c_begin, head.c_end
        mem, life = torch.zeros(Cn, 512, dtype=torch.float64), torch.zeros(Cn, dtype=torch.float64)
        for step in range(3):
            x, W, labels = vo.make_inputs(Bl * world, Cn, 512, 70 + step)
            with torch.no_grad():
                head.shard_parameter().copy_(W[b:e].cuda())
            head.shard_parameter().grad = None
            xl = x[rank * Bl:(rank + 1) * Bl].cuda().requires_grad_(True)
            o = head.fused_loss(xl, labels[rank * Bl:(rank + 1) * Bl].cuda())
            o.loss.backward()
            torch.cuda.synchronize()
            ref = vo.loss_and_grads(cfg, x, W, labels, mem, life, True, 1.0)
            mem, life = ref["mem"], ref["life"]
            assert abs(float(o.loss) - float(ref["loss"])) <= LOSS_REL_TC * abs(float(ref["loss"])), (step, float(o.loss), float(ref["loss"]))
            assert abs(float(o.acc1) - float(ref["acc1"])) < 1.1
            assert cosim(xl.grad, ref["dx"][rank * Bl:(rank + 1) * Bl]) >= GRAD_COS_TC
            assert cosim(head.shard_parameter().grad, ref["dW"][b:e]) >= GRAD_COS_TC
            assert torch.allclose(head.local.mem.double().cpu(), mem[b:e], atol=1e-5)
            assert torch.equal(head.local.life.double().cpu(), life[b:e])
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_class_sharded_vpl_matches_oracle():
    """VPL-ArcFace with the class centres, the memory bank and the lifetimes sharded over 2 ranks (NCCL)."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
