"""Per-step cost of the class-sharded head's collectives, timed alone on the compute stream (torchrun, NCCL):
all-gather x (fp32 [B,512]) + labels, all-reduce of the target cosines, all-gather of the row statistics,
reduce-scatter of dx^ (fp32 [R*B,512]).  Compare with the step time bench.py prints for the same N.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/time_collectives.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_recognition_models_b200 import _lib as L  # noqa: E402
from face_recognition_models_b200.sharded import ShardComm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
saved = os.dup(1)
os.dup2(2, 1)                                    # NCCL's banner goes to stderr, the JSON line to stdout
dist.init_process_group("nccl", device_id=dev)
comm = ShardComm()
B = int(os.environ.get("B", 1024))
Bg = B * world
B_pad = (Bg + 255) // 256 * 256
x = torch.randn(B, 512, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
t_raw = torch.randn(Bg, device=dev)
stats = torch.randn(L.ST_PLANES, B_pad, device=dev)
all_stats = torch.empty(world, L.ST_PLANES, B_pad, device=dev)
dxh = torch.randn(Bg, 512, device=dev)


def step():
    comm.gather_rows(x)
    comm.gather_rows(y)
    comm.allreduce_sum_(t_raw)
    comm.allgather_stats(stats, out=all_stats)
    comm.reduce_scatter_rows(dxh)


for _ in range(10):
    step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    os.dup2(saved, 1)
    print(json.dumps({"n_gpus": world, "B_per_gpu": B, "collectives_ms_per_step": round(float(ms), 4),
                      "bytes": {"allgather_x": Bg * 512 * 4, "allgather_labels": Bg * 8, "allreduce_t": Bg * 4,
                                "allgather_stats": world * L.ST_PLANES * B_pad * 4, "reduce_scatter_dx": Bg * 512 * 4}}), flush=True)
dist.destroy_process_group()
