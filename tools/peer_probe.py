"""Latency of the two small exchanges of a case-sharded step, NCCL vs the peer-memory kernels (csrc/peer.cu), measured between
a barrier-aligned start and the end of the exchange on every rank (max over ranks).  Run under torchrun."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from madeleine_b200 import parallel  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
px = parallel.PeerExchange.get(dev)
res = {"world": world, "peer": parallel.PeerExchange.status()}
warm = torch.zeros(1 << 22, device=dev)
for name, numel in (("allgather_64KB", 16 * 2 * 512), ("allreduce_2MB", 524288 + 2048)):
    x = torch.randn(numel, device=dev)
    out_nccl = torch.empty(world * numel, device=dev)

    def nccl():
        if name.startswith("allgather"):
            dist.all_gather_into_tensor(out_nccl, x)
        else:
            dist.all_reduce(x)

    def peer():
        if name.startswith("allgather"):
            px.all_gather(x)
        else:
            px.all_reduce_(x)

    big = torch.randn(64 << 20, device=dev)

    def busy():                                  # ~0.3 ms of device work queued first: the host is ahead of the device when the
        big.mul_(1.0000001)                      # exchange is launched, as in a real step

    for label, fn in (("none", lambda: None), ("nccl", nccl), ("peer", peer)):
        if label == "peer" and px is None:
            continue
        times = []
        for it in range(30):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            busy()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= 5:
                times.append(e0.elapsed_time(e1) * 1e3)
        t = torch.tensor([sorted(times)[len(times) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"{name}_{label}_us"] = round(float(t), 1)
    del big
if px is not None:      # correctness of both exchanges
    x = torch.arange(4096, device=dev, dtype=torch.float32) + 1000 * rank
    g = px.all_gather(x)
    ref = torch.cat([torch.arange(4096, device=dev, dtype=torch.float32) + 1000 * r for r in range(world)])
    assert torch.equal(g, ref)
    y = torch.full((8192,), float(rank + 1), device=dev)
    px.all_reduce_(y)
    assert torch.equal(y, torch.full_like(y, world * (world + 1) / 2))
    res["checked"] = True
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
