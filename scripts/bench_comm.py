"""Latency of the peer-memory all-reduce at the arena sizes of the step, under torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N scripts/bench_comm.py
20 collectives captured in one CUDA graph, replayed between events; per-call microseconds, max over ranks; NCCL's
all_reduce on the same buffers beside it.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import recnn_b200
from recnn_b200.dist import PeerComm

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
comm = PeerComm(None, dev, 430000)
out = {"world": world}
for n in (4, 429828):
    x = torch.randn(n, device=dev)
    for _ in range(3):
        comm.all_reduce(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            comm.all_reduce(x)
    ts = []
    for _ in range(10):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) * 1e3 / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    ts.sort()
    out["peer_us_n%d" % n] = ts[len(ts) // 2]
    tn = []
    for _ in range(10):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(x)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) * 1e3 / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tn.append(float(t.item()))
    tn.sort()
    out["nccl_us_n%d" % n] = tn[len(tn) // 2]
dist.barrier()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
