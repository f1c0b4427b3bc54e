"""Data-parallel host logic on CPU (gloo, world_size 2): row sharding, replica sync and the
gradient all-reduce convention (local losses pre-scaled by 1/N_global, SUM over ranks).
The device kernels are not involved -- this covers recnn_b200/dist.py and the contract the
engine's split-phase path relies on; the N>1 numerics on GPUs are covered by
tests/test_gpu_parity.py::test_two_rank_equals_one_rank (needs 2 GPUs) and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import recnn_b200
from recnn_b200.dist import shard_rows
from oracle import recnn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_rows_partitions_exactly():
    for n in (1, 7, 4096, 8192, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) replicas become bit-identical to rank 0
        torch.manual_seed(100 + rank)
        nets = {"policy_net": recnn_b200.nn.Actor(68, 16, 32), "value_net": recnn_b200.nn.Critic(68, 16, 32)}
        recnn_b200.dist.broadcast_nets(nets)
        digest = torch.cat([p.detach().flatten() for n in sorted(nets) for p in nets[n].parameters()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        same = all(torch.equal(gathered[0], g) for g in gathered)

        # (2) sharded critic gradient: each rank differentiates its rows with the loss scaled by
        #     1/N_global; SUM all-reduce must equal the full-batch gradient (numpy oracle math)
        rng = np.random.default_rng(5)
        p = O.make_critic(rng, 68, 16, 32, 54e-2)
        n = 48
        s = rng.standard_normal((n, 68), dtype=np.float32)
        a = rng.standard_normal((n, 16), dtype=np.float32)
        y = rng.standard_normal((n, 1), dtype=np.float32)
        masks = O.synth_masks(rng, 2, n, 32)

        def grads(lo, hi):
            q, cache = O.critic_forward(p, s[lo:hi], a[lo:hi], [m[lo:hi] for m in masks])
            d_q = (2.0 * (q - y[lo:hi]) / n).astype(np.float32)          # 1/N_global, not 1/N_local
            g, _ = O._mlp_backward(p, cache, d_q, need_dx=False)
            return np.concatenate([g[k].reshape(-1) for k in O.PARAM_ORDER])

        lo, hi = shard_rows(n, rank, world)
        local = torch.from_numpy(grads(lo, hi))
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        full = grads(0, n)
        err = float(np.abs(local.numpy() - full).max() / np.abs(full).max())

        # (3) enable_data_parallel wires the group into (future) engines
        recnn_b200.dist.enable_data_parallel(dict(nets, target_policy_net=nets["policy_net"]), sync_weights=False)
        wired = nets["policy_net"].__dict__["_recnn_dp"][1] == world
        q.put((rank, same, err, wired))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, err, wired in results:
        assert same, "replicas differ after broadcast"
        assert err < 1e-5, err
        assert wired
