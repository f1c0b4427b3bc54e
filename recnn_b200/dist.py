"""Data-parallel sharding of the update step across the GPUs of one node.

The path shards by sample row: every rank holds a replica of the
embedding table, the nets and the optimizer state, processes its own rows, and
the ranks exchange exactly one thing -- the summed weight gradients (critic every
step, actor on policy steps) plus the loss scalars -- with an all-reduce before
the (identical) optimizer step.  Local loss terms are already scaled by
1/N_global on the device, so SUM over ranks gives the single-device gradient.

One process per GPU (torchrun); torch.distributed supplies the NCCL communicator.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .nn.arena import param_arena


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous, balanced row range of this rank (first ``n_rows % world`` ranks get one more)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_nets(nets: dict, group=None, src: int = 0):
    """Make every replica bit-identical to rank ``src`` (one broadcast per net arena)."""
    for name in sorted(nets):
        dist.broadcast(param_arena(nets[name]), src=src, group=group)


def enable_data_parallel(agent_or_nets, group=None, sync_weights=True):
    """Turn on gradient all-reduce for an Algo (or a nets dict).  Must be called on every
    rank after the nets are on their CUDA device."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    nets = agent_or_nets.nets if hasattr(agent_or_nets, "nets") else agent_or_nets
    world = dist.get_world_size(group)
    if sync_weights:
        broadcast_nets(nets, group)
    policy = nets["policy_net"]
    policy.__dict__["_recnn_dp"] = (group, world)
    for eng in policy.__dict__.get("_recnn_engines", {}).values():
        eng.group, eng.world = group, world
        eng.graphs.clear()
    return agent_or_nets
