"""One process per GPU: bootstrap of the NCCL communicator the kernel library uses for its per-iteration exchange.

torch.distributed is plumbing only (rendezvous, shipping the 128-byte NCCL unique id, barriers, max-over-ranks of
timings); the data path -- (cost, index, position) of one particle per rank per ParticleSwarm pass or sequential round, a
histogram or n doubles per CrossEntropy stage -- is moved by the kernels themselves over the peer-memory mailboxes set up
in eqs_ctx_comm_init (csrc/comm.cu, csrc/common.cuh: ll_*), or by the library's own NCCL communicator where peer mapping is
unavailable.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

from .global_optimisers import Context, unique_id


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def exchange_unique_id(rank: int, make_id=unique_id) -> bytes:
    """Rank 0 creates the id, everyone receives it (works on any initialised torch.distributed backend)."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    assert isinstance(uid, (bytes, bytearray)) and len(uid) == 128
    return bytes(uid)


def init_context(device: Optional[int] = None) -> Context:
    """Context for this rank; joins the library communicator when WORLD_SIZE > 1 (torch.distributed must be up)."""
    rank, world, local = env_rank_world()
    ctx = Context(local if device is None else device)
    if world > 1:
        ctx.comm_init(rank, world, exchange_unique_id(rank))
    return ctx
