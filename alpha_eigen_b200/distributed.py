"""Multi-GPU set-up: one process per GPU, sweep rows (angle x group) sharded by angle, the flux update
sharded by cells: one NCCL reduce-scatter of the partial scalar flux and one all-gather of the new source
per source iteration."""
from __future__ import annotations

import os

_comm = None


def init(backend: str = "nccl"):
    """Join the torchrun rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), bind this process
    to its GPU and create the NCCL communicator the C ABI uses.  No-op for a single process."""
    global _comm
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return None
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    from .device import Comm
    _comm = Comm(rank, world)
    return _comm


def communicator():
    return _comm


def shutdown():
    global _comm
    import torch.distributed as dist
    if _comm is not None:
        _comm.close()
        _comm = None
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
