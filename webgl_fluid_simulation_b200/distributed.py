"""One-process-per-GPU launcher glue: torch.distributed is only the plumbing (rendezvous +
broadcast of the ncclUniqueId); the halo exchange itself runs inside libfluid_b200 over NCCL."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .sim import FluidSimulation


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = _lib.lib().fluid_nccl_unique_id(buf, 128)
    if rc != 0:
        raise _lib.FluidError(rc, (_lib.lib().fluid_last_error(None) or b"").decode())
    return buf.raw


def create_slab_simulation(config=None, canvas_width=1024, canvas_height=1024, device=-1, **kw) -> FluidSimulation:
    """Every rank of an initialised torch.distributed group calls this; returns the rank's slab."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return FluidSimulation(config, canvas_width, canvas_height, device=device, rank=rank, world=world,
                           nccl_uid=box[0], **kw)
