"""Multi-GPU bootstrap: one process per GPU, the libsharp-MPI layout of
commander3/src/comm_map_mod.f90:197-261 (m's and ring pairs round-robin over ranks).

In Commander the group is an MPI communicator (comm_chain, commander3/src/comm_param_mod.f90:322);
here the launcher is torchrun, torch.distributed carries the 128-byte NCCL id from rank 0 to the
other ranks, and the library opens its own NCCL communicator for the phase all-to-all and the CG
dot-product all-reduces (include/cmdr_sht.h part 2).
"""
from __future__ import annotations

import ctypes as C
import os

from . import sharp


class Comm:
    """What comm_mapinfo needs from a communicator: rank, size and the integer handle that
    sharp_execute_mpi_fortran receives as `comm` (an MPI_Fint in the reference)."""

    def __init__(self, rank: int, size: int, handle):
        self.rank, self.size, self.handle = rank, size, handle

    def allreduce_sum_(self, t):
        """In-place sum over ranks of a float64 CUDA tensor (mpi_dot_product's MPI_Allreduce,
        commander3/src/comm_utils.f90:599-614)."""
        if self.size > 1:
            import torch
            sharp.lib().cmdr_sht_allreduce_sum(self.handle, t.data_ptr(), t.numel(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return t

    def set_exchange(self, mode: int) -> None:
        """Collective: 1 = phase exchange fused into the Legendre kernels (peer stores over NVLink), 0 = NCCL
        all-to-all, -1 = automatic (cmdr_sht_comm_set_exchange)."""
        if self.size > 1:
            sharp.lib().cmdr_sht_comm_set_exchange(self.handle, int(mode))


_next_handle = [1]


def init_from_torch(device=None) -> Comm:
    """Collective over the default torch.distributed group (must be initialised, any backend)."""
    import torch
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()
    handle = _next_handle[0]
    _next_handle[0] += 1
    if size == 1:
        return Comm(0, 1, None)
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(device)
    L = sharp.lib()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        L.cmdr_sht_get_unique_id(buf)
    idt = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        idt = idt.to(device)
    dist.broadcast(idt, src=0)
    raw = bytes(idt.cpu().tolist())
    rc = L.cmdr_sht_comm_register(handle, rank, size, (C.c_ubyte * 128).from_buffer_copy(raw))
    if rc != 0:
        raise RuntimeError("cmdr_sht_comm_register failed")
    return Comm(rank, size, handle)


def destroy(comm: Comm) -> None:
    if comm.handle is not None:
        sharp.lib().cmdr_sht_comm_destroy(comm.handle)
