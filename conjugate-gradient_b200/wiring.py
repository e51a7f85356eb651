"""Rank wiring for one-process-per-GPU launches (torchrun): what MPI_Init_thread /
MPI_COMM_WORLD did for the reference (code/MPI/cg_main.cc:15-20).

`torch.distributed` is used only to ship two small opaque blobs between the ranks -- the NCCL
unique id (ncclAllGather baseline) and the per-rank exchange blobs (CUDA IPC handles of the
gather buffers, for the exchange fused into the mat-vec kernel).  Works with any backend
(nccl on the GPU box, gloo in the CPU tests, where `ctx` is a stand-in).
"""
from __future__ import annotations


def shard_of(n: int, rank: int, world: int):
    """(first_row, rows) of `rank` under partition_matrix (code/MPI/cg.cc:236-268)."""
    n_loc = n // world
    if rank < world - 1:
        return rank * n_loc, n_loc
    return (world - 1) * n_loc, n - (world - 1) * n_loc


def wire(ctx, rank: int, world: int, dist, unique_id_fn, nccl: bool = True, fused: bool = True):
    """Collective over all ranks.  `ctx` needs comm_init / exchange_export / exchange_import
    (conjugate-gradient_b200.Context); `unique_id_fn` is called on rank 0 only.  When both
    exchanges are wired the fused one is left selected (option "exchange" = 1)."""
    if world == 1:
        return
    if nccl:
        box = [unique_id_fn() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) == 0:
            raise RuntimeError("rank %d did not receive the NCCL unique id" % rank)
        ctx.comm_init(bytes(box[0]))
    if fused:
        blobs = [None] * world
        dist.all_gather_object(blobs, ctx.exchange_export())
        if any(b is None for b in blobs):
            raise RuntimeError("rank %d did not receive every exchange blob" % rank)
        ctx.exchange_import([bytes(b) for b in blobs])
