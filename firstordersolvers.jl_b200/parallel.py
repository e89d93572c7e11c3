"""Multi-GPU plumbing: one process per GPU (``torchrun``), ``torch.distributed`` only for the
rendezvous.  The data path is inside the CUDA library: A is row-sharded, every rank runs the
fused mat-vec on its rows and the A' partial sums (+ the owners' A x rows) are all-reduced
with NCCL over NVLink (new functionality, the reference has no parallelism).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SHARD_ALIGN = 16  # row_begin must be a multiple of the 16-row tile (fos_load_conic_dense)


def row_shard(m: int, rank: int, nranks: int, align: int = SHARD_ALIGN):
    """Rows [begin, begin+count) owned by `rank`: contiguous, multiple-of-`align` boundaries,
    balanced to within one `align` block; ranks past the end get count = 0."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    blocks = (m + align - 1) // align
    base, extra = divmod(blocks, nranks)
    b0 = rank * base + min(rank, extra)
    nb = base + (1 if rank < extra else 0)
    begin = min(b0 * align, m)
    end = min((b0 + nb) * align, m)
    return begin, end - begin


def batch_shard(nproblems: int, rank: int, nranks: int):
    """Independent problems [begin, begin+count) of a batch owned by `rank` (no collective)."""
    base, extra = divmod(nproblems, nranks)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def exchange_comm_id(rank: int, make_id, dist=None):
    """Rank 0 creates the 128-byte NCCL unique id (``make_id()``), everyone receives it through
    ``torch.distributed`` (any backend; gloo on CPU in the tests)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    buf = torch.zeros(_lib.FOS_COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.from_numpy(np.frombuffer(make_id(), dtype=np.uint8).copy()))
    if dist.get_backend() == "nccl":  # NCCL process groups only move device tensors
        dbuf = buf.cuda()
        dist.broadcast(dbuf, src=0)
        buf = dbuf.cpu()
    else:
        dist.broadcast(buf, src=0)
    return bytes(buf.numpy().tobytes())


def nccl_unique_id() -> bytes:
    L = _lib.load()
    arr = (C.c_uint8 * _lib.FOS_COMM_ID_BYTES)()
    rc = L.fos_comm_unique_id(arr)
    if rc != 0:
        raise _lib.FosError(rc, (L.fos_last_error(None) or b"").decode())
    return bytes(arr)


def init_comm(handle, rank: int, nranks: int, comm_id: bytes = None):
    """fos_comm_init on a ``Handle`` (before loading the problem).  ``comm_id=None`` passes the all-zero id: no NCCL
    communicator, the peer-memory exchange (``enable_p2p_exchange``) must follow the load."""
    if comm_id is None:
        comm_id = bytes(_lib.FOS_COMM_ID_BYTES)
    arr = (C.c_uint8 * _lib.FOS_COMM_ID_BYTES).from_buffer_copy(comm_id)
    handle.ck(handle.L.fos_comm_init(handle.h, rank, nranks, arr))


def enable_p2p_exchange(handle, rank: int, nranks: int, dist=None) -> bool:
    """After ``fos_load_conic_dense`` on every rank: export this rank's CUDA-IPC exchange slot,
    all-gather the 64-byte handles through ``torch.distributed`` (plumbing only) and import the
    table -- from then on every pass over A uses the fused peer-memory exchange kernel instead of
    fold + ncclAllReduce.

    Returns True when EVERY rank mapped every peer.  If any rank could not (CUDA IPC disabled in the
    container, no peer access), all ranks switch back to the NCCL exchange together and False is returned:
    the ranks must agree, a mixed configuration would dead-lock."""
    import torch
    if dist is None:
        import torch.distributed as dist
    on_gpu = dist.get_backend() == "nccl"
    ok = 1
    mine = torch.zeros(_lib.FOS_IPC_HANDLE_BYTES, dtype=torch.uint8)
    try:
        arr = (C.c_uint8 * _lib.FOS_IPC_HANDLE_BYTES)()
        handle.ck(handle.L.fos_comm_p2p_export(handle.h, arr))
        mine = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
    except _lib.FosError:
        ok = 0
    if on_gpu:
        mine = mine.cuda()
    table = [torch.empty_like(mine) for _ in range(nranks)]
    dist.all_gather(table, mine)
    flag = torch.tensor([ok], dtype=torch.int32, device=mine.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        try:
            flat = np.concatenate([t.cpu().numpy() for t in table]).astype(np.uint8)
            buf = (C.c_uint8 * (nranks * _lib.FOS_IPC_HANDLE_BYTES)).from_buffer_copy(flat.tobytes())
            handle.ck(handle.L.fos_comm_p2p_import(handle.h, buf))
        except _lib.FosError:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=mine.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    all_ok = int(flag.item()) == 1
    if not all_ok and ok == 1:
        handle.set_option("exchange_impl", 0)  # this rank mapped its peers but another one did not
    dist.barrier()
    return all_ok
