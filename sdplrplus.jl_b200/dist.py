"""Host-side multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the
rendezvous, the NCCL id broadcast and the max-over-ranks timing; the data path itself runs
inside libsdplrp_b200.so (its own NCCL communicator on the handle's stream).

The reference has no distributed layer; this mirrors how a Julia host would
bootstrap one handle per process."""
from __future__ import annotations

import os

import numpy as np


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (127.0.0.1 rendezvous)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload, nbytes, src=0):
    """Broadcast a fixed-size byte string (the 128-byte NCCL unique id) from `src`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


ROW_COST = 22.0  # streaming bytes per vertex in units of one gathered nonzero (csrc/comm.cu, comm_partition)


def balanced_row_blocks(rowptr, world):
    """Contiguous row blocks with ~equal (nnz + ROW_COST * rows) weight -- the same rule as comm_partition()
    in csrc/comm.cu (kept in Python for tests and for sizing host buffers)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = rowptr.size - 1
    w = rowptr[:-1] + ROW_COST * np.arange(n, dtype=np.float64)
    total = float(rowptr[-1] + ROW_COST * n)
    starts = [0]
    for p in range(1, world):
        starts.append(int(np.searchsorted(w, total * p / world, side="left")))
    starts.append(n)
    return np.asarray(starts, dtype=np.int64)


def _equal_block_params(n, world):
    B = (n + world - 1) // world
    last = n - (world - 1) * B
    return B, last, n % world


def dealt_row_starts(n, world):
    """Row blocks of the multi-GPU layout of relabelled (skewed) patterns -- the same rule as k_deal_rows() in
    csrc/preprocess.cu: the degree-sorted vertices are dealt round-robin to the ranks; ranks 0..world-2 hold exactly
    B = ceil(n / world) rows (so one in-place ncclAllGather moves a factor), the last rank the remainder."""
    B, last, _ = _equal_block_params(n, world)
    if last >= 1:
        return np.asarray([min(n, q * B) for q in range(world)] + [n], dtype=np.int64)
    starts = [0]
    for q in range(world):
        starts.append(starts[-1] + (n - q + world - 1) // world)
    return np.asarray(starts, dtype=np.int64)


def deal_order(sorted_vertices, world):
    """Internal order produced by the deal: position k of the degree-sorted list goes to rank k % world, slot k // world;
    the ranks this leaves one row short of B take the last (lowest-degree) rows of the last rank."""
    sorted_vertices = np.asarray(sorted_vertices)
    n = sorted_vertices.size
    B, last, rem = _equal_block_params(n, world)
    starts = dealt_row_starts(n, world)
    k = np.arange(n, dtype=np.int64)
    p, slot = k % world, k // world
    if last >= 1 and rem != 0:
        move = (p == world - 1) & (slot >= last)
        p = np.where(move, rem + (slot - last), p)
        slot = np.where(move, B - 1, slot)
    out = np.empty_like(sorted_vertices)
    out[starts[p] + slot] = sorted_vertices
    return out


def make_handle(Handle):
    """Create this rank's handle; rank 0 makes the NCCL id and every rank receives it."""
    rank, world, local = init_process_group()
    if world == 1:
        return Handle(device=local)
    nid = Handle.nccl_unique_id() if rank == 0 else bytes(128)
    nid = broadcast_bytes(nid, 128, src=0)
    return Handle(device=local, rank=rank, world=world, nccl_id=nid)
