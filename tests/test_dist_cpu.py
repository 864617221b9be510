"""Host-side multi-rank logic on CPU: world_size-2 gloo process group (the GPU data path runs NCCL inside the
library; what is testable without GPUs is the rendezvous, the id broadcast, the max-over-ranks timing and the
row-partition rules that every rank must evaluate identically)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from sdplrplus.jl_b200 import dist as spdist
    r, w, l = spdist.init_process_group(backend="gloo")
    assert (r, w, l) == (rank, world, rank)
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = spdist.broadcast_bytes(payload, 128, src=0)           # the 128-byte NCCL unique id travels this way
    assert got == bytes(range(128))
    mx = spdist.max_over_ranks(10.0 + rank)                     # timing rule: max over ranks
    assert mx == 10.0 + world - 1
    spdist.barrier()
    # every rank derives the same partition from the same row pointer
    rng = np.random.default_rng(5)
    deg = np.sort(rng.zipf(2.2, size=5000).clip(max=800))[::-1]
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    blocks = spdist.balanced_row_blocks(rowptr, world)
    dealt = spdist.dealt_row_starts(5000, world)
    results[rank] = (blocks.tolist(), dealt.tolist())
    dist.destroy_process_group()


def test_gloo_world2_rendezvous_and_partition():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert results[0] == results[1]
        blocks, dealt = results[0]
    assert blocks[0] == 0 and blocks[-1] == 5000 and all(a <= b for a, b in zip(blocks, blocks[1:]))
    assert dealt[0] == 0 and dealt[-1] == 5000


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dealt_layout_balances_rows_and_nonzeros(world):
    """Round-robin deal of the degree-sorted vertices (csrc/preprocess.cu, k_deal_rows): every rank's block gets
    the same number of rows (+-1) and the same number of nonzeros (+- one maximum degree)."""
    from sdplrplus.jl_b200 import dist as spdist
    rng = np.random.default_rng(1)
    n = 10007
    deg = rng.zipf(2.3, size=n).clip(max=2000)
    order = np.argsort(-deg, kind="stable")            # hub-first
    internal = spdist.deal_order(order, world)          # internal label -> vertex
    assert sorted(internal.tolist()) == list(range(n))  # a permutation
    starts = spdist.dealt_row_starts(n, world)
    rows = np.diff(starts)
    assert rows.max() - rows.min() <= world - 1 and rows.sum() == n  # ranks 0..P-2 hold exactly ceil(n/P) rows
    nnz = np.array([deg[internal[starts[q]:starts[q + 1]]].sum() for q in range(world)])
    assert nnz.max() - nnz.min() <= deg.max()
    for q in range(world):                              # each block is itself hub-first
        d = deg[internal[starts[q]:starts[q + 1]]]
        assert np.all(d[:-1] >= d[1:])


def test_balanced_row_blocks_weights():
    from sdplrplus.jl_b200 import dist as spdist
    rowptr = np.arange(0, 1001 * 5, 5)                  # 1000 rows of 5 nonzeros
    b = spdist.balanced_row_blocks(rowptr, 4)
    assert b.tolist() == [0, 250, 500, 750, 1000]


def _gen_worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import sdplrplus.jl_b200 as sp
    from sdplrplus.jl_b200 import dist as spdist
    import bench
    spdist.init_process_group(backend="gloo")
    torch.manual_seed(100 + rank)          # the ranks' own generators disagree: only rank 0's graph may be used
    asm, b, normC, E, _ = bench.generate(sp, 3000, 20000, 42)
    results[rank] = (asm.I.tolist(), asm.J.tolist(), asm.V.tolist(), asm.mat_off.tolist(), asm.gids.tolist(), b.tolist(), normC, E)
    dist.destroy_process_group()


def test_every_rank_gets_rank0s_problem():
    """bench.generate under world 2: rank 0 draws the graph, the triplets are broadcast (a per-device generator would give
    every rank its own graph, which is what the full-size 2-GPU run of round 1 caught)."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_gen_worker, args=(world, port, results), nprocs=world, join=True)
        r0, r1 = results[0], results[1]
    assert r0 == r1
    I, J, V, off, gids, b, normC, E = r0
    assert len(I) == len(J) == len(V) == off[-1] and off[-2] == 3000 and gids[-1] == 3001 and len(b) == 3000 and E > 15000
