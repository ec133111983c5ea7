"""Host-side logic of the multi-GPU path on CPU: two gloo ranks agree on the split of B's sequences and on
the NCCL id exchange; creating a communicator without a device fails loudly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bwtm_b200


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


def _worker(rank, world, port, totals, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bwtm_b200.build_library(); bwtm_b200.lib()
        # every rank's block of sequence ids, gathered everywhere
        mine = torch.tensor([list(bwtm_b200.shard_range(t, rank, world)) for t in totals], dtype=torch.int64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        # the NCCL id travels from rank 0 through the process group
        ident = torch.zeros(bwtm_b200.COMM_ID_BYTES, dtype=torch.uint8)
        id_ok = True
        if rank == 0:
            try:
                ident = torch.from_numpy(bwtm_b200.comm_unique_id().copy())
            except bwtm_b200.BwtmError:
                id_ok = False   # no NCCL library on this host
        dist.broadcast(ident, 0)
        error = None
        if not torch.cuda.is_available():
            try:
                bwtm_b200.Communicator(ident.numpy(), rank, world)
            except bwtm_b200.BwtmError as e:
                error = e.code
        queue.put((rank, [g.tolist() for g in gathered], ident.numpy().tobytes(), id_ok, error))
    finally:
        dist.destroy_process_group()


def test_two_ranks_agree_on_shards_and_id():
    world, totals = 2, [0, 1, 2, 7, 100000, 10_000_001]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, totals, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(queue.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, shards0, id0, ok0, err0), (_, shards1, id1, ok1, err1) = results
    assert shards0 == shards1 and id0 == id1
    for k, total in enumerate(totals):      # blocks are contiguous, disjoint and cover [0, total)
        position = 0
        for r in range(world):
            first, count = shards0[r][k]
            assert first == position
            position += count
        assert position == total
    if ok0:
        assert any(b != 0 for b in id0)
    if not torch.cuda.is_available():
        assert err0 in (-2, -7) and err1 in (-2, -7)     # BWTM_ERR_CUDA (no device) or BWTM_ERR_COMM


def test_shard_range_matches_reference_blocks():
    """One block per rank: same boundaries as floor(total * k / world)."""
    bwtm_b200.build_library()
    for total in (1, 5, 64, 1000, 123457):
        for world in (1, 2, 3, 8):
            blocks = [bwtm_b200.shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == total
            assert all(f == (total * r) // world for r, (f, _) in enumerate(blocks))


def _model_worker(rank, world, port, batches, queue):
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
        for p in (root, os.path.join(root, "bwt-merge_b200"), here):
            if p not in sys.path:
                sys.path.insert(0, p)
        from oracle.oracle import Oracle
        from conftest import make_collection
        from dist_model import distributed_merge_model
        orc = Oracle()
        _, bwt_a = make_collection(orc, 3000, 260, 45, 0.02, 42, 1, 0.01)
        _, bwt_b = make_collection(orc, 3000, 150, 45, 0.02, 42, 2, 0.01)
        A, B = orc.from_comps(bwt_a), orc.from_comps(bwt_b)
        begin, end, symbols = distributed_merge_model(orc, A, B, batches)
        want = orc.merge(A, B).decode()
        queue.put((rank, begin, end, bool(np.array_equal(symbols, want[begin:end])), len(want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batches", [(2, 1), (2, 3), (3, 2)])
def test_protocol_of_the_distributed_merge_on_gloo(world, batches):
    """tests/dist_model.py restates the protocol of bwtm_merge_distributed (shards, splitter rule, piece layout of the
    one-shot and the batched exchange, slice boundaries) over gloo with the oracle doing the per-rank work: the slices of
    the ranks must tile the merged BWT of the oracle exactly."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, batches, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(queue.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    position = 0
    for rank, begin, end, same, total in results:
        assert begin == position and same, (rank, begin, end, same)
        position = end
    assert position == results[0][4]
