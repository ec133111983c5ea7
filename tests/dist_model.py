"""TEST INFRASTRUCTURE -- a CPU restatement of the protocol of bwtm_merge_distributed (csrc/bwtm_dist.cu) over
torch.distributed (gloo): the same split of B's sequences, the same splitter rule, the same piece layout of the exchange
(one-shot and batched: S sorted runs per rank, G S pieces per receiver), the same slice boundaries. The per-rank work
(walk, sort, interleave) is done by the oracle and numpy instead of the CUDA kernels, so what this checks is the
ARITHMETIC OF THE PROTOCOL: that the slices the ranks produce tile the merged BWT exactly."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):          # bwtm_shard_range
    first = (total * rank) // world
    return first, (total * (rank + 1)) // world - first


def all_reduce_sum(values):
    t = torch.tensor(np.asarray(values, dtype=np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def all_gather(values):
    t = torch.tensor(np.asarray(values, dtype=np.int64))
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def distributed_merge_model(oracle, A, B, batches=1):
    """Returns (begin, end, symbols of this rank's slice of the merged sequence)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n_a, n_b, m_b = A.size, B.size, B.sequences
    a_seq, b_seq = A.decode(), B.decode()

    # 1. search + sort of this rank's sequences, in `batches` sorted runs
    seq_first, seq_count = shard_range(m_b, rank, world)
    runs = []
    for k in range(batches):
        first = seq_first + (seq_count * k) // batches; last = seq_first + (seq_count * (k + 1)) // batches
        keys = oracle.build_ra_walk(A, B, first, last - 1) if last > first else np.zeros(0, dtype=np.uint64)
        runs.append(np.sort(keys.astype(np.int64)))

    def below(x):
        return sum(int(np.searchsorted(run, x, side="left")) for run in runs)

    # 2. splitters: smallest p with p + #{keys < p} >= k (n_a + n_b) / world (the counts are summed over the ranks)
    splitter = [0]
    for k in range(1, world):
        target = ((n_a + n_b) * k) // world
        lo, hi = 0, n_a + 1
        while lo < hi:        # every rank runs the same bisection: the all-reduce keeps them in step
            mid = (lo + hi) // 2
            if mid + int(all_reduce_sum([below(mid)])[0]) < target:
                lo = mid + 1
            else:
                hi = mid
        splitter.append(max(lo, splitter[-1]))
    splitter.append(n_a + 1)

    # 3. pieces[dst][k]: the part of run k that goes to rank dst; everybody learns all of them
    bound = [[int(np.searchsorted(run, splitter[d], side="left")) for run in runs] for d in range(world)] + [[len(run) for run in runs]]
    my_pieces = [[bound[d + 1][k] - bound[d][k] for k in range(batches)] for d in range(world)]
    pieces = all_gather(np.asarray(my_pieces).reshape(-1))          # pieces[src][dst * batches + k]
    count = lambda src, dst, k: int(pieces[src][dst * batches + k])
    total = sum(count(s, d, k) for s in range(world) for d in range(world) for k in range(batches))
    assert total == n_b
    b_lo = sum(count(s, d, k) for s in range(world) for d in range(rank) for k in range(batches))

    # 4. the exchange: every piece lands in the receiver's buffer ordered by (source rank, run)
    send = [np.concatenate([runs[k][bound[d][k]:bound[d + 1][k]] for k in range(batches)]) if batches else np.zeros(0, np.int64) for d in range(world)]
    sizes = [int(sum(count(src, rank, k) for k in range(batches))) for src in range(world)]
    received = [torch.zeros(sizes[src], dtype=torch.int64) for src in range(world)]
    requests = []
    for peer in range(world):
        if peer == rank:
            received[rank] = torch.tensor(send[rank])
            continue
        requests.append(dist.isend(torch.tensor(send[peer]), peer))
        requests.append(dist.irecv(received[peer], peer))
    for request in requests:
        request.wait()
    # 5. G S sorted pieces -> the keys of the slice (the device merges them pairwise, or range by range)
    arrived = [received[src].numpy() for src in range(world)]
    slice_keys = np.sort(np.concatenate(arrived)) if arrived else np.zeros(0, np.int64)
    for src in range(world):          # every piece is sorted and lies inside this rank's range of A positions
        offset = 0
        for k in range(batches):
            piece = arrived[src][offset:offset + count(src, rank, k)]; offset += count(src, rank, k)
            assert np.all(piece[:-1] <= piece[1:]) and (len(piece) == 0 or (piece[0] >= splitter[rank] and piece[-1] < splitter[rank + 1]))

    # 6. the slice: A positions [a_lo, a_hi) and B positions [b_lo, b_lo + recv) interleaved by "RA[j] symbols of A come
    #    before B[j]" (bwt.cpp:215-282): B[j] sits at merged position j + RA[j]
    a_lo, a_hi = min(splitter[rank], n_a), min(splitter[rank + 1], n_a)
    begin, end = a_lo + b_lo, a_hi + b_lo + len(slice_keys)
    merged = np.zeros(end - begin, dtype=np.uint8)
    from_b = np.zeros(end - begin, dtype=bool)
    positions = np.arange(len(slice_keys), dtype=np.int64) + b_lo + slice_keys - begin
    from_b[positions] = True
    merged[positions] = b_seq[b_lo:b_lo + len(slice_keys)]
    merged[~from_b] = a_seq[a_lo:a_hi]
    return begin, end, merged
