"""Config 3 of BASELINE.json on N GPUs (torchrun): two-input merge of 2 x R reads, A and B replicated on every
GPU, B's sequences split across GPUs, RA values exchanged by A-position range.

The inputs are too large for the sort-based fixture builder in one piece, so every collection is built in
chunks that are merged with the product itself (single-GPU merges, outside the timed region). The result of
the distributed merge is then compared byte for byte, on rank 0, with an independent route to the same BWT:
inserting B's chunks into A one after the other with single-GPU merges.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bwt-merge_b200"))

import bwtm_b200                                    # noqa: E402
from bwtm_b200 import FMI, MergeParameters, synth   # noqa: E402


def build_collection(args, seed, chunks, thr, log):
    per = args.reads // chunks
    index = None
    for c in range(chunks):
        n = per if c < chunks - 1 else args.reads - per * (chunks - 1)
        part = FMI.synthetic(args.genome, 42, args.read_len, thr, [(seed, n, c * per)])
        index = part if index is None else FMI.merge(index, part)
        log("  seed %d chunk %d/%d: %d symbols so far" % (seed, c + 1, chunks, index.size()))
    return index


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=500_000_000)
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--patterns", type=int, default=100_000)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--sequence-blocks", type=int, default=0, help="search batches (needed where two full key buffers do not fit: 2 GPUs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); bwtm_b200.set_device(local)
    if world > 1:
        from datetime import timedelta
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timedelta(minutes=30))
    comm = bwtm_b200.Communicator.from_torch(dist, rank, world) if world > 1 else None
    thr = synth.error_threshold(0.01)

    def log(msg):
        if rank == 0:
            print("[%7.1f s] %s" % (time.time() - t0, msg), flush=True)
    t0 = time.time()
    log("building A and B (%d reads x %d bp each, %d chunks) on every rank" % (args.reads, args.read_len, args.chunks))
    A = build_collection(args, 1, args.chunks, thr, log)
    B = build_collection(args, 2, args.chunks, thr, log)
    n_a, n_b = A.size(), B.size()
    log("A: %d symbols, %d RLE bytes; B: %d symbols, %d RLE bytes" % (n_a, A.bytes(), n_b, B.bytes()))
    pats = [p for p in synth.patterns(synth.genome(min(args.genome, 50_000_000), 42), args.patterns, 32, 99)] if args.genome <= 50_000_000 else None
    if pats is None:   # patterns from the first 50 Mbp of the genome (host generator: the prefix is the same function of the index)
        pats = [p for p in synth.patterns(synth.genome(50_000_000, 42), args.patterns, 32, 99)]
    pre = A.count(pats) + B.count(pats)

    params = MergeParameters(); params.sequence_blocks = args.sequence_blocks
    times = []
    M = None
    for step in range(args.steps + 1):
        if M is not None:
            M.close()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t = time.perf_counter()
        M = comm.merge(A, B, params, keep_inputs=True) if comm is not None else FMI.merge(A, B, params, keep_inputs=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t)
        t = M.timings.as_dict()
        log("merge %d: %.1f ms  %s  batches %d, walk records of %d bytes" % (
            step, times[-1] * 1e3, {k: round(v * 1e3, 1) for k, v in t.items() if k.endswith("_seconds")}, t["search_batches"], t["walk_record_bytes"]))
    best = min(times[1:]) if len(times) > 1 else times[0]
    post = M.count(pats)
    ok_counts = bool(np.array_equal(pre, post))
    stage = {k[:-8]: round(v * 1e3, 1) for k, v in M.timings.as_dict().items() if k.endswith("_seconds")}
    result = {"config": "two-input merge 2x%dx%dbp reads, %d bp genome, %d GPUs" % (args.reads, args.read_len, args.genome, world),
              "stages_ms_last_merge": stage, "search_batches": int(M.timings.search_batches), "device_memory_peak_bytes": int(bwtm_b200.memory_stats()[1]),
              "inserted_bases": int(n_b), "merged_symbols": int(n_a + n_b), "merged_rle_bytes": int(M.bytes()),
              "merge_ms": best * 1e3, "merged_bases_per_second": n_b / best, "pattern_counts_match": ok_counts,
              "pattern_occurrences": int(post.sum())}

    if not args.no_verify and rank == 0:
        log("verification route: inserting B's chunks into A with single-GPU merges")
        per = args.reads // args.chunks
        V = None
        for c in range(args.chunks):
            n = per if c < args.chunks - 1 else args.reads - per * (args.chunks - 1)
            part = FMI.synthetic(args.genome, 42, args.read_len, thr, [(2, n, c * per)])
            V = FMI.merge(A if V is None else V, part, keep_inputs=(V is None))
            log("  inserted chunk %d/%d: %d symbols" % (c + 1, args.chunks, V.size()))
        same = (V.bytes() == M.bytes())
        if same:
            got, want = M.rle(), V.rle()
            same = bool(np.array_equal(got, want))
        result["byte_identical_to_sequential_route"] = same
        log("byte-identical: %s" % same)
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(result), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
