"""Launched by torchrun on N GPUs (tests/test_gpu_dist.py): distributed merge == single-GPU merge == oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bwt-merge_b200"))

import bwtm_b200                                    # noqa: E402
from bwtm_b200 import FMI, MergeParameters, synth   # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); bwtm_b200.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = bwtm_b200.Communicator.from_torch(dist, rank, world)
    thr = synth.error_threshold(0.01)
    cases = [(200000, 20000, 100, 0), (200000, 20000, 100, 524288), (200000, 16000, 100, 262144), (50000, 3000, 60, 8192), (3000, 5, 40, 4096), (100, 1, 30, 0)]
    # Three passes: peer windows sized by the first (largest) case; windows growing from the smallest case up
    # in a fresh communicator; the NCCL send/recv + broadcast route.
    runs = [(comm, cases), (None, cases[::-1]), ("nccl", cases[:1] + cases[3:])]
    for which, (use, todo) in enumerate(runs):
        if use is None:
            dist.barrier(); comm.close(); comm = bwtm_b200.Communicator.from_torch(dist, rank, world)
        elif use == "nccl":
            dist.barrier(); comm.close(); os.environ["BWTM_NCCL_EXCHANGE"] = "1"
            comm = bwtm_b200.Communicator.from_torch(dist, rank, world)
        check_cases(comm, rank, thr, todo)
    os.environ.pop("BWTM_NCCL_EXCHANGE", None)
    dist.barrier(); comm.close(); comm = bwtm_b200.Communicator.from_torch(dist, rank, world)
    check_cases(comm, rank, thr, cases[:4], sequence_blocks=3)      # the search in batches, pieces merged range by range
    check_cases(comm, rank, thr, cases[3:], sequence_blocks=7)
    check_failures(comm, rank, world, thr)
    dist.barrier()
    if rank == 0:
        print("dist_check ok: %d ranks, %d cases (+ batched search, + injected failures)" % (world, sum(len(todo) for _, todo in runs)))
    comm.close()
    dist.destroy_process_group()


def check_failures(comm, rank, world, thr):
    """A rank that fails locally must not leave the others waiting: every rank gets an error from the same merge, and the
    communicator is still usable afterwards (bwtm_dist.cu: agree_on_status, poisoned writer state)."""
    A = FMI.synthetic(50000, 42, 60, thr, [(1, 3000)]); B = FMI.synthetic(50000, 42, 60, thr, [(2, 1500)])
    want = FMI.merge(A, B, keep_inputs=True).rle()
    for phase in ("search", "interleave", "writer"):
        for failing in sorted({0, world - 1, world // 2}):
            os.environ["BWTM_INJECT_FAILURE"] = "%s:%d" % (phase, failing)
            try:
                comm.merge(A, B, keep_inputs=True)
                raise AssertionError("rank %d: the merge succeeded although rank %d failed in '%s'" % (rank, failing, phase))
            except bwtm_b200.BwtmError as error:
                assert ("injected" in str(error)) == (rank == failing) or "failed" in str(error), str(error)
            finally:
                os.environ.pop("BWTM_INJECT_FAILURE", None)
            assert np.array_equal(comm.merge(A, B, keep_inputs=True).rle(), want), "rank %d: merge after a failed merge differs" % rank


def check_cases(comm, rank, thr, cases, sequence_blocks=1):
    for G, n, L, slab in cases:
        A = FMI.synthetic(G, 42, L, thr, [(1, n)]); B = FMI.synthetic(G, 42, L, thr, [(2, max(1, n // 2))])
        p = MergeParameters(); p.slab_symbols = slab; p.sequence_blocks = sequence_blocks
        single = FMI.merge(A, B, p, keep_inputs=True)
        multi = comm.merge(A, B, p, keep_inputs=True)
        want = single.rle()
        got = multi.rle()
        assert np.array_equal(got, want), "rank %d: distributed merge differs (case %s)" % (rank, (G, n, L, slab))
        assert np.array_equal(multi.counts(), single.counts()) and multi.hash() == single.hash()
        # sequential: the distributed result is the next A on every rank
        C_ = FMI.synthetic(G, 42, L, thr, [(3, max(1, n // 3))])
        again = comm.merge(multi, C_, p, keep_inputs=True)
        direct = FMI.synthetic(G, 42, L, thr, [(1, n), (2, max(1, n // 2)), (3, max(1, n // 3))])
        assert np.array_equal(again.rle(), direct.rle()), "rank %d: sequential distributed merge differs" % rank


if __name__ == "__main__":
    main()
