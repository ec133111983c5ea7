#!/usr/bin/env python
"""bench.py -- merged bases/s of the rank-array path (BASELINE.json metric) on B200.

  python bench.py --gpus 1 --steps 3 --warmup 3          # our arm, config 2 of BASELINE.json
  python bench.py --impl reference                        # the unmodified reference binary on the host cores
  python bench.py --config 5                              # another named configuration (1, 2 or 5)

A step is one merge of the synthetic collections A and B (B inserted into A):
  * `value`  : |B| / device time of bwtm_merge with both indexes resident in HBM (CUDA events);
  * `e2e`    : the same through the C ABI with HOST buffers: upload of both run-length BWTs from pinned
               memory + rank-structure build (K0) + merge + download of the merged run-length BWT;
  * `roofline`: the rank/LF walk kernel (K1), algorithmic bytes 168 B per inserted base;
  * `verified`: after the warm-up merges the merged run-length bytes are compared (SHA-256, byte count, counts)
               with the result of the UNMODIFIED reference binary on the same inputs, recorded in
               tests/golden/reference_merge_digests.json (written by `--impl reference --record-digest`);
  * `cpu_baseline`: the unmodified reference (oracle/_ref/bwt_merge -t nproc) on a bounded sample.

Reference arm (`--impl reference`): the process never loads the CUDA library. bin/bwtm_fixture (a child process)
writes the inputs as native files, then the unmodified reference binary oracle/_ref/bwt_merge merges the FULL
workload once with -t <host cores>; its own timers are the result. A prefix sample is merged first as the warm-up
and its extrapolation is reported beside the measured number.
"""
import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "bwt-merge_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

ALGORITHMIC_BYTES_PER_BASE = 168.0   # LF_B 80 B + rank_A 80 B + 8 B RA value
L2_BYTES = 126 << 20
FIXTURE_TOOL = os.path.join(ROOT, "bwt-merge_b200", "bin", "bwtm_fixture")
REF_MERGE = os.path.join(ROOT, "oracle", "_ref", "bwt_merge")
DIGESTS = os.path.join(ROOT, "tests", "golden", "reference_merge_digests.json")

# The named configurations of BASELINE.json that are two-input merges.
CONFIGS = {
    1: dict(genome=1_000_000, reads=100_000, read_len=100, error=0.01, genome_seed=42),
    2: dict(genome=50_000_000, reads=10_000_000, read_len=150, error=0.01, genome_seed=42),
    5: dict(genome=1_000_000_000, reads=20_000_000, read_len=250, error=0.05, genome_seed=43),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="named configuration of BASELINE.json")
    ap.add_argument("--genome", type=int, default=None)
    ap.add_argument("--reads", type=int, default=None)
    ap.add_argument("--read-len", type=int, default=None)
    ap.add_argument("--error", type=float, default=None)
    ap.add_argument("--genome-seed", type=int, default=None)
    ap.add_argument("--seed-a", type=int, default=1)
    ap.add_argument("--seed-b", type=int, default=2)
    ap.add_argument("--chunk-reads", type=int, default=0, help="build the inputs in chunks of this many reads (0 = automatic)")
    ap.add_argument("--sequence-blocks", type=int, default=0, help="bwtm_merge_options.sequence_blocks (0 = automatic)")
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000, help="reads of B in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--verify-direct", action="store_true", help="also compare with BWT(A ++ B) built directly by the fixture builder")
    ap.add_argument("--record-digest", action="store_true", help="reference arm: store the digest of its result in tests/golden/")
    ap.add_argument("--digest-out", default=None, help="reference arm: also write the digest entry to this file")
    ap.add_argument("--reference-prefix-only", action="store_true", help="reference arm: only the bounded sample (no full run)")
    ap.add_argument("--walk", default="auto", choices=["auto", "single", "pairs"],
                    help="rank/LF walk: 64-byte records, one step per read (single) or 128-byte pair records, two steps per read (pairs)")
    ap.add_argument("--gather-bench", action="store_true", help="measure the random-access HBM peak as well")
    args = ap.parse_args()
    for key, value in CONFIGS[args.config].items():
        if getattr(args, key) is None:
            setattr(args, key, value)
    return args


def workload_name(args):
    return "two-input merge 2x%dx%dbp reads, synthetic %d bp genome, %g%% substitutions" % (
        args.reads, args.read_len, args.genome, 100 * args.error)


def workload_key(args):
    return "%s; seeds %d/%d/%d" % (workload_name(args), args.genome_seed, args.seed_a, args.seed_b)


def config_dict(args, n_a, n_b, rle_bytes):
    """Identical in both arms: a function of the workload alone."""
    records = (n_a + n_b) // 2     # the device rank records take at least 0.5 B per symbol
    return {"workload": workload_name(args), "seeds": [args.genome_seed, args.seed_a, args.seed_b],
            "inserted_bases": int(n_b), "merged_symbols": int(n_a + n_b), "rle_bytes": [int(x) for x in rle_bytes],
            "l2": ("rank records of both inputs (>= %.2f GB) exceed the 126 MB L2; no flush between steps" % (records / 1e9)
                   if records >= 2 * L2_BYTES else "inputs fit in L2: a 252 MB buffer is written between timed steps")}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_digests():
    try:
        return json.load(open(DIGESTS))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device; self.proc = None; self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); out["sm_max_mhz"] = float(f[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# The reference binary and its input files. Nothing here imports the CUDA library or the oracle into
# this process: bin/bwtm_fixture and oracle/_ref/bwt_merge are child processes.

def scratch_dir():
    return tempfile.mkdtemp(prefix="bwtm_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def write_fixture(args, path, segments, device=0):
    """BWT of the given read segments as a native file, built on the GPU by the child process bin/bwtm_fixture."""
    cmd = [FIXTURE_TOOL, "--genome", str(args.genome), "--genome-seed", str(args.genome_seed), "--read-len", str(args.read_len),
           "--error", repr(args.error), "--device", str(device), "--format", "native", "--output", path]
    for seg in segments:
        cmd += ["--segment", ":".join(str(x) for x in seg)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("bwtm_fixture failed: " + res.stderr[-500:])
    return json.loads(res.stdout.strip().splitlines()[-1])


def run_reference_binary(path_a, path_b, path_out, cores, temp_dir):
    """oracle/_ref/bwt_merge -t <cores> -d <temp_dir> A B out (native files). Returns the reference's own timers."""
    t0 = time.perf_counter()
    res = subprocess.run([REF_MERGE, "-t", str(cores), "-d", temp_dir, path_a, path_b, path_out], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        raise RuntimeError("reference bwt_merge failed: " + res.stderr[-500:])

    def timer(text, pattern):
        m = re.search(pattern, text)
        return float(m.group(1)) if m else None
    return {"merge_seconds": timer(res.stdout, r"BWTs merged in ([0-9.e+-]+) seconds \("),      # bwt_merge.cpp:296-297
            "ra_seconds": timer(res.stderr, r"RA built in ([0-9.e+-]+) seconds"),               # fmi.cpp:362
            "interleave_seconds": timer(res.stderr, r"bwt_merge: BWTs merged in ([0-9.e+-]+) seconds"),   # bwt.cpp:302
            "samples_seconds": timer(res.stderr, r"rank/select built in ([0-9.e+-]+) seconds"),  # bwt.cpp:312
            "process_seconds": wall}


def native_file_digest(path):
    """Header fields and SHA-256 of the run-length bytes of a native file (NativeHeader 24 bytes, formats.cpp:488-499;
    then BlockArray::serialize: u64 byte count + whole 8 MiB blocks, support.cpp:296-309)."""
    with open(path, "rb") as f:
        header = np.frombuffer(f.read(24), dtype=np.uint64)
        size = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
        h = hashlib.sha256(); left = size
        while left > 0:
            chunk = f.read(min(left, 64 << 20))
            if not chunk:
                raise RuntimeError("truncated native file " + path)
            h.update(chunk); left -= len(chunk)
    return {"sequences": int(header[1]), "bases": int(header[2]), "rle_bytes": size, "sha256": h.hexdigest()}


def reference_sample(args, tmp, path_a, info_a, cores):
    """Bounded sample: A = the full collection, B = the first --cpu-sample-reads reads of collection B, merged by the
    reference binary; its timers scaled to the full workload (RA by |B|/|B_sample|, interleave + samples by merged length)."""
    sample_reads = min(args.cpu_sample_reads, args.reads)
    path_bs = os.path.join(tmp, "B_sample.native")
    info_bs = write_fixture(args, path_bs, [(args.seed_b, sample_reads)])
    t = run_reference_binary(path_a, path_bs, os.path.join(tmp, "sample_out.native"), cores, tmp)
    os.unlink(os.path.join(tmp, "sample_out.native")); os.unlink(path_bs)
    n_a = info_a["bases"]; n_b = args.reads * (args.read_len + 1); n_bs = info_bs["bases"]
    ra = t["ra_seconds"] if t["ra_seconds"] is not None else t["merge_seconds"]
    rest = (t["interleave_seconds"] or 0.0) + (t["samples_seconds"] or 0.0)
    scale_b = n_b / n_bs; scale_m = (n_a + n_b) / (n_a + n_bs)
    estimate = ra * scale_b + rest * scale_m
    desc = ("unmodified reference binary (oracle/_ref/bwt_merge, SDSL stand-in) -t %d, defaults; A = full %d-symbol collection, "
            "B = first %d reads of collection B (%d symbols): RA %.2f s, interleave %.2f s, samples %.2f s, merge %.2f s; RA time "
            "scaled by |B|/|B_sample| = %.1f and interleave+samples by merged length %.2f"
            % (cores, n_a, sample_reads, n_bs, ra, t["interleave_seconds"] or 0.0, t["samples_seconds"] or 0.0,
               t["merge_seconds"] or 0.0, scale_b, scale_m))
    return {"value": n_b / estimate, "seconds_estimated": estimate, "sample_timers": t, "sample": desc}


def reference_arm(args):
    if not (os.path.exists(REF_MERGE) and os.path.exists(FIXTURE_TOOL)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bwt_merge or bin/bwtm_fixture is not built on this box"}))
        return 0
    cores = os.cpu_count() or 1
    n_a = n_b = args.reads * (args.read_len + 1)
    tmp = scratch_dir()
    try:
        t0 = time.perf_counter()
        path_a, path_b, path_out = (os.path.join(tmp, name) for name in ("A.native", "B.native", "merged.native"))
        info_a = write_fixture(args, path_a, [(args.seed_a, args.reads)])
        info_b = write_fixture(args, path_b, [(args.seed_b, args.reads)])
        build_seconds = time.perf_counter() - t0
        sample = reference_sample(args, tmp, path_a, info_a, cores)      # also the warm-up (page cache, A's file)
        if args.reference_prefix_only:
            timers, seconds, digest, verified = None, sample["seconds_estimated"], None, None
            rle_m = 0
        else:
            timers = run_reference_binary(path_a, path_b, path_out, cores, tmp)
            seconds = timers["merge_seconds"]
            digest = native_file_digest(path_out)
            rle_m = digest["rle_bytes"]
            entry = dict(digest, source="oracle/_ref/bwt_merge -t %d (unmodified reference, SDSL stand-in), inputs from bin/bwtm_fixture" % cores)
            known = load_digests().get(workload_key(args))
            verified = {"against": "digest committed in tests/golden/reference_merge_digests.json" if known else "nothing committed yet",
                        "ok": (known is None or known["sha256"] == digest["sha256"]), "sha256": digest["sha256"]}
            if args.record_digest:
                all_digests = load_digests(); all_digests[workload_key(args)] = entry
                json.dump(all_digests, open(DIGESTS, "w"), indent=1, sort_keys=True)
            if args.digest_out:
                json.dump({workload_key(args): entry}, open(args.digest_out, "w"), indent=1, sort_keys=True)
        value = n_b / seconds
        stage = None
        if timers is not None:
            sort_note = "included in search (run/thread/merge buffers work inside buildRA)"
            stage = {"search": (timers["ra_seconds"] or 0.0) * 1e3, "sort": sort_note,
                     "interleave": (timers["interleave_seconds"] or 0.0) * 1e3, "index": (timers["samples_seconds"] or 0.0) * 1e3}
        line = {
            "impl": "reference", "metric": "merged_bases_per_second", "value": value, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": seconds * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": config_dict(args, n_a, n_b, [info_a["rle_bytes"], info_b["rle_bytes"], rle_m]),
            "input_build_seconds": build_seconds,
            "note": ("one merge of the FULL workload by the unmodified reference binary (FMI::FMI(a, b, parameters) as timed by "
                     "bwt_merge.cpp:290-297), after one warm-up merge of a prefix sample; the run is not repeated --steps times "
                     "because one merge takes minutes" if timers is not None else "prefix sample only (--reference-prefix-only)"),
            "stages_ms": stage, "reference_timers": timers, "prefix_extrapolation": sample, "verified": verified,
            "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "reference",
                             "sample": ("the full workload, once: A = %d symbols, B = %d symbols, -t %d, default buffers, temporary files in %s"
                                        % (n_a, n_b, cores, os.path.dirname(tmp)) if timers is not None else sample["sample"])},
            "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------

def build_collection(FMI, args, thr, seed, reads, chunk_reads):
    """BWT of `reads` reads of one seed on the device; collections beyond the builder's memory are built in chunks that
    are merged with the product itself (outside every timed region)."""
    if chunk_reads <= 0 or chunk_reads >= reads:
        return FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(seed, reads)])
    index, first = None, 0
    while first < reads:
        n = min(chunk_reads, reads - first)
        part = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(seed, n, first)])
        index = part if index is None else FMI.merge(index, part)
        first += n
    return index


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args) if rank == 0 else 0

    import torch
    import bwtm_b200
    from bwtm_b200 import FMI, MergeParameters, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    bwtm_b200.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from datetime import timedelta
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=timedelta(minutes=30))

    thr = synth.error_threshold(args.error)
    n_a = args.reads * (args.read_len + 1); n_b = n_a
    # The sort-based fixture builder needs about 32 bytes per symbol: larger collections are built in chunks.
    chunk_reads = args.chunk_reads
    if chunk_reads == 0 and n_a > (3 << 30):
        chunk_reads = max(1, (2 << 30) // (args.read_len + 1))

    t_build = time.perf_counter()
    A = build_collection(FMI, args, thr, args.seed_a, args.reads, chunk_reads)
    B = build_collection(FMI, args, thr, args.seed_b, args.reads, chunk_reads)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    info_a, info_b = A.info(), B.info()
    assert info_a.bases == n_a and info_b.bases == n_b

    # The pair records (two backward steps per record read) belong to an index like the basic rank records K0 builds:
    # for the device-timed `value` they are resident with the inputs. (The e2e legs below start from host bytes and pay
    # for every structure they use inside the timed region.)
    if args.walk != "auto":
        os.environ["BWTM_WALK"] = args.walk
    pair_build_ms = None
    if args.walk != "single":
        torch.cuda.synchronize(); t0 = time.perf_counter()
        A.build_pairs(); B.build_pairs()
        torch.cuda.synchronize(); pair_build_ms = (time.perf_counter() - t0) * 1e3

    comm = None
    if world > 1:
        comm = bwtm_b200.Communicator.from_torch(dist, rank, world)

    params = MergeParameters()
    params.sequence_blocks = args.sequence_blocks

    def one_merge():
        if comm is None:
            return FMI.merge(A, B, params, keep_inputs=True)
        return comm.merge(A, B, params, keep_inputs=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- verification (outside the timed region): the first merge against the reference's result ----------------
    verified = None
    args.warmup = max(1, args.warmup)
    warmup_ms = []
    for it in range(args.warmup):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        M = one_merge()
        torch.cuda.synchronize(); warmup_ms.append(round((time.perf_counter() - t0) * 1e3, 2))
        if it == 0 and not args.no_verify and rank == 0:
            got = M.rle()
            digest = hashlib.sha256(got.tobytes()).hexdigest()
            known = load_digests().get(workload_key(args))
            if known is not None:
                ok = (known["sha256"] == digest and known["rle_bytes"] == len(got) and known["bases"] == M.size()
                      and known["sequences"] == M.sequences())
                verified = {"against": "unmodified reference binary: %s (tests/golden/reference_merge_digests.json)" % known.get("source", "oracle/_ref/bwt_merge"),
                            "ok": bool(ok), "sha256": digest, "rle_bytes": int(len(got))}
            if known is None or args.verify_direct:
                AB = build_collection(FMI, args, thr, args.seed_a, args.reads, chunk_reads)
                B2 = build_collection(FMI, args, thr, args.seed_b, args.reads, chunk_reads)
                if chunk_reads > 0 and chunk_reads < args.reads:
                    AB = FMI.merge(AB, B2, MergeParameters())       # no one-piece build at this size: sequential single-GPU route
                    how = "the same inputs merged by the single-GPU path (self-referential; no reference digest committed for this workload)"
                else:
                    AB.close(); B2.close()
                    AB = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_a, args.reads), (args.seed_b, args.reads)])
                    how = "BWT(A ++ B) built directly by the repo's fixture builder (self-referential)"
                same = bool(np.array_equal(AB.rle(), got)); AB.close()
                if verified is None:
                    verified = {"against": how, "ok": same, "sha256": digest, "rle_bytes": int(len(got))}
                else:
                    verified["direct_construction_ok"] = same
            del got
        M.close()
    if verified is not None and not verified["ok"]:
        print(json.dumps({"error": "merged BWT differs from the reference result", "verified": verified}))
        return 1

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    time.sleep(0.25)      # nvidia-smi needs a moment before its first sample; the timed region can be shorter than that
    launches0 = bwtm_b200.kernel_launches()
    stage = {k: 0.0 for k in ("search", "sort", "exchange", "interleave", "encode", "index", "pair_index")}
    last = None
    # Inputs smaller than twice the L2 would stay cached between steps: flush it (outside the timed events).
    flush = None
    if (n_a + n_b) // 2 < 2 * L2_BYTES:
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device="cuda")
    profiling = os.environ.get("BWTM_PROFILE_RANGE") == "1"   # ncu --profile-from-start off
    if profiling:
        torch.cuda.profiler.start()
    ms_total = 0.0
    step_ms_device = []
    memory_before = bwtm_b200.memory_stats(reset_peak=True)[0]
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1); torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        M = one_merge()
        end.record()
        end.synchronize()
        step_ms_device.append(start.elapsed_time(end)); ms_total += step_ms_device[-1]
        for k in stage:
            stage[k] += getattr(M.timings, k + "_seconds")
        last = M.timings.as_dict()
        merged_bytes = M.bytes()
        M.close()
    barrier()
    memory_peak = bwtm_b200.memory_stats()[1]
    if profiling:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = bwtm_b200.kernel_launches() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_b / (ms_step * 1e-3)
    for k in stage:
        stage[k] = stage[k] / args.steps

    # ---- e2e: host buffers through the C ABI -----------------------------------------------------
    e2e = None
    if not args.no_e2e and world == 1:
        rle_a = torch.from_numpy(A.rle()).pin_memory(); rle_b = torch.from_numpy(B.rle()).pin_memory()
        out = torch.empty(merged_bytes + 4096, dtype=torch.uint8).pin_memory()
        na, nb_, no = rle_a.numpy(), rle_b.numpy(), out.numpy()
        e2e_steps = max(1, args.steps)

        stream_out = MergeParameters(); stream_out.host_output = no    # merged bytes are copied out while encoding
        stream_out.sequence_blocks = args.sequence_blocks

        def e2e_step():
            a, b = FMI.from_rle_pair(na, nb_)
            m = FMI.merge(a, b, stream_out)
            got = int(m.timings.merged_bytes); m.close()
            return got
        for _ in range(args.warmup):
            e2e_step()
        step_ms = []
        for _ in range(e2e_steps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            got = e2e_step()
            torch.cuda.synchronize(); step_ms.append((time.perf_counter() - t0) * 1e3)
        e2e_s = float(np.mean(step_ms)) * 1e-3
        e2e = {"value": n_b / e2e_s, "unit": "bases/s", "h2d_bytes_per_step": int(len(na) + len(nb_)),
               "d2h_bytes_per_step": int(got), "ms_per_step": e2e_s * 1e3, "steps_ms": [round(x, 2) for x in step_ms]}

    if not args.no_e2e and world > 1 and os.environ.get("BWTM_BENCH_NO_DIST_E2E") is None:
        # N GPUs: every rank uploads its replica of both inputs from pinned host memory and builds K0, the merge is the
        # distributed one, rank 0 reads the merged bytes back. Wall clock between barriers, max over ranks.
        rle_a = torch.from_numpy(A.rle()).pin_memory(); rle_b = torch.from_numpy(B.rle()).pin_memory()
        out = torch.empty(merged_bytes + 4096, dtype=torch.uint8).pin_memory()
        na, nb_, no = rle_a.numpy(), rle_b.numpy(), out.numpy()

        def e2e_step_dist():
            a, b = FMI.from_rle_pair(na, nb_)
            m = comm.merge(a, b, params)
            got = int(m.download_into(no)) if rank == 0 else int(m.bytes())
            m.close()
            return got
        for _ in range(args.warmup):
            e2e_step_dist()
        step_ms = []
        for _ in range(max(1, args.steps)):
            barrier(); t0 = time.perf_counter()
            got = e2e_step_dist()
            barrier(); elapsed = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda")
            dist.all_reduce(elapsed, op=dist.ReduceOp.MAX); step_ms.append(float(elapsed.item()))
        e2e_s = float(np.mean(step_ms)) * 1e-3
        e2e = {"value": n_b / e2e_s, "unit": "bases/s", "h2d_bytes_per_step": int(world * (len(na) + len(nb_))),
               "d2h_bytes_per_step": int(got), "ms_per_step": e2e_s * 1e3, "steps_ms": [round(x, 2) for x in step_ms]}

    if rank != 0:
        return 0

    peak, peak_source = measured_peaks()
    k1_s = stage["search"]
    achieved = ALGORITHMIC_BYTES_PER_BASE * (n_b / world) / k1_s / 1e9 if k1_s > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k1_walk", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_source,
                "algorithmic_bytes_per_base": ALGORITHMIC_BYTES_PER_BASE, "kernel_ms": k1_s * 1e3}
    ncu_traffic = os.path.join(ROOT, "profiles", "k1_walk_traffic.json")
    if os.path.exists(ncu_traffic) and world == 1:
        try:   # one ncu --set full capture of this kernel on this workload (config 2); null for other workloads
            t = json.load(open(ncu_traffic))
            if t.get("algorithmic_bytes_per_launch") == int(ALGORITHMIC_BYTES_PER_BASE * n_b):
                roofline["traffic"] = t.get("dram_bytes_per_launch")
                roofline["traffic_kernel"] = t.get("kernel")
        except Exception:
            pass
    if world == 1:
        # Random-access denominator: dependent record reads in the kernel's own access shape over a table
        # of the size of the structures the walk reads. The microbenchmark has no reuse, so its rate is a rate of DRAM
        # MISSES; the kernel's requests are compared with it after removing the share that hits in L2 (ncu, committed).
        table = int(last.get("walk_table_bytes", 0)) or int(info_a.device_bytes + info_b.device_bytes)
        record_bytes = int(last.get("walk_record_bytes", 0)) or 64
        requests_per_base = (1.0 if record_bytes == 128 else 2.0)   # one record of each index per step; a pair record answers two steps
        peak_records = bwtm_b200.chase_bench(max(table, 1 << 28), record_bytes, 1 << 28, 2048) / float(record_bytes)
        achieved_records = requests_per_base * (n_b / world) / k1_s / 1e9 if k1_s > 0 else 0.0
        l2_hit = None
        try:
            t = json.load(open(ncu_traffic))
            if t.get("algorithmic_bytes_per_launch") == int(ALGORITHMIC_BYTES_PER_BASE * n_b):
                l2_hit = t.get("l2_hit_rate")
        except Exception:
            pass
        random = {"requests_per_second": achieved_records, "record_bytes": record_bytes, "requests_per_base": requests_per_base,
                  "peak_misses_per_second": peak_records, "unit": "G records/s", "l2_hit_rate": l2_hit,
                  "note": "peak = measured dependent random record reads without reuse (a DRAM miss rate)"}
        if l2_hit is not None and peak_records > 0:
            random["misses_per_second"] = achieved_records * (1.0 - l2_hit)
            random["frac"] = random["misses_per_second"] / peak_records
        roofline["random_records"] = random
    if args.gather_bench:
        roofline["random_access_gbs"] = {str(g): bwtm_b200.gather_bench(8 << 30, g, 1 << 28) for g in (32, 64, 128)}

    cpu = None
    if not args.no_cpu_baseline and world == 1 and os.path.exists(REF_MERGE) and os.path.exists(FIXTURE_TOOL):
        A.close(); B.close()      # the child process builds its own copy of A on this GPU
        tmp = scratch_dir()
        try:
            path_a = os.path.join(tmp, "A.native")
            info_file_a = write_fixture(args, path_a, [(args.seed_a, args.reads)], device=local_rank)
            res = reference_sample(args, tmp, path_a, info_file_a, os.cpu_count() or 1)
            cpu = {"value": res["value"], "unit": "bases/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": res["sample"]}
        except Exception as e:   # a bench line without the CPU leg is still a bench line
            cpu = {"value": None, "unit": "bases/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": "failed: %s" % e}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    line = {
        "metric": "merged_bases_per_second", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": config_dict(args, n_a, n_b, [info_a.rle_bytes, info_b.rle_bytes, merged_bytes]),
        "input_build_seconds": t_build, "steps_ms": [round(x, 2) for x in step_ms_device],
        "warmup_ms": warmup_ms,    # host wall clock of the untimed warm-up merges: the first one pays for the memory pools' growth
        "walk": {"record_bytes": int(last["walk_record_bytes"]), "table_bytes": int(last["walk_table_bytes"]),
                 "search_batches": int(last["search_batches"]),
                 "pair_records_build_ms": pair_build_ms,
                 "note": "pair records of both inputs built once before the timed region (bwtm_index_build_pairs), like K0; "
                         "e2e builds them inside its timed region" if pair_build_ms is not None else "single-step walk"},
        "verified": verified,
        "device_memory": {"inputs_resident_bytes": int(memory_before), "peak_bytes_during_merge": int(memory_peak),
                          "work_bytes": int(memory_peak - memory_before)},
        "stages_ms": {k: v * 1e3 for k, v in stage.items()},
        "stage_bases_per_second": {k: (n_b / v if v > 0 else None) for k, v in stage.items()},
        "ra_runs": last["ra_runs"], "merged_runs": last["merged_runs"],
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
