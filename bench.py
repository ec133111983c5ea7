#!/usr/bin/env python
"""bench.py -- merged bases/s of the rank-array path (BASELINE.json metric) on B200.

  python bench.py --gpus 1 --steps 3 --warmup 3          # our arm, config 2 of BASELINE.json
  python bench.py --impl reference --steps 2 --warmup 1   # the reference's CPU path on the host cores

A step is one merge of the synthetic collections A and B (B inserted into A):
  * `value`  : |B| / device time of bwtm_merge with both indexes resident in HBM (CUDA events);
  * `e2e`    : the same through the C ABI with HOST buffers: upload of both run-length BWTs from pinned
               memory + rank-structure build (K0) + merge + download of the merged run-length BWT;
  * `roofline`: the rank/LF walk kernel (K1), algorithmic bytes 168 B per inserted base;
  * `cpu_baseline`: the unmodified reference (oracle/_ref) on the host cores on a bounded sample.
Inputs are built on the device by the fixture builder (outside every timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "bwt-merge_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

ALGORITHMIC_BYTES_PER_BASE = 168.0   # LF_B 80 B + rank_A 80 B + 8 B RA value


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload: config 2 of BASELINE.json
    ap.add_argument("--genome", type=int, default=50_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--error", type=float, default=0.01)
    ap.add_argument("--genome-seed", type=int, default=42)
    ap.add_argument("--seed-a", type=int, default=1)
    ap.add_argument("--seed-b", type=int, default=2)
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000, help="reads of B merged by the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify", action="store_true", help="also build BWT(A ++ B) directly and compare bytes")
    ap.add_argument("--gather-bench", action="store_true", help="measure the random-access HBM peak as well")
    return ap.parse_args()


def workload_name(args):
    return "two-input merge 2x%dx%dbp reads, synthetic %d bp genome, %.0f%% substitutions" % (
        args.reads, args.read_len, args.genome, 100 * args.error)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
# CPU reference leg (the only place oracle/ is executed from bench.py)

class StderrCapture:
    """Redirects fd 2 to a file so that the reference's VERBOSE_STATUS_INFO stage timers can be read."""
    def __enter__(self):
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        sys.stderr.flush()
        self.saved = os.dup(2); os.dup2(self.tmp.fileno(), 2)
        return self

    def __exit__(self, *exc):
        os.dup2(self.saved, 2); os.close(self.saved)
        self.tmp.seek(0); self.text = self.tmp.read().decode(errors="replace"); self.tmp.close()

    def timer(self, label):
        for line in self.text.splitlines():
            if label in line:
                try:
                    return float(line.split(label)[1].split()[0])
                except Exception:
                    pass
        return None


def reference_merge_rate(args, rle_a, rle_b_sample, n_a, n_b, steps, warmup):
    """Times FMI::FMI(a, b, parameters) of the unmodified reference (oracle/_ref/libref_hooks.so) with
    -t nproc on A = the full collection and B = a prefix of collection B, and scales the stage times to
    the full workload. Returns (bases/s, cores, description, ms per step)."""
    from oracle.oracle import Oracle, RefHooks, ref_available
    from bwtm_b200 import synth
    if not ref_available():
        return None
    hooks = RefHooks(); orc = Oracle()
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="bwtm_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = []
        for name, rle in (("A", rle_a), ("B", rle_b_sample)):
            path = os.path.join(tmp, name + ".plain")
            synth.comps_to_chars(orc.from_rle(rle).decode()).tofile(path)
            paths.append(path)
        fa, fb = hooks.load(paths[0]), hooks.load(paths[1])
        n_bs = fb.size
        times = []
        for it in range(warmup + steps):
            ca, cb = hooks.copy(fa), hooks.copy(fb)
            with StderrCapture() as cap:
                t0 = time.perf_counter()
                merged = hooks.merge_params(ca, cb, threads=cores, temp_dir=tmp)
                wall = time.perf_counter() - t0
            ra = cap.timer("RA built in"); il = cap.timer("BWTs merged in"); sa = cap.timer("rank/select built in")
            del merged, ca, cb
            if it >= warmup:
                times.append((wall, ra, il, sa))
        wall = float(np.mean([t[0] for t in times]))
        ra = float(np.mean([t[1] if t[1] is not None else t[0] for t in times]))
        il = float(np.mean([t[2] or 0.0 for t in times])); sa = float(np.mean([t[3] or 0.0 for t in times]))
        scale_b = n_b / n_bs; scale_m = (n_a + n_b) / (n_a + n_bs)
        full = ra * scale_b + (il + sa) * scale_m
        desc = ("unmodified reference (oracle/_ref, SDSL stand-in) -t %d, defaults; A = full %d-symbol collection, B = first %d "
                "reads of collection B (%d symbols); measured RA %.2f s, interleave %.2f s, samples %.2f s, merge wall %.2f s; "
                "RA time scaled by |B|/|B_sample| = %.1f and interleave+samples by merged length %.2f to the full workload"
                % (cores, n_a, args.cpu_sample_reads, n_bs, ra, il, sa, wall, scale_b, scale_m))
        return n_b / full, cores, desc, wall * 1e3
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import bwtm_b200
    from bwtm_b200 import FMI, MergeParameters, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    bwtm_b200.set_device(local_rank)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    thr = synth.error_threshold(args.error)
    n_a = args.reads * (args.read_len + 1); n_b = n_a

    # ---- reference arm ---------------------------------------------------------------------------
    if args.impl == "reference":
        A = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_a, args.reads)])
        Bs = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_b, min(args.cpu_sample_reads, args.reads))])
        rle_a, rle_bs = A.rle(), Bs.rle()
        A.close(); Bs.close()
        res = reference_merge_rate(args, rle_a, rle_bs, n_a, n_b, args.steps, max(1, args.warmup))
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
            return 0
        value, cores, desc, ms = res
        print(json.dumps({
            "impl": "reference", "metric": "merged_bases_per_second", "value": value, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args)},
            "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "reference", "sample": desc},
            "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---- our arm ---------------------------------------------------------------------------------
    t_build = time.perf_counter()
    A = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_a, args.reads)])
    B = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_b, args.reads)])
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    info_a, info_b = A.info(), B.info()
    assert info_a.bases == n_a and info_b.bases == n_b

    comm = None
    if world > 1:
        comm = bwtm_b200.Communicator.from_torch(dist, rank, world)

    params = MergeParameters()

    def one_merge():
        if comm is None:
            return FMI.merge(A, B, params, keep_inputs=True)
        return comm.merge(A, B, params, keep_inputs=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    reference_bytes = None
    if args.verify and rank == 0:
        AB = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_a, args.reads), (args.seed_b, args.reads)])
        reference_bytes = AB.rle(); AB.close()

    # The stream-ordered allocator pool settles after a few identical steps (a first step can cost 50+ ms more):
    # never fewer than three warm-up merges, whatever was asked for.
    args.warmup = max(3, args.warmup)
    for _ in range(args.warmup):
        M = one_merge()
        if reference_bytes is not None:
            assert np.array_equal(M.rle(), reference_bytes), "merged BWT differs from the directly built BWT(A ++ B)"
            reference_bytes = None
        M.close()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = bwtm_b200.kernel_launches()
    stage = {k: 0.0 for k in ("search", "sort", "exchange", "interleave", "encode", "index")}
    last = None
    # Inputs smaller than twice the L2 would stay cached between steps: flush it (outside the timed events).
    l2_bytes = 126 << 20
    flush = None
    if info_a.device_bytes + info_b.device_bytes < 2 * l2_bytes:
        flush = torch.empty(2 * l2_bytes, dtype=torch.uint8, device="cuda")
    profiling = os.environ.get("BWTM_PROFILE_RANGE") == "1"   # ncu --profile-from-start off
    if profiling:
        torch.cuda.profiler.start()
    ms_total = 0.0
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1); torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        M = one_merge()
        end.record()
        end.synchronize()
        ms_total += start.elapsed_time(end)
        for k in stage:
            stage[k] += getattr(M.timings, k + "_seconds")
        last = M.timings.as_dict()
        merged_bytes = M.bytes()
        M.close()
    barrier()
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

        def e2e_step():
            a, b = FMI.from_rle_pair(na, nb_)
            m = FMI.merge(a, b, stream_out)
            got = int(m.timings.merged_bytes); m.close()
            return got
        for _ in range(max(3, args.warmup)):     # the allocator pool settles after a few identical steps
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
        for _ in range(max(3, args.warmup)):
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
        except Exception:
            pass
    if world == 1:
        # Random-access denominator: dependent 64-byte record reads in the kernel's own access shape
        # over a table of the size of both rank structures; K1 reads two records per inserted base.
        table = int(info_a.device_bytes + info_b.device_bytes)
        peak_records = bwtm_b200.chase_bench(max(table, 1 << 28), 64, 1 << 28, 2048) / 64.0
        achieved_records = 2.0 * n_b / k1_s / 1e9 if k1_s > 0 else 0.0
        roofline["random_records"] = {"achieved": achieved_records, "peak": peak_records, "unit": "G records/s",
                                      "frac": achieved_records / peak_records if peak_records > 0 else None,
                                      "note": "peak = measured dependent random 64-byte record reads (no reuse); the kernel exceeds it "
                                              "through L2 hits and its coalesced first step"}
    if args.gather_bench:
        roofline["random_access_gbs"] = {str(g): bwtm_b200.gather_bench(8 << 30, g, 1 << 28) for g in (32, 64, 128)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        Bs = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_b, min(args.cpu_sample_reads, args.reads))])
        rle_bs = Bs.rle(); Bs.close()
        res = reference_merge_rate(args, A.rle(), rle_bs, n_a, n_b, 1, 0)
        if res is not None:
            cpu = {"value": res[0], "unit": "bases/s", "cores": res[1], "kind": "reference", "sample": res[2]}

    line = {
        "metric": "merged_bases_per_second", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "inserted_bases": n_b, "merged_symbols": n_a + n_b,
                   "rle_bytes": [int(info_a.rle_bytes), int(info_b.rle_bytes), int(merged_bytes)],
                   "l2": ("inputs (2 x %.2f GB of rank records) exceed the 126 MB L2; no flush" % (info_a.device_bytes / 1e9)
                          if flush is None else "inputs fit in L2: a 252 MB buffer is written between timed steps"),
                   "input_build_seconds": t_build},
        "stages_ms": {k: v * 1e3 for k, v in stage.items()},
        "stage_bases_per_second": {k: (n_b / v if v > 0 else None) for k, v in stage.items()},
        "ra_runs": last["ra_runs"], "merged_runs": last["merged_runs"],
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
