#!/usr/bin/env python
"""Config 4 of BASELINE.json: sequential multi-input merge of 8 collections (50, 40, 30, 20, 15, 10, 7, 5 M x 150 bp reads
of a 500 Mbp genome, largest to smallest) with -v pattern verification (1 M 32-mers), through the command line.

  python scripts/run_config4.py --scale 1.0                 # full size: bin/bwt_merge_b200 -v, then an independent route
  python scripts/run_config4.py --scale 0.1 --reference     # 1/10 size, also oracle/_ref/bwt_merge: same file, same report

Inputs are native files written by bin/bwtm_fixture. Checks:
  * the tool's own -v verdict (pattern counts of the inputs add up to those of the output, bwt_merge.cpp:178-194);
  * --reference: the unmodified reference binary on the same files: byte-identical output file and identical report;
  * an independent route through the library: the same eight inputs merged as a balanced tree ((1+2)+(3+4))+((5+6)+(7+8))
    (other intermediate sizes, other kernels' regimes) must give the same run-length bytes as the sequential merge.
Prints one JSON line.
"""
import argparse
import filecmp
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bwt-merge_b200"))
import numpy as np          # noqa: E402
import bench                # noqa: E402

READS_M = [50, 40, 30, 20, 15, 10, 7, 5]
TOOL = os.path.join(ROOT, "bwt-merge_b200", "bin", "bwt_merge_b200")
REF_OVER_ABI = os.path.join(ROOT, "oracle", "_ref", "bwt_merge_b200")


def pattern_rows(genome_len, genome_seed, n, length, seed):
    """n length-mers at uniform genome offsets as text rows; only the genome positions that are needed are generated
    (same function of the index as bwtm_b200.synth.genome / synth.patterns)."""
    from bwtm_b200 import synth
    starts = synth.rnd(seed, 4, np.arange(n, dtype=np.uint64)) % np.uint64(genome_len - length + 1)
    idx = starts[:, None] + np.arange(length, dtype=np.uint64)[None, :]
    comps = ((synth.rnd(genome_seed, 0, idx.reshape(-1)) & np.uint64(3)) + np.uint64(1)).astype(np.uint8).reshape(n, length)
    chars = synth.COMP2CHAR[comps]
    return [row.tobytes().decode() for row in chars]


def report_lines(stdout):
    keep = []
    for line in stdout.splitlines():
        m = re.match(r"(Input|Output):\s+Found (\d+) patterns with (\d+) occ", line)
        if m:
            keep.append(list(m.groups()))
        m = re.match(r"(Input|Output):\s+([0-9.e+-]+) MB \(([0-9.e+-]+) bpc\)", line)
        if m:
            keep.append(list(m.groups()))
        if line.startswith(("Verification", "Read ")):
            keep.append(line)
    return keep


def run_tool(tool, files, out, patterns, tmp, threads):
    t0 = time.perf_counter()
    res = subprocess.run([tool, "-t", str(threads), "-d", tmp, "-v", patterns] + files + [out], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        raise RuntimeError("%s failed: %s" % (tool, res.stderr[-800:]))
    merges = [float(x) for x in re.findall(r"BWTs merged in ([0-9.e+-]+) seconds \(", res.stdout)]
    stage = {"search+sort": [float(x) for x in re.findall(r"RA built in ([0-9.e+-]+) seconds", res.stderr)],
             "interleave+encode": [float(x) for x in re.findall(r"bwt_merge: BWTs merged in ([0-9.e+-]+) seconds", res.stderr)],
             "index": [float(x) for x in re.findall(r"rank/select built in ([0-9.e+-]+) seconds", res.stderr)]}
    return {"wall_seconds": wall, "merge_seconds": merges, "stage_seconds": stage, "report": report_lines(res.stdout),
            "verification_successful": "Verification successful" in res.stdout, "stdout_tail": res.stdout[-600:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reference", action="store_true", help="also run oracle/_ref/bwt_merge on the same files and compare")
    ap.add_argument("--reference-cli-over-abi", action="store_true", help="also run the reference's own CLI bound to the C ABI")
    ap.add_argument("--patterns", type=int, default=0)
    ap.add_argument("--no-tree", action="store_true")
    opts = ap.parse_args()
    genome = int(500_000_000 * opts.scale); read_len = 150; error = 0.01; genome_seed = 42
    reads = [max(1, int(m * 1_000_000 * opts.scale)) for m in READS_M]
    n_patterns = opts.patterns or max(1000, int(1_000_000 * opts.scale))
    threads = os.cpu_count() or 1

    class A:   # the arguments bench.write_fixture reads
        pass
    args = A(); args.genome = genome; args.genome_seed = genome_seed; args.read_len = read_len; args.error = error
    tmp = bench.scratch_dir()
    try:
        t0 = time.perf_counter()
        files, rles, infos = [], [], []
        for k, n in enumerate(reads):
            path = os.path.join(tmp, "in%d.native" % (k + 1)); rle = os.path.join(tmp, "in%d.rle" % (k + 1))
            cmd = [bench.FIXTURE_TOOL, "--genome", str(genome), "--genome-seed", str(genome_seed), "--read-len", str(read_len), "--error", repr(error),
                   "--segment", "%d:%d" % (k + 1, n), "--format", "native", "--output", path, "--rle-output", rle]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("bwtm_fixture failed: " + res.stderr[-500:])
            infos.append(json.loads(res.stdout.strip().splitlines()[-1])); files.append(path); rles.append(rle)
        build_seconds = time.perf_counter() - t0
        patterns = os.path.join(tmp, "patterns.txt")
        with open(patterns, "w") as f:
            f.write("\n".join(pattern_rows(genome, genome_seed, n_patterns, 32, 99)) + "\n")

        out = os.path.join(tmp, "merged.native")
        mine = run_tool(TOOL, files, out, patterns, tmp, threads)
        digest = bench.native_file_digest(out)
        total_inserted = sum(i["bases"] for i in infos[1:])
        result = {"config": "sequential multi-input merge of 8 collections (%s reads x %d bp, %d bp genome) with -v (%d 32-mers)" % (
                      ", ".join(str(r) for r in reads), read_len, genome, n_patterns),
                  "scale": opts.scale, "symbols": [i["bases"] for i in infos], "merged_symbols": digest["bases"], "merged_rle_bytes": digest["rle_bytes"],
                  "sha256": digest["sha256"], "input_build_seconds": build_seconds,
                  "tool": "bin/bwt_merge_b200 -v (host driver above the C ABI, one B200)", "tool_run": mine,
                  "inserted_bases": total_inserted, "merged_bases_per_second": total_inserted / sum(mine["merge_seconds"]),
                  "verified": {"minus_v": mine["verification_successful"]}}

        if opts.reference_cli_over_abi and os.path.exists(REF_OVER_ABI):
            out2 = os.path.join(tmp, "merged_refcli.native")
            other = run_tool(REF_OVER_ABI, files, out2, patterns, tmp, threads)
            result["reference_cli_over_abi"] = {"run": other, "same_file": filecmp.cmp(out, out2, shallow=False), "same_report": other["report"] == mine["report"]}
            os.unlink(out2)
        if opts.reference:
            out3 = os.path.join(tmp, "merged_ref.native")
            ref = run_tool(bench.REF_MERGE, files, out3, patterns, tmp, threads)
            result["reference"] = {"tool": "oracle/_ref/bwt_merge -t %d (unmodified reference)" % threads, "run": ref,
                                   "merged_bases_per_second": total_inserted / sum(ref["merge_seconds"])}
            result["verified"]["same_file_as_reference"] = filecmp.cmp(out, out3, shallow=False)
            result["verified"]["same_report_as_reference"] = (ref["report"] == mine["report"])
            os.unlink(out3)

        if not opts.no_tree:
            import bwtm_b200
            from bwtm_b200 import FMI
            bwtm_b200.set_device(0)
            t0 = time.perf_counter()
            level = [FMI.from_rle(np.fromfile(p, dtype=np.uint8)) for p in rles]
            while len(level) > 1:
                level = [FMI.merge(level[k], level[k + 1]) if k + 1 < len(level) else level[k] for k in range(0, len(level), 2)]
            tree = level[0].rle()
            result["verified"]["same_bytes_as_balanced_tree_route"] = (hashlib.sha256(tree.tobytes()).hexdigest() == digest["sha256"] and len(tree) == digest["rle_bytes"])
            result["tree_route_seconds"] = time.perf_counter() - t0
        result["verified"]["ok"] = all(bool(v) for v in result["verified"].values())
        print(json.dumps(result))
        return 0 if result["verified"]["ok"] else 1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
