"""TEST INFRASTRUCTURE -- digests of merges done by the UNMODIFIED reference binary (oracle/_ref/bwt_merge).

  python oracle/make_reference_digests.py --config 1        # CPU only: inputs from the oracle's suffix sorter

For configurations small enough for the CPU, the inputs are built WITHOUT the CUDA library: synthetic reads from
bwtm_b200/synth.py (numpy), BWT by the oracle's suffix sorter (oracle/bwtm_oracle.c), native files by the reference's
own bwt_convert. The digest entry (SHA-256 of the merged run-length bytes, byte count, sequences, bases) goes to
tests/golden/reference_merge_digests.json under the same key bench.py uses, so `bench.py --config 1` checks the GPU
result -- and the GPU fixture builder, which has to produce the same inputs -- against it.
Configurations beyond the CPU sorter (config 2) are recorded on the GPU box by `bench.py --impl reference
--record-digest`, where bin/bwtm_fixture builds the input files and the reference binary merges them.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.abspath(__file__))]   # oracle/ is a package here
for p in (ROOT, os.path.join(ROOT, "bwt-merge_b200")):
    sys.path.insert(0, p)

import bench                                  # noqa: E402  (workload_key, native_file_digest, CONFIGS)
from bwtm_b200 import synth                   # noqa: E402  (numpy only)
from oracle.oracle import Oracle, REF_DIR     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    opts = ap.parse_args()
    sys.argv = [sys.argv[0], "--config", str(opts.config)]
    args = bench.parse_args()
    orc = Oracle()
    genome = synth.genome(args.genome, args.genome_seed)
    with tempfile.TemporaryDirectory(prefix="bwtm_digest_") as tmp:
        natives = []
        for name, seed in (("A", args.seed_a), ("B", args.seed_b)):
            reads = synth.reads(genome, args.reads, args.read_len, args.error, seed)
            bwt = orc.bwt_of_reads([row for row in reads])
            plain, native = os.path.join(tmp, name + ".plain"), os.path.join(tmp, name + ".native")
            synth.comps_to_chars(bwt).tofile(plain)
            subprocess.check_call([os.path.join(REF_DIR, "bwt_convert"), "-i", "plain_default", "-o", "native", plain, native],
                                  stdout=subprocess.DEVNULL)
            natives.append(native)
        out = os.path.join(tmp, "merged.native")
        timers = bench.run_reference_binary(natives[0], natives[1], out, opts.threads, tmp)
        entry = bench.native_file_digest(out)
    entry["source"] = ("oracle/_ref/bwt_merge -t %d (unmodified reference, SDSL stand-in); inputs built on the CPU by the oracle's "
                       "suffix sorter from bwtm_b200/synth.py reads (oracle/make_reference_digests.py)" % opts.threads)
    digests = bench.load_digests(); digests[bench.workload_key(args)] = entry
    json.dump(digests, open(bench.DIGESTS, "w"), indent=1, sort_keys=True)
    print(json.dumps({bench.workload_key(args): entry, "reference_timers": timers}, indent=1))


if __name__ == "__main__":
    main()
