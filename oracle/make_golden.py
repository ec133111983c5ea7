"""Generates tests/golden/*.json with the UNMODIFIED reference (oracle/_ref/libref_hooks.so).

Run in the container that has /root/reference:  python oracle/make_golden.py
The reference ships no golden vectors; these are outputs of the reference's own
classes (Run::write, ByteCode, FMI load / rank / inverse_select / find, FMI(a, b, parameters)) on small
synthetic inputs, committed so that the oracle and the CUDA path can be checked where /root/reference
and oracle/_ref do not exist.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bwt-merge_b200"))

from oracle.oracle import Oracle, RefHooks, build   # noqa: E402
from bwtm_b200 import synth                          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [  # name, genome, reads A, reads B, read length, error rate, N fraction
    ("reads", 3000, 120, 80, 40, 0.01, 0.0),
    ("noisy_n", 1500, 90, 60, 33, 0.05, 0.03),
    ("repeats", 30, 200, 150, 25, 0.0, 0.0),
]


def collection(orc, G, n, L, e, seed, nfrac):
    g = synth.genome(G, 42)
    r = synth.reads(g, n, L, e, seed)
    if nfrac > 0:
        rng = np.random.default_rng(seed)
        r = r.copy(); r[rng.random(r.shape) < nfrac] = 5
    return r, orc.bwt_of_reads([row for row in r])


def main():
    build(ref=True)
    orc = Oracle(); ref = RefHooks()
    os.makedirs(OUT, exist_ok=True)

    kat = {"run_write": [], "bytecode": []}
    lengths = [1, 41, 42, 43, 82, 83, 84, 169, 170, 171, 1000, 16425, 16426, 100000, (1 << 21) + 42, (1 << 35) + 7]
    for off in (0, 1, 30, 55, 56, 57, 60, 61, 62, 63):
        for length in lengths:
            for comp in (0, 2, 5):
                out = ref.run_write(bytes(off), comp, length)
                kat["run_write"].append({"offset": off, "comp": comp, "length": length, "bytes": out[off:].hex()})
    for v in [0, 1, 127, 128, 16383, 16384, (1 << 21) - 1, 1 << 21, (1 << 64) - 1]:
        kat["bytecode"].append({"value": v, "bytes": ref.bytecode_write(v).hex()})
    json.dump(kat, open(os.path.join(OUT, "codec_kat.json"), "w"), indent=0)

    with tempfile.TemporaryDirectory() as tmp:
        for name, G, na, nb, L, e, nfrac in CASES:
            ra, bwt_a = collection(orc, G, na, L, e, 1, nfrac)
            rb, bwt_b = collection(orc, G, nb, L, e, 2, nfrac)
            fa, fb = os.path.join(tmp, "A"), os.path.join(tmp, "B")
            synth.comps_to_chars(bwt_a).tofile(fa); synth.comps_to_chars(bwt_b).tofile(fb)
            A, B = ref.load(fa), ref.load(fb)
            rng = np.random.default_rng(5)
            pos = [int(x) for x in rng.integers(0, A.size + 1, 60)] + [0, A.size]
            ranks = [[A.rank(i, c) for c in range(6)] for i in pos]
            lf_pos = [int(x) for x in rng.integers(0, A.size, 60)]
            lf = [list(A.inverse_select(i)) for i in lf_pos]
            g = synth.genome(G, 42)
            pats = [p.tolist() for p in synth.patterns(g, 25, min(9, G // 2), 99)]
            found = [list(A.find(synth.comps_to_chars(np.array(p, np.uint8)).tobytes())) for p in pats]
            ends, cum = A.samples()
            case = {
                "name": name, "params": {"genome": G, "reads_a": na, "reads_b": nb, "read_len": L, "error": e, "n_frac": nfrac},
                "bwt_a": bytes(synth.comps_to_chars(bwt_a)).decode(), "bwt_b": bytes(synth.comps_to_chars(bwt_b)).decode(),
                "rle_a": bytes(A.rle()).hex(), "rle_b": bytes(B.rle()).hex(),
                "C_a": [int(x) for x in A.C()], "hash_a": A.hash(),
                "rank_positions": pos, "ranks": ranks, "lf_positions": lf_pos, "lf": lf,
                "patterns": pats, "find": found,
                "block_ends_a": [int(x) for x in ends], "cumulative_a": [[int(x) for x in row] for row in cum],
            }
            M = ref.merge(A, B, threads=2, sequence_blocks=5, temp_dir=tmp)
            case.update({"rle_merged": bytes(M.rle()).hex(), "hash_merged": M.hash(), "C_merged": [int(x) for x in M.C()],
                         "sequences_merged": M.sequences, "size_merged": M.size,
                         "find_merged": [list(M.find(synth.comps_to_chars(np.array(p, np.uint8)).tobytes())) for p in pats]})
            json.dump(case, open(os.path.join(OUT, "merge_%s.json" % name), "w"))
            print("wrote", name, A.size, B.size, M.bytes)


if __name__ == "__main__":
    main()
