"""Pins oracle/bwtm_oracle.c against the UNMODIFIED reference built into oracle/_ref
(in-process through libref_hooks.so and through the bwt_merge binary)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import make_collection
from bwtm_b200 import synth


LENGTHS = [1, 2, 40, 41, 42, 43, 82, 83, 84, 42 + 127, 42 + 128, 42 + 129, 1000, 42 + 16383, 42 + 16384,
           100000, (1 << 21) + 41, (1 << 21) + 42, (1 << 35) + 7, (1 << 62)]


def test_bytecode_kat(oracle, refhooks):
    for v in [0, 1, 127, 128, 129, 16383, 16384, (1 << 21) - 1, 1 << 21, (1 << 63) + 5, (1 << 64) - 1]:
        enc = oracle.bytecode_write(v)
        assert enc == refhooks.bytecode_write(v)
        assert oracle.bytecode_read(enc) == (v, len(enc)) == refhooks.bytecode_read(enc)


def test_run_write_every_offset(oracle, refhooks):
    """Run::write for every entry offset 0..63 and the edge lengths of SURVEY appendix A.1."""
    for off in range(64):
        prefix = bytes([0] * off)  # `off` runs ($, 1)
        for length in LENGTHS:
            for comp in (0, 3, 5):
                assert oracle.run_write(prefix, comp, length) == refhooks.run_write(prefix, comp, length), (off, length, comp)


def test_run_write_known_answers(oracle):
    # len 42 exactly -> head 246+c and a zero extension byte (bit_length(0) == 1)
    assert oracle.run_write(b"", 1, 42) == bytes([6 * 41 + 1, 0])
    assert oracle.run_write(b"", 1, 41) == bytes([6 * 40 + 1])
    # one byte left in the block: 41-head, rest continues in the next block
    assert oracle.run_write(bytes(63), 2, 42) == bytes(63) + bytes([6 * 40 + 2, 2])
    assert oracle.run_write(bytes(63), 2, 83) == bytes(63) + bytes([6 * 40 + 2, 6 * 41 + 2, 0])
    # two bytes left and an extension that needs two bytes: truncated to 127
    assert oracle.run_write(bytes(62), 4, 42 + 128) == bytes(62) + bytes([6 * 41 + 4, 127, 4])


def _write_plain(path, comps):
    synth.comps_to_chars(comps).tofile(path)


@pytest.mark.parametrize("shape", [(3000, 150, 40, 0.01, 0.0), (500, 300, 25, 0.05, 0.02), (64, 400, 30, 0.0, 0.0)])
def test_queries_and_merge_vs_reference(oracle, refhooks, tmp_path, shape):
    G, n, L, e, nfrac = shape
    ra, bwt_a = make_collection(oracle, G, n, L, e, 42, 1, nfrac)
    rb, bwt_b = make_collection(oracle, G, n // 2 + 1, L, e, 42, 2, nfrac)
    fa, fb = str(tmp_path / "A.plain"), str(tmp_path / "B.plain")
    _write_plain(fa, bwt_a); _write_plain(fb, bwt_b)

    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    RA, RB = refhooks.load(fa), refhooks.load(fb)
    assert np.array_equal(A.rle(), RA.rle()) and np.array_equal(B.rle(), RB.rle())
    assert A.size == RA.size and A.sequences == RA.sequences and A.hash() == RA.hash()
    assert np.array_equal(A.C(), RA.C())
    ea, ca = A.samples(); er, cr = RA.samples()
    assert np.array_equal(ea, er) and np.array_equal(ca, cr)

    rng = np.random.default_rng(7)
    for i in list(rng.integers(0, A.size + 1, 400)) + [0, A.size, A.size + 5]:
        for c in range(6):
            assert A.rank(i, c) == RA.rank(i, c)
        assert np.array_equal(A.ranks(i)[1:], RA.ranks(i)[1:])
        if i < A.size:
            assert A.inverse_select(i) == RA.inverse_select(i)
            assert A.access(i) == RA.access(i)
    for sp in rng.integers(0, A.size - 1, 100):
        ep = min(A.size - 1, sp + int(rng.integers(1, 256)))
        f, s = A.ranks_range(sp, ep); rf, rs = RA.ranks_range(sp, ep)
        occ = s[1:] > f[1:]
        assert np.array_equal(occ, rs[1:] > rf[1:])
        assert np.array_equal(f[1:][occ], rf[1:][occ]) and np.array_equal(s[1:][occ], rs[1:][occ])

    g = synth.genome(G, 42)
    pats = synth.patterns(g, 50, 12, 99)
    for p in pats:
        assert A.find(p) == RA.find(synth.comps_to_chars(p).tobytes())

    M_dfs = oracle.merge(A, B, use_dfs=True)
    M_walk = oracle.merge(A, B, use_dfs=False)
    RM = refhooks.merge(RA, RB, threads=3, sequence_blocks=7, temp_dir=str(tmp_path))
    ref_rle = RM.rle()
    assert np.array_equal(M_dfs.rle(), ref_rle)
    assert np.array_equal(M_walk.rle(), ref_rle)
    assert M_dfs.hash() == RM.hash() and np.array_equal(M_dfs.C(), RM.C())

    # definition: merge(BWT(A), BWT(B)) == BWT(A ++ B)
    direct = oracle.bwt_of_reads([r for r in ra] + [r for r in rb])
    assert np.array_equal(M_dfs.decode(), direct)


def test_reference_binary_invariance(oracle, tmp_path):
    """bwt_merge output does not depend on -t/-s/-r/-b/-m and equals the oracle."""
    from oracle.oracle import REF_DIR, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ra, bwt_a = make_collection(oracle, 2000, 300, 50, 0.01, 42, 1)
    rb, bwt_b = make_collection(oracle, 2000, 200, 50, 0.01, 42, 2)
    fa, fb = str(tmp_path / "A.plain"), str(tmp_path / "B.plain")
    _write_plain(fa, bwt_a); _write_plain(fb, bwt_b)
    outs = []
    for k, opts in enumerate([["-t", "1"], ["-t", "4", "-s", "50"], ["-t", "2", "-r", "1", "-b", "1", "-m", "1"]]):
        out = str(tmp_path / ("M%d.plain" % k))
        subprocess.check_call([os.path.join(REF_DIR, "bwt_merge"), "-i", "plain_default", "-o", "plain_default",
                               "-d", str(tmp_path)] + opts + [fa, fb, out],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        outs.append(np.fromfile(out, dtype=np.uint8))
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    M = oracle.merge(oracle.from_comps(bwt_a), oracle.from_comps(bwt_b))
    assert np.array_equal(synth.comps_to_chars(M.decode()), outs[0])
