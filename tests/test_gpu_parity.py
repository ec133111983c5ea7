"""GPU parity: the CUDA path behind the C ABI against the CPU oracle (bit-exact integer/byte work)."""
import os
import subprocess

import numpy as np
import pytest

import bwtm_b200
from bwtm_b200 import FMI, MergeParameters, synth
from conftest import make_collection

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _library():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    bwtm_b200.lib()


SHAPES = {
    "reads":     dict(G=5000, n=400, L=60, e=0.01, nfrac=0.0),
    "noisy_N":   dict(G=3000, n=300, L=45, e=0.05, nfrac=0.03),
    "repeats":   dict(G=40, n=500, L=30, e=0.0, nfrac=0.0),      # tiny genome: very long runs
    "single":    dict(G=200, n=1, L=150, e=0.0, nfrac=0.0),
}


def collections(oracle, shape):
    s = SHAPES[shape]
    ra, bwt_a = make_collection(oracle, s["G"], s["n"], s["L"], s["e"], 42, 1, s["nfrac"])
    rb, bwt_b = make_collection(oracle, s["G"], max(1, s["n"] // 3), s["L"], s["e"], 42, 2, s["nfrac"])
    return ra, bwt_a, rb, bwt_b


@pytest.mark.parametrize("shape", list(SHAPES))
def test_index_and_queries(oracle, shape):
    ra, bwt_a, rb, bwt_b = collections(oracle, shape)
    A = oracle.from_comps(bwt_a)
    D = FMI.from_rle(A.rle(), expected_counts=A.counts())
    assert D.size() == A.size and D.sequences() == A.sequences and D.bytes() == A.bytes
    assert np.array_equal(D.counts(), A.counts()) and np.array_equal(D.C(), A.C())
    assert np.array_equal(D.rle(), A.rle())
    assert np.array_equal(D.extract(), A.decode())
    assert D.hash() == A.hash()
    ends, cum = D.samples(); oe, oc = A.samples()
    assert np.array_equal(ends, oe) and np.array_equal(cum, oc)

    rng = np.random.default_rng(3)
    pos = np.concatenate([rng.integers(0, A.size + 1, 3000), [0, A.size, A.size + 9]]).astype(np.uint64)
    comps = rng.integers(0, 6, len(pos)).astype(np.uint8)
    got = D.rank(pos, comps)
    want = np.array([A.rank(int(i), int(c)) for i, c in zip(pos, comps)], dtype=np.uint64)
    assert np.array_equal(got, want)

    pos = rng.integers(0, A.size, 3000).astype(np.uint64)
    nxt, cc = D.LF(pos)
    C_ = A.C()
    for i, n_, c_ in zip(pos, nxt, cc):
        r, c = A.inverse_select(int(i))
        assert (int(n_), int(c_)) == (r + int(C_[c]), c)

    g = synth.genome(SHAPES[shape]["G"], 42)
    pats = [p for p in synth.patterns(g, 64, min(12, SHAPES[shape]["G"] // 2), 99)] + [np.array([5, 5, 5], np.uint8), np.array([0], np.uint8)]
    got = D.count(pats)
    want = np.array([A.count(p) for p in pats], dtype=np.uint64)
    assert np.array_equal(got, want)
    chars = [synth.comps_to_chars(p).tobytes() for p in pats]
    assert np.array_equal(D.count(chars, char2comp=bwtm_b200.DEFAULT_CHAR2COMP), want)


@pytest.mark.parametrize("shape", list(SHAPES))
def test_rank_array(oracle, shape):
    ra, bwt_a, rb, bwt_b = collections(oracle, shape)
    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    DA, DB = FMI.from_rle(A.rle()), FMI.from_rle(B.rle())
    got = bwtm_b200.rank_array(DA, DB)
    want = np.sort(oracle.build_ra_walk(A, B))
    assert np.array_equal(got, want)
    # expanding the reference's (pos, len) runs from the DFS gives the same multiset
    runs = oracle.build_ra_dfs(A, B)
    assert np.array_equal(np.sort(np.repeat(runs[:, 0], runs[:, 1].astype(np.int64))), got)
    if B.sequences > 2:
        part = bwtm_b200.rank_array(DA, DB, 1, B.sequences - 2)
        assert np.array_equal(part, np.sort(oracle.build_ra_walk(A, B, 1, B.sequences - 2)))


@pytest.mark.parametrize("slab", [0, 4096, 8192])
@pytest.mark.parametrize("shape", list(SHAPES))
def test_merge_bit_exact(oracle, shape, slab):
    ra, bwt_a, rb, bwt_b = collections(oracle, shape)
    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    want = oracle.merge(A, B, use_dfs=True)
    p = MergeParameters(); p.slab_symbols = slab
    M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
    assert np.array_equal(M.rle(), want.rle())
    assert np.array_equal(M.counts(), want.counts()) and M.sequences() == want.sequences and M.size() == want.size
    assert M.hash() == want.hash()
    assert M.timings.ra_values == B.size
    assert M.timings.ra_runs == len(oracle.sort_compress(oracle.build_ra_dfs(A, B)))   # the reference's RA run count
    # definition: the merge is the BWT of A's reads followed by B's reads
    direct = oracle.bwt_of_reads([r for r in ra] + [r for r in rb])
    assert np.array_equal(M.extract(), direct)
    # -v: per-pattern counts add up (bwt_merge.cpp:178-194)
    g = synth.genome(SHAPES[shape]["G"], 42)
    pats = [p_ for p_ in synth.patterns(g, 40, min(10, SHAPES[shape]["G"] // 2), 7)]
    pre = FMI.from_rle(A.rle()).count(pats) + FMI.from_rle(B.rle()).count(pats)
    assert np.array_equal(M.count(pats), pre)
    assert np.array_equal(M.count(pats), np.array([want.count(p_) for p_ in pats], dtype=np.uint64))


def test_sequential_multi_merge(oracle):
    """bwt_merge.cpp:166-173: the merged index becomes the next A."""
    parts = [make_collection(oracle, 3000, n, 50, 0.01, 42, seed) for seed, n in ((1, 300), (2, 200), (3, 100), (4, 37))]
    want = oracle.from_comps(parts[0][1]); dev = FMI.from_rle(want.rle())
    for reads, bwt in parts[1:]:
        inc = oracle.from_comps(bwt)
        want = oracle.merge(want, inc)
        dev = FMI.merge(dev, FMI.from_rle(inc.rle()))
        assert np.array_equal(dev.rle(), want.rle())
    direct = oracle.bwt_of_reads([r for reads, _ in parts for r in reads])
    assert np.array_equal(dev.extract(), direct)


def _coalesce(runs):
    out = []
    for comp, length in runs:
        if out and out[-1][0] == comp:
            out[-1] = (comp, out[-1][1] + length)
        else:
            out.append((comp, length))
    return out


def test_encoder_offsets_and_long_runs(oracle):
    """Run::write edge cases through the device encoder (K3 + K5): long runs of every length class at
    arbitrary block offsets. Homopolymer reads give BWTs that are a few very long runs; mixing read
    lengths of 41..130 moves them over all offsets."""
    rng = np.random.default_rng(11)
    for trial in range(3):   # oracle self-check: split runs decode back to the maximal runs
        runs = []
        for _ in range(400):
            length = int(rng.integers(1, 42)) if rng.integers(0, 4) == 0 else int(rng.choice([42, 43, 83, 169, 170, 171, 300, 5000, 16425, 16426, 70000]))
            comp = int(rng.integers(1, 5))
            if runs and runs[-1][0] == comp:
                comp = comp % 4 + 1
            runs.append((comp, length))
        assert _coalesce(oracle.decode_runs(oracle.encode_runs(runs))) == runs
    for trial in range(8):
        runs = []
        for _ in range(int(rng.integers(1, 700))):
            length = int(rng.integers(1, 42)) if rng.integers(0, 3) else int(rng.choice([42, 43, 82, 83, 84, 169, 170, 171, 300, 5000, 16425, 16426, 40000]))
            comp = int(rng.integers(0, 6))
            if runs and runs[-1][0] == comp:
                comp = (comp + 1) % 6
            runs.append((comp, length))
        comps = np.repeat(np.array([r[0] for r in runs], np.uint8), [r[1] for r in runs])
        want = oracle.encode_runs(runs)
        for slab in (0, 4096, 12288):
            D = FMI.from_comps(comps, slab_symbols=slab)
            assert np.array_equal(D.rle(), want), (trial, slab)
            assert np.array_equal(D.counts(), np.bincount(comps, minlength=6).astype(np.uint64))
    assert np.array_equal(FMI.from_comps(np.full(100000, 3, np.uint8), slab_symbols=4096).rle(), oracle.encode_runs([(3, 100000)]))
    for m, L in ((50, 41), (97, 64), (300, 100), (1000, 130), (43, 42), (20000, 3)):
        for base in (1, 4):
            reads = np.full((m, L), base, dtype=np.uint8)
            reads[m // 2:, L // 2:] = 2          # two read types: runs of several lengths
            D = FMI.from_reads(reads)
            want = oracle.from_comps(oracle.bwt_of_reads([r for r in reads]))
            assert np.array_equal(D.rle(), want.rle()), (m, L, base)


@pytest.mark.parametrize("shape", ["reads", "noisy_N", "repeats"])
def test_device_builder(oracle, shape):
    s = SHAPES[shape]
    g = synth.genome(s["G"], 42)
    reads = synth.reads(g, s["n"], s["L"], s["e"], 5)
    want = oracle.from_comps(oracle.bwt_of_reads([r for r in reads]))
    assert np.array_equal(FMI.from_reads(reads).rle(), want.rle())
    D = FMI.synthetic(s["G"], 42, s["L"], synth.error_threshold(s["e"]), [(5, s["n"])])
    assert np.array_equal(D.rle(), want.rle())


def test_config1_shape_against_reference_binary(oracle, tmp_path):
    """Config 1 of BASELINE.json at 1/4 scale against the unmodified reference binary (oracle/_ref), and
    against a direct device construction of BWT(A ++ B)."""
    from oracle.oracle import REF_DIR, ref_available
    G, n, L, thr = 250000, 25000, 100, synth.error_threshold(0.01)
    A = FMI.synthetic(G, 42, L, thr, [(1, n)])
    B = FMI.synthetic(G, 42, L, thr, [(2, n)])
    AB = FMI.synthetic(G, 42, L, thr, [(1, n), (2, n)])
    rle_a, rle_b = A.rle(), B.rle()
    M = FMI.merge(A, B)
    assert np.array_equal(M.rle(), AB.rle())
    if not ref_available():
        pytest.skip("oracle/_ref not present")
    fa, fb, fm = (str(tmp_path / x) for x in ("A.plain", "B.plain", "M.plain"))
    synth.comps_to_chars(oracle.from_rle(rle_a).decode()).tofile(fa)
    synth.comps_to_chars(oracle.from_rle(rle_b).decode()).tofile(fb)
    subprocess.check_call([os.path.join(REF_DIR, "bwt_merge"), "-i", "plain_default", "-o", "plain_default", "-t", "4",
                           "-r", "8", "-d", str(tmp_path), fa, fb, fm], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref_chars = np.fromfile(fm, dtype=np.uint8)
    assert np.array_equal(synth.comps_to_chars(M.extract()), ref_chars)


def _variable_reads(rng, genome, n, lo, hi):
    reads = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1)); s = int(rng.integers(0, len(genome) - L + 1))
        reads.append(genome[s:s + L].copy())
    return reads


def test_variable_length_sequences(oracle):
    """Sequences of very different lengths (1..400): lanes of a warp finish at different times and refill."""
    rng = np.random.default_rng(5)
    g = synth.genome(3000, 42)
    ra = _variable_reads(rng, g, 300, 1, 400); rb = _variable_reads(rng, g, 257, 1, 400)
    A, B = oracle.from_comps(oracle.bwt_of_reads(ra)), oracle.from_comps(oracle.bwt_of_reads(rb))
    want = oracle.merge(A, B)
    DA, DB = FMI.from_rle(A.rle()), FMI.from_rle(B.rle())
    assert np.array_equal(bwtm_b200.rank_array(DA, DB), np.sort(oracle.build_ra_walk(A, B)))
    M = FMI.merge(DA, DB)
    assert np.array_equal(M.rle(), want.rle())
    assert np.array_equal(M.extract(), oracle.bwt_of_reads(ra + rb))


@pytest.mark.parametrize("na,nb", [(1, 400), (400, 1), (1, 1), (3, 2)])
def test_unbalanced_collections(oracle, na, nb):
    ra, bwt_a = make_collection(oracle, 2000, na, 80, 0.01, 42, 1)
    rb, bwt_b = make_collection(oracle, 2000, nb, 80, 0.01, 42, 2)
    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()))
    assert np.array_equal(M.rle(), oracle.merge(A, B).rle())


def test_wide_key_and_position_paths(oracle, monkeypatch):
    """The 64-bit key / position instantiations (used when a BWT has >= 2^32 symbols) on small inputs."""
    monkeypatch.setenv("BWTM_FORCE_WIDE", "1")
    for shape in ("reads", "noisy_N", "repeats"):
        ra, bwt_a, rb, bwt_b = collections(oracle, shape)
        A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
        p = MergeParameters(); p.slab_symbols = 4096
        M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
        assert np.array_equal(M.rle(), oracle.merge(A, B).rle()), shape


def test_superblock_boundary():
    """More than 2^25 symbols per input: several superblocks of the rank structure; merge == direct build."""
    G, n, L, thr = 3_000_000, 400_000, 100, synth.error_threshold(0.01)
    A = FMI.synthetic(G, 42, L, thr, [(1, n)]); B = FMI.synthetic(G, 42, L, thr, [(2, n // 2)])
    assert A.size() > (1 << 25)
    AB = FMI.synthetic(G, 42, L, thr, [(1, n), (2, n // 2)])
    pats = [p for p in synth.patterns(synth.genome(G, 42), 500, 24, 3)]
    pre = A.count(pats) + B.count(pats)
    M = FMI.merge(A, B)
    assert np.array_equal(M.rle(), AB.rle())
    assert np.array_equal(M.count(pats), pre) and pre.sum() > 0


def test_invalid_inputs_fail_loudly(oracle):
    with pytest.raises(bwtm_b200.BwtmError):
        FMI.from_rle(np.zeros(0, np.uint8))
    ra, bwt_a = make_collection(oracle, 500, 20, 30, 0.0, 42, 1)
    A = oracle.from_comps(bwt_a)
    with pytest.raises(bwtm_b200.BwtmError):          # wrong expected counts
        FMI.from_rle(A.rle(), expected_counts=A.counts() + np.uint64(1))
    # a "BWT" without endmarkers cannot be inserted
    with pytest.raises(bwtm_b200.BwtmError):
        FMI.merge(FMI.from_rle(A.rle()), FMI.from_comps(np.full(100, 2, np.uint8)))
    # a sequence whose walks do not cover it (not a BWT of a collection) is rejected, not mis-merged
    bogus = np.concatenate([np.zeros(3, np.uint8), np.full(50, 1, np.uint8), np.full(50, 2, np.uint8)])
    with pytest.raises(bwtm_b200.BwtmError):
        FMI.merge(FMI.from_rle(A.rle()), FMI.from_comps(bogus))


@pytest.mark.parametrize("slab", [0, 4096])
@pytest.mark.parametrize("batches", [2, 3, 5, 7])
def test_search_in_batches(oracle, batches, slab):
    """options.sequence_blocks > 1 (the counterpart of the reference's bounded buffers, fmi.cpp:164-257): b's sequences
    are searched in batches kept as sorted runs, and the interleave gathers and merges the pieces of every range of A
    positions. Same bytes as the one-shot merge, including the reference's RA run count."""
    for shape in ("reads", "noisy_N", "repeats", "single"):
        ra, bwt_a, rb, bwt_b = collections(oracle, shape)
        A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
        p = MergeParameters(); p.sequence_blocks = batches; p.slab_symbols = slab
        M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
        want = oracle.merge(A, B)
        assert np.array_equal(M.rle(), want.rle()), (shape, batches)
        assert M.timings.search_batches == min(batches, B.sequences)
        assert M.timings.ra_runs == len(oracle.sort_compress(oracle.build_ra_dfs(A, B)))
        assert M.hash() == want.hash()
    rng = np.random.default_rng(9)
    g = synth.genome(2000, 42)
    ra = _variable_reads(rng, g, 200, 1, 300); rb = _variable_reads(rng, g, 150, 1, 300)
    A, B = oracle.from_comps(oracle.bwt_of_reads(ra)), oracle.from_comps(oracle.bwt_of_reads(rb))
    p = MergeParameters(); p.sequence_blocks = batches; p.slab_symbols = slab
    out = np.zeros(A.bytes + B.bytes + 4096, dtype=np.uint8); p.host_output = out
    M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
    want = oracle.merge(A, B).rle()
    assert np.array_equal(M.rle(), want) and np.array_equal(out[:len(want)], want)


@pytest.mark.parametrize("shape", list(SHAPES))
def test_pair_records_answer_two_steps(oracle, shape):
    """bwtm_lf2 (the pair records of bwtm_pairs.cu) equals two applications of FMI::LF(i) (fmi.h:147-150), and both
    equal the reference's inverse_select restated in the oracle."""
    ra, bwt_a, rb, bwt_b = collections(oracle, shape)
    A = oracle.from_comps(bwt_a)
    D = FMI.from_rle(A.rle())
    n = A.size
    pos = np.arange(n, dtype=np.uint64) if n <= 40000 else np.random.default_rng(5).integers(0, n, 40000).astype(np.uint64)
    first, second, c1, c2 = D.LF2(pos)
    one, comp1 = D.LF(pos)
    assert np.array_equal(c1, comp1)
    live = comp1 != 0
    assert np.array_equal(first[live], one[live]) and not first[~live].any() and not second[~live].any() and not c2[~live].any()
    two, comp2 = D.LF(one[live])
    assert np.array_equal(c2[live], comp2)
    live2 = comp2 != 0
    assert np.array_equal(second[live][live2], two[live2]) and not second[live][~live2].any()
    C_ = A.C()
    for i in pos[:300]:
        r, c = A.inverse_select(int(i))
        k = int(np.nonzero(pos == i)[0][0])
        assert c == c1[k] and (c == 0 or first[k] == C_[c] + r)


@pytest.mark.parametrize("walk", ["single", "pairs"])
@pytest.mark.parametrize("wide", [False, True])
def test_both_walks_give_the_reference_rank_array(oracle, monkeypatch, walk, wide):
    """The single-step walk (64-byte records) and the two-step walk (128-byte pair records) emit the multiset that
    buildRA (fmi.cpp:272-334) emits, with 32- and 64-bit positions, and merges through either are byte-identical."""
    monkeypatch.setenv("BWTM_WALK", walk)
    if wide:
        monkeypatch.setenv("BWTM_FORCE_WIDE", "1")
    for shape in SHAPES:
        ra, bwt_a, rb, bwt_b = collections(oracle, shape)
        A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
        DA, DB = FMI.from_rle(A.rle()), FMI.from_rle(B.rle())
        runs = oracle.build_ra_dfs(A, B)
        want = np.sort(np.repeat(runs[:, 0], runs[:, 1].astype(np.int64)))
        assert np.array_equal(bwtm_b200.rank_array(DA, DB), want), (shape, walk)
        M = FMI.merge(DA, DB)
        assert M.timings.walk_record_bytes == (128 if walk == "pairs" else 64)
        assert np.array_equal(M.rle(), oracle.merge(A, B).rle()), (shape, walk)
    # sequences of odd and even lengths, length 0 included: the two-step walk stops after either step
    rng = np.random.default_rng(21)
    g = synth.genome(1500, 42)
    ra = _variable_reads(rng, g, 120, 1, 90); rb = _variable_reads(rng, g, 200, 1, 7)
    A, B = oracle.from_comps(oracle.bwt_of_reads(ra)), oracle.from_comps(oracle.bwt_of_reads(rb))
    assert np.array_equal(FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle())).rle(), oracle.merge(A, B).rle())


def test_pair_records_across_superblocks():
    """More than 2^20 positions (pair superblocks) and a sequential merge whose inputs keep or lack pair records."""
    thr = synth.error_threshold(0.02)
    A = FMI.synthetic(300_000, 42, 100, thr, [(1, 30_000)])      # 3.03 M symbols: three pair superblocks
    B = FMI.synthetic(300_000, 42, 100, thr, [(2, 20_000)])
    pos = np.random.default_rng(2).integers(0, A.size(), 50_000).astype(np.uint64)
    first, second, c1, c2 = A.LF2(pos)
    one, comp1 = A.LF(pos)
    live = comp1 != 0
    two, comp2 = A.LF(one[live])
    assert np.array_equal(c1, comp1) and np.array_equal(first[live], one[live]) and np.array_equal(c2[live], comp2)
    assert np.array_equal(second[live][comp2 != 0], two[comp2 != 0])
    os.environ["BWTM_WALK"] = "single"
    try:
        want = FMI.merge(A, B, keep_inputs=True).rle()
    finally:
        del os.environ["BWTM_WALK"]
    os.environ["BWTM_WALK"] = "pairs"
    try:
        got = FMI.merge(A, B, keep_inputs=True)
        assert got.timings.walk_record_bytes == 128
    finally:
        del os.environ["BWTM_WALK"]
    assert np.array_equal(got.rle(), want)
    direct = FMI.synthetic(300_000, 42, 100, thr, [(1, 30_000), (2, 20_000)])
    assert np.array_equal(direct.rle(), want)


@pytest.mark.parametrize("slab", [0, 4096])
def test_streaming_download(oracle, slab):
    """options.host_output: the merged bytes arrive in the host buffer while later slabs are encoded."""
    ra, bwt_a, rb, bwt_b = collections(oracle, "reads")
    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    want = oracle.merge(A, B).rle()
    p = MergeParameters(); p.slab_symbols = slab; p.host_output = np.zeros(len(want) + 100, dtype=np.uint8)
    M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
    assert M.timings.merged_bytes == len(want)
    assert np.array_equal(p.host_output[:len(want)], want) and np.array_equal(M.rle(), want)
    p.host_output = np.zeros(len(want) - 1, dtype=np.uint8)      # too small: refused, not truncated
    with pytest.raises(bwtm_b200.BwtmError) as err:
        FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
    assert err.value.code == -5


@pytest.mark.parametrize("slab", [0, 4096, 12288])
def test_result_index_from_merged_symbols(oracle, monkeypatch, slab):
    """The result's rank structure is filled from the merged symbols; it must equal the one decoded from the
    encoded bytes (BWTM_RLE_INDEX=1) and answer like the reference's BWT::rank / FMI::LF (bwt.cpp:318-341)."""
    for shape in ("reads", "noisy_N", "repeats"):
        ra, bwt_a, rb, bwt_b = collections(oracle, shape)
        A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
        want = oracle.merge(A, B)
        p = MergeParameters(); p.slab_symbols = slab
        fast = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
        monkeypatch.setenv("BWTM_RLE_INDEX", "1")
        slow = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()), p)
        monkeypatch.delenv("BWTM_RLE_INDEX")
        n = fast.size()
        assert n == slow.size() == want.size and fast.sequences() == slow.sequences()
        positions = np.unique(np.concatenate([np.arange(min(n + 1, 300)), np.arange(max(0, n - 300), n + 1),
                                              np.random.default_rng(3).integers(0, n + 1, 2000)])).astype(np.uint64)
        for c in range(6):
            comps = np.full(len(positions), c, dtype=np.uint8)
            got = fast.rank(positions, comps)
            assert np.array_equal(got, slow.rank(positions, comps)), (shape, slab, c)
            assert np.array_equal(got, np.array([want.rank(int(i), c) for i in positions], dtype=np.uint64)), (shape, c)
        inside = positions[positions < n]
        assert all(np.array_equal(x, y) for x, y in zip(fast.LF(inside), slow.LF(inside)))
        assert np.array_equal(fast.extract(0, n), want.decode())
        assert fast.hash() == slow.hash() == want.hash()


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("limit", [None, "1", "64", "1000", "wide-counters"])
def test_counting_sort_of_low_bits(oracle, monkeypatch, wide, limit):
    """K2 for large inputs: radix sort on the high bits, then one counting pass per range of A positions
    (forced on small inputs here). Ranges holding more keys than the limit are radix-sorted on their own
    (limits 64 and 1000 make some ranges heavy); more than 64 such ranges (limit 1) fall back to the plain sort."""
    monkeypatch.setenv("BWTM_LOCAL_SORT_MIN", "1"); monkeypatch.setenv("BWTM_LOCAL_SORT_DENSITY", "0")
    monkeypatch.setenv("BWTM_LOCAL_SORT_MAX_DENSITY", "0")
    if limit == "wide-counters":      # the kernel for ranges of more than 65535 keys, on all ranges
        monkeypatch.setenv("BWTM_LOCAL_SORT_SMALL", "0")
    elif limit is not None:
        monkeypatch.setenv("BWTM_LOCAL_SORT_LIMIT", limit)
    if wide:
        monkeypatch.setenv("BWTM_FORCE_WIDE", "1")
    for shape in ("reads", "noisy_N", "repeats", "single"):
        ra, bwt_a, rb, bwt_b = collections(oracle, shape)
        A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
        DA, DB = FMI.from_rle(A.rle()), FMI.from_rle(B.rle())
        assert np.array_equal(bwtm_b200.rank_array(DA, DB), np.sort(oracle.build_ra_walk(A, B))), shape
        M = FMI.merge(DA, DB)
        assert np.array_equal(M.rle(), oracle.merge(A, B).rle()), shape
    # an A of 70000 symbols: 17-bit keys, 12 counted bits, 32 ranges
    big_a = make_collection(oracle, 20000, 1150, 60, 0.01, 42, 5)[1]; small_b = make_collection(oracle, 20000, 300, 60, 0.02, 42, 6)[1]
    A, B = oracle.from_comps(big_a), oracle.from_comps(small_b)
    assert A.size > 65536
    M = FMI.merge(FMI.from_rle(A.rle()), FMI.from_rle(B.rle()))
    assert np.array_equal(M.rle(), oracle.merge(A, B).rle())


def test_create_pair(oracle):
    """bwtm_index_create_pair = two bwtm_index_create calls (the second upload overlaps the first K0)."""
    ra, bwt_a, rb, bwt_b = collections(oracle, "reads")
    A, B = oracle.from_comps(bwt_a), oracle.from_comps(bwt_b)
    DA, DB = FMI.from_rle_pair(A.rle(), B.rle())
    for D, O in ((DA, A), (DB, B)):
        assert D.size() == O.size and D.sequences() == O.sequences and np.array_equal(D.counts(), O.counts())
        assert np.array_equal(D.rle(), O.rle()) and np.array_equal(D.extract(), O.decode()) and D.hash() == O.hash()
    assert np.array_equal(FMI.merge(DA, DB).rle(), oracle.merge(A, B).rle())
    with pytest.raises(bwtm_b200.BwtmError):
        FMI.from_rle_pair(A.rle(), np.zeros(0, dtype=np.uint8))


def _run_bytes(comps, layout, rng=None):
    """One (comp, length <= 31) run per byte; with rng the maximal runs are cut at random places as well."""
    starts = np.concatenate([[0], np.flatnonzero(np.diff(comps)) + 1]); lengths = np.diff(np.concatenate([starts, [len(comps)]]))
    out = []
    for c, n in zip(comps[starts].tolist(), lengths.tolist()):
        while n > 0:
            take = min(31, n) if rng is None else int(rng.integers(1, min(31, n) + 1))
            out.append((take << 3 | c) if layout == FMI.RUNS_ROPEBWT else (c << 5 | take)); n -= take
    return np.array(out, dtype=np.uint8)


@pytest.mark.parametrize("layout", [FMI.RUNS_ROPEBWT, FMI.RUNS_SGA])
def test_run_byte_transcoder(oracle, tmp_path, layout):
    """bwtm_index_create_runs: RopeData::read / SGAData::read on the device (formats.cpp:286-310, 403-429). Pieces of one
    symbol are joined into maximal runs, as the reference's RunBuffer does, whatever way the file cut them."""
    rng = np.random.default_rng(5)
    for shape in ("reads", "noisy_N", "repeats", "single"):
        bwt = collections(oracle, shape)[1]
        if shape == "reads":   # long runs of every class as well
            bwt = np.concatenate([bwt[:3000], np.full(70000, 2, np.uint8), bwt[3000:], np.full(43, 5, np.uint8), np.full(200, 0, np.uint8)])
        want = oracle.from_comps(bwt)
        for cut, slab in ((None, 0), (rng, 0), (rng, 4096)):
            D = FMI.from_run_bytes(_run_bytes(bwt, layout, cut), layout, slab)
            assert np.array_equal(D.rle(), want.rle()), (shape, slab)
            assert np.array_equal(D.counts(), want.counts()) and D.size() == want.size and D.hash() == want.hash()
        assert np.array_equal(D.extract(), bwt)
    # a file written by the reference's own bwt_convert, when it is here
    from oracle.oracle import REF_DIR, ref_available
    if ref_available():
        bwt = collections(oracle, "noisy_N")[1]
        src, dst = str(tmp_path / "src.plain"), str(tmp_path / "dst.runs")
        synth.comps_to_chars(bwt).tofile(src)
        fmt = "ropebwt" if layout == FMI.RUNS_ROPEBWT else "sga"
        subprocess.check_call([os.path.join(REF_DIR, "bwt_convert"), "-i", "plain_default", "-o", fmt, src, dst],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        data = np.fromfile(dst, dtype=np.uint8)
        # 4-byte tag / 30-byte SGA header (tag u16, sequences, bases, bytes u64, flags u32) that states the number of run bytes
        body = data[4:] if layout == FMI.RUNS_ROPEBWT else data[30:30 + int(data[18:26].view(np.uint64)[0])]
        assert np.array_equal(FMI.from_run_bytes(body, layout).rle(), oracle.from_comps(bwt).rle())
    # refused: a zero-length run, a comp value of 6
    bad = np.array([(3 << 3) | 1, 0 | 2, (6 << 3) | 6], dtype=np.uint8) if layout == FMI.RUNS_ROPEBWT else np.array([(1 << 5) | 3, (2 << 5), (6 << 5) | 6], dtype=np.uint8)
    with pytest.raises(bwtm_b200.BwtmError) as err:
        FMI.from_run_bytes(bad, layout)
    assert err.value.code == -4
