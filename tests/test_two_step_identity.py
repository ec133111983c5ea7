"""The identity behind the pair records (csrc/bwtm_pairs.cu), checked on the CPU against the reference-pinned oracle:

    LF(LF(i, c1), c2) = C[c2] + rank(C[c1], c2) + pairrank(i, c1 c2),
    pairrank(i, c1 c2) = #{k < i : BWT[k] = c1 and BWT[LF(k)] = c2},

for EVERY position i and every pair (c1, c2) of non-endmarker symbols -- the rank-array search applies it with (c1, c2) read
from B and i a position of A, so it must hold for pairs that do not occur at i as well. The right-hand side is what a
128-byte pair record + one superblock row deliver; the left-hand side is two applications of FMI::LF(i, c) (fmi.h:152-155)."""
import numpy as np
import pytest

from conftest import make_collection


@pytest.mark.parametrize("shape", [dict(G=3000, n=200, L=40, e=0.02, nfrac=0.02), dict(G=60, n=150, L=25, e=0.0, nfrac=0.0)])
def test_two_step_lf_identity(oracle, shape):
    _, bwt = make_collection(oracle, shape["G"], shape["n"], shape["L"], shape["e"], 42, 3, shape["nfrac"])
    A = oracle.from_comps(bwt)
    n = A.size
    C = A.C().astype(np.int64)
    seq = A.decode().astype(np.int64)
    # prefix counts: occ[c][i] = rank(i, c)
    occ = np.zeros((6, n + 1), dtype=np.int64)
    for c in range(6):
        occ[c, 1:] = np.cumsum(seq == c)
    for i in (0, 1, n // 3, n - 1, n):          # the prefix table is the oracle's rank
        for c in range(1, 6):
            assert occ[c, i] == A.rank(i, c)
    lf = np.where(seq > 0, C[seq] + occ[seq, np.arange(n)], 0)           # LF(k) for BWT[k] != $
    second = np.where(seq > 0, seq[np.minimum(lf, n - 1)], 0)            # BWT[LF(k)]
    positions = np.arange(n + 1)
    for c1 in range(1, 6):
        for c2 in range(1, 6):
            pair = np.zeros(n + 1, dtype=np.int64)
            pair[1:] = np.cumsum((seq == c1) & (second == c2))
            first = C[c1] + occ[c1, positions]                           # LF(i, c1), i = 0 .. n
            left = C[c2] + occ[c2, first]                                # LF(LF(i, c1), c2)
            right = C[c2] + occ[c2, C[c1]] + pair
            assert np.array_equal(left, right), (c1, c2)
