"""Synthetic read collections, host/numpy twin of the CUDA generator in
csrc/bwtm_tools.cu.  Counter-based (splitmix64), so any element can be produced independently
and the host and device versions agree bit for bit.

comp values: 0 = $, 1 = A, 2 = C, 3 = G, 4 = T, 5 = N  (support.cpp:40-63 in the reference).
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
GOLDEN = 0x9E3779B97F4A7C15
STREAM = 0xD1342543DE82EF95


def _mix(x):
    """splitmix64 output function on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(GOLDEN)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rnd(seed, stream, index):
    """rnd(seed, stream, i) = mix(mix(seed + stream * STREAM) + i)."""
    with np.errstate(over="ignore"):
        base = _mix(np.array([(seed + stream * STREAM) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        return _mix(np.asarray(index, dtype=np.uint64) + base)


def error_threshold(error_rate):
    """Substitute a base iff (rnd >> 11) < threshold; threshold = floor(e * 2^53)."""
    return int(error_rate * (1 << 53))


def genome(length, seed):
    """Uniform i.i.d. genome over ACGT as comp values 1..4."""
    return ((rnd(seed, 0, np.arange(length, dtype=np.uint64)) & np.uint64(3)) + np.uint64(1)).astype(np.uint8)


def reads(genome_comps, n_reads, read_len, error_rate, seed, first_read=0):
    """n_reads x read_len comp matrix: forward-strand reads at uniform starts with i.i.d. substitutions."""
    G = len(genome_comps)
    ids = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    starts = rnd(seed, 1, ids) % np.uint64(G - read_len + 1)
    idx = starts[:, None] + np.arange(read_len, dtype=np.uint64)[None, :]
    base = genome_comps[idx.astype(np.int64)].astype(np.uint64) - np.uint64(1)
    cell = ids[:, None] * np.uint64(read_len) + np.arange(read_len, dtype=np.uint64)[None, :]
    u = rnd(seed, 2, cell) >> np.uint64(11)
    sub = u < np.uint64(error_threshold(error_rate))
    shift = (rnd(seed, 3, cell) % np.uint64(3)) + np.uint64(1)
    base = np.where(sub, (base + shift) & np.uint64(3), base)
    return (base + np.uint64(1)).astype(np.uint8)


def patterns(genome_comps, n_patterns, pattern_len, seed):
    """pattern_len-mers at uniform genome offsets (comp values)."""
    G = len(genome_comps)
    starts = rnd(seed, 4, np.arange(n_patterns, dtype=np.uint64)) % np.uint64(G - pattern_len + 1)
    idx = starts[:, None] + np.arange(pattern_len, dtype=np.uint64)[None, :]
    return genome_comps[idx.astype(np.int64)]


COMP2CHAR = np.frombuffer(b"$ACGTN", dtype=np.uint8)


def comps_to_chars(comps):
    return COMP2CHAR[np.asarray(comps, dtype=np.uint8)]
