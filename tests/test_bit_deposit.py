"""The bit-deposit network used by the pair-record builder (pairs_gather in csrc/bwtm_pairs.cu: deposit_network /
deposit, Hacker's Delight 7-5 "expand"), restated with 32-bit integer arithmetic and checked against the definition: the
low popcount(mask) bits of x go to the set positions of mask, in order. (The device version is exercised by the GPU
parity tests of the pair records; this pins the algorithm.)"""
import numpy as np

M32 = 0xFFFFFFFF


def deposit_network(mask):
    moves, m = [], mask
    mk = (~mask << 1) & M32
    for i in range(5):
        mp = mk ^ ((mk << 1) & M32)
        for shift in (2, 4, 8, 16):
            mp ^= (mp << shift) & M32
        mv = mp & m
        moves.append(mv)
        m = ((m ^ mv) | (mv >> (1 << i))) & M32
        mk &= ~mp & M32
    return moves


def deposit(moves, mask, x):
    for i in range(4, -1, -1):
        mv = moves[i]
        x = ((x & ~mv) | ((x << (1 << i)) & mv)) & M32
    return x & mask


def definition(x, mask):
    out, k = 0, 0
    for t in range(32):
        if (mask >> t) & 1:
            out |= ((x >> k) & 1) << t
            k += 1
    return out


def test_deposit_network_matches_the_definition():
    rng = np.random.default_rng(7)
    masks = [0, M32, 1, 0x80000000, 0x55555555, 0xAAAAAAAA, 0x0000FFFF, 0xFFFF0000] + [int(v) for v in rng.integers(0, 1 << 32, 3000, dtype=np.uint64)]
    masks += [int(a & b) for a, b in zip(rng.integers(0, 1 << 32, 1000, dtype=np.uint64), rng.integers(0, 1 << 32, 1000, dtype=np.uint64))]
    for mask in masks:
        moves = deposit_network(mask)
        for x in (0, M32, 0x12345678, int(rng.integers(0, 1 << 32, dtype=np.uint64))):
            assert deposit(moves, mask, x) == definition(x, mask), (hex(mask), hex(x))
