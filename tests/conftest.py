import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bwt-merge_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build(ref=True)
    return Oracle()


@pytest.fixture(scope="session")
def refhooks(oracle):
    from oracle.oracle import RefHooks, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return RefHooks()


def make_collection(oracle, genome_len, n_reads, read_len, error, gseed, rseed, n_frac=0.0):
    """Synthetic reads -> (reads matrix, BWT comps) via the oracle's suffix sorter."""
    from bwtm_b200 import synth
    g = synth.genome(genome_len, gseed)
    r = synth.reads(g, n_reads, read_len, error, rseed)
    if n_frac > 0:
        rng = np.random.default_rng(rseed)
        r = r.copy(); r[rng.random(r.shape) < n_frac] = 5
    bwt = oracle.bwt_of_reads([row for row in r])
    return r, bwt
