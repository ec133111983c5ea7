"""CPU checks of the C-ABI library: it builds, loads and exports every symbol include/bwtm.h declares,
and it fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import bwtm_b200
from conftest import ROOT


@pytest.fixture(scope="module")
def library():
    bwtm_b200.build_library()
    return bwtm_b200.lib()


def test_header_symbols_are_exported(library):
    header = open(os.path.join(ROOT, "include", "bwtm.h")).read()
    declared = set(re.findall(r"\b(bwtm_[a-z_0-9]+)\s*\(", header))
    assert declared == set(bwtm_b200.EXPORTS)
    for name in declared:
        assert hasattr(library, name), name


def test_version_and_error_string(library):
    assert b"sm_100a" in library.bwtm_version()
    assert isinstance(library.bwtm_last_error(), bytes)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(bwtm_b200.IndexInfo) == 8 * (3 + 6 + 7 + 1)
    assert ctypes.sizeof(bwtm_b200.MergeOptions) == 8 * 6 + 8 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(bwtm_b200.Timings) == 8 * 13


def test_no_cpu_fallback(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rle = np.array([6 * 3 + 1, 0], dtype=np.uint8)
    with pytest.raises(bwtm_b200.BwtmError) as err:
        bwtm_b200.FMI.from_rle(rle)
    assert err.value.code == -2
    assert "no CPU fallback" in str(err.value)
