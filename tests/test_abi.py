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
    assert ctypes.sizeof(bwtm_b200.Timings) == 8 * 17


def test_no_cpu_fallback(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rle = np.array([6 * 3 + 1, 0], dtype=np.uint8)
    with pytest.raises(bwtm_b200.BwtmError) as err:
        bwtm_b200.FMI.from_rle(rle)
    assert err.value.code == -2
    assert "no CPU fallback" in str(err.value)


def test_reference_cli_links_the_abi_and_has_no_cpu_fallback(library, oracle, tmp_path):
    """oracle/_ref/bwt_merge_b200 is the reference's own bwt_merge.cpp with FMI::FMI(FMI&, FMI&, MergeParameters)
    (fmi.cpp:336-369) replaced by bwt-merge_b200/integration/fmi_b200.cpp: it must import the C ABI and, without a
    device, stop with the library's error instead of merging on the CPU."""
    import subprocess
    import torch
    from bwtm_b200 import synth
    from conftest import make_collection
    tool = os.path.join(ROOT, "oracle", "_ref", "bwt_merge_b200")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/bwt_merge_b200 not built (reference sources absent)")
    symbols = subprocess.run(["nm", "-D", "--undefined-only", tool], capture_output=True, text=True).stdout
    for name in ("bwtm_index_create_pair", "bwtm_merge", "bwtm_index_download"):
        assert name in symbols
    if torch.cuda.is_available():
        return
    paths = []
    for k in (1, 2):
        _, bwt = make_collection(oracle, 2000, 60, 40, 0.01, 42, k)
        paths.append(str(tmp_path / ("in%d.plain" % k))); synth.comps_to_chars(bwt).tofile(paths[-1])
    res = subprocess.run([tool, "-i", "plain_default", "-o", "plain_default"] + paths + [str(tmp_path / "out")],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr
