"""Host-side file formats (bwt-merge_b200/host) against the reference's own bwt_convert (oracle/_ref):
every format is written byte-identically and read back identically. CPU only."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import bwtm_b200
from bwtm_b200 import synth
from conftest import ROOT, make_collection

FORMATS = ["native", "plain_default", "plain_sorted", "rfm", "sdsl", "ropebwt", "sga"]
DEFAULT_ORDER = ["native", "plain_default", "ropebwt", "sga"]
SORTED_ORDER = ["native", "plain_sorted", "rfm", "sdsl"]
MINE = os.path.join(ROOT, "bwt-merge_b200", "bin", "bwt_convert_b200")


@pytest.fixture(scope="module")
def tools():
    bwtm_b200.build_library()
    from oracle.oracle import REF_DIR, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    return MINE, os.path.join(REF_DIR, "bwt_convert")


def run(tool, src, dst, fi, fo):
    subprocess.check_call([tool, "-i", fi, "-o", fo, src, dst], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.mark.parametrize("family", ["default", "sorted"])
def test_formats_match_reference(oracle, tools, tmp_path, family):
    mine, ref = tools
    reads, bwt = make_collection(oracle, 1500, 400, 60, 0.02, 42, 3, 0.02)
    # long runs as well: a homopolymer block in the middle
    bwt = np.concatenate([bwt[:5000], np.full(70000, 2, np.uint8), bwt[5000:], np.full(43, 5, np.uint8)])
    plain_tag = "plain_default" if family == "default" else "plain_sorted"
    chars = synth.comps_to_chars(bwt) if family == "default" else np.frombuffer(b"$ACGNT", dtype=np.uint8)[bwt]
    src = str(tmp_path / "src.plain"); chars.tofile(src)
    formats = DEFAULT_ORDER if family == "default" else SORTED_ORDER
    for fmt in formats:
        a, b = str(tmp_path / ("mine." + fmt)), str(tmp_path / ("ref." + fmt))
        run(mine, src, a, plain_tag, fmt); run(ref, src, b, plain_tag, fmt)
        assert filecmp.cmp(a, b, shallow=False), fmt
        # read each other's files back to plain
        back_a, back_b = str(tmp_path / "back_a"), str(tmp_path / "back_b")
        run(mine, b, back_a, fmt, plain_tag); run(ref, a, back_b, fmt, plain_tag)
        assert filecmp.cmp(back_a, src, shallow=False) and filecmp.cmp(back_b, src, shallow=False), fmt
        # and into every other format of the family, through both tools
        for other in formats:
            c, d = str(tmp_path / "x"), str(tmp_path / "y")
            run(mine, a, c, fmt, other); run(ref, a, d, fmt, other)
            assert filecmp.cmp(c, d, shallow=False), (fmt, other)


def test_plain_reader_quirks(oracle, tools, tmp_path):
    """PlainData::read coalesces equal CHARACTERS, not equal comps: 'Aa' stays two runs (formats.cpp:145-152)."""
    mine, ref = tools
    src = str(tmp_path / "mixed.plain")
    open(src, "wb").write(b"AAAaaa$$\x00\x00CcCcXYZNNnn" * 50 + b"T" * 100)
    a, b = str(tmp_path / "a.native"), str(tmp_path / "b.native")
    run(mine, src, a, "plain_default", "native"); run(ref, src, b, "plain_default", "native")
    assert filecmp.cmp(a, b, shallow=False)


def test_native_loader_does_not_depend_on_the_sparse_vector_layout(oracle, tools, tmp_path):
    """The seven sparse vectors of a native file are SDSL's to lay out; the loader reads the header,
    the run-length bytes and the alphabet (the last 352 bytes) and skips whatever lies between. A file whose sample
    section has another size (as a real-SDSL build would write it) must load; malformed files must fail with a message."""
    mine, ref = tools
    reads, bwt = make_collection(oracle, 1500, 300, 50, 0.02, 42, 7)
    src = str(tmp_path / "src.plain"); synth.comps_to_chars(bwt).tofile(src)
    native = str(tmp_path / "a.native"); run(ref, src, native, "plain_default", "native")
    data = open(native, "rb").read()
    rle_bytes = int(np.frombuffer(data[24:32], dtype=np.uint64)[0])
    payload_end = 32 + ((rle_bytes + (8 << 20) - 1) // (8 << 20)) * (8 << 20)
    foreign = str(tmp_path / "foreign.native")
    open(foreign, "wb").write(data[:payload_end] + bytes(np.random.default_rng(1).integers(0, 256, 12345, dtype=np.uint8)) + data[-352:])
    back = str(tmp_path / "back.plain"); run(mine, foreign, back, "native", "plain_default")
    assert filecmp.cmp(back, src, shallow=False)
    for name, blob in (("truncated", data[:payload_end - 100]), ("no_alphabet", data[:payload_end] + b"\0" * 400)):
        bad = str(tmp_path / (name + ".native")); open(bad, "wb").write(blob)
        res = subprocess.run([mine, "-i", "native", "-o", "plain_default", bad, str(tmp_path / "out")], capture_output=True, text=True)
        assert res.returncode != 0 and "BWT::load()" in res.stderr, name
