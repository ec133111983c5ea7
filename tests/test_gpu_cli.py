"""Two command lines above the C ABI against the unmodified reference binary oracle/_ref/bwt_merge (same files,
same -v report):
  * bin/bwt_merge_b200          -- the product's host driver;
  * oracle/_ref/bwt_merge_b200  -- the REFERENCE's own bwt_merge.cpp and libraries, unmodified, with only the merging
                                   constructor FMI::FMI(FMI&, FMI&, MergeParameters) bound to libbwtm_b200.so by
                                   bwt-merge_b200/integration/fmi_b200.cpp (oracle/Makefile, target ref_b200)."""
import filecmp
import os
import re
import subprocess

import numpy as np
import pytest

from bwtm_b200 import synth
from conftest import ROOT, make_collection

pytestmark = pytest.mark.gpu
MINE = os.path.join(ROOT, "bwt-merge_b200", "bin", "bwt_merge_b200")
REF_OVER_ABI = os.path.join(ROOT, "oracle", "_ref", "bwt_merge_b200")
TOOLS = {"host_driver": MINE, "reference_cli_over_abi": REF_OVER_ABI}


def report(stdout):
    """The parts of the report that do not depend on timing."""
    keep = []
    for line in stdout.splitlines():
        m = re.match(r"(Input|Output):\s+Found (\d+) patterns with (\d+) occ", line)
        if m:
            keep.append(m.groups())
        m = re.match(r"(Input|Output):\s+([0-9.e+-]+) MB \(([0-9.e+-]+) bpc\)", line)
        if m:
            keep.append(m.groups())
        if line.startswith(("Verification", "Read ", "Input: ", "Output: ", "Patterns:")) and "Found" not in line and " MB (" not in line:
            keep.append(line)
    return keep


@pytest.mark.parametrize("tool", list(TOOLS))
@pytest.mark.parametrize("fmt_in,fmt_out", [("plain_default", "native"), ("native", "plain_default"), ("sga", "ropebwt"), ("ropebwt", "sga")])
def test_cli_matches_reference(oracle, tmp_path, fmt_in, fmt_out, tool):
    from oracle.oracle import REF_DIR, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not present")
    under_test = TOOLS[tool]
    if not os.path.exists(under_test):
        pytest.skip("%s not built" % under_test)
    ref_merge, ref_convert = os.path.join(REF_DIR, "bwt_merge"), os.path.join(REF_DIR, "bwt_convert")
    inputs = []
    for k, n in enumerate((500, 300, 120)):
        reads, bwt = make_collection(oracle, 4000, n, 70, 0.01, 42, k + 1, 0.01)
        plain = str(tmp_path / ("in%d.plain" % k)); synth.comps_to_chars(bwt).tofile(plain)
        path = str(tmp_path / ("in%d.%s" % (k, fmt_in)))
        subprocess.check_call([ref_convert, "-i", "plain_default", "-o", fmt_in, plain, path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        inputs.append(path)
    g = synth.genome(4000, 42)
    patterns = str(tmp_path / "patterns.txt")
    with open(patterns, "w") as f:
        for p in synth.patterns(g, 200, 14, 9):
            f.write(synth.comps_to_chars(p).tobytes().decode() + "\n")
        f.write("\nNNNNNN\nACGT\n")
    outs = []
    for binary, name in ((under_test, "mine"), (ref_merge, "ref")):
        out = str(tmp_path / (name + "." + fmt_out))
        res = subprocess.run([binary, "-t", "4", "-r", "1", "-b", "1", "-d", str(tmp_path), "-v", patterns, "-i", fmt_in, "-o", fmt_out]
                             + inputs + [out], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append((out, res.stdout.replace(out, "OUTPUT")))
    assert filecmp.cmp(outs[0][0], outs[1][0], shallow=False)
    assert report(outs[0][1]) == report(outs[1][1])
    assert "Verification successful" in outs[0][1]


def test_cli_rejects_different_alphabets(oracle, tmp_path):
    reads, bwt = make_collection(oracle, 1000, 50, 40, 0.0, 42, 1)
    a, b = str(tmp_path / "a.plain"), str(tmp_path / "b.plain")
    synth.comps_to_chars(bwt).tofile(a); np.frombuffer(b"$ACGNT", dtype=np.uint8)[bwt].tofile(b)
    res = subprocess.run([MINE, "-i", "plain_default,plain_sorted", "-o", "plain_default", a, b, str(tmp_path / "out")],
                         capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and "Cannot merge BWTs with different alphabets" in res.stderr   # fmi.cpp:338-342
