"""Command-line behaviour of bin/bwt_merge_b200 that needs no GPU, beside the unmodified reference binary: usage text,
option errors and their messages (bwt_merge.cpp:47-142). Everything past option parsing needs a device and is covered by
tests/test_gpu_cli.py."""
import os
import subprocess

import pytest

import bwtm_b200
from conftest import ROOT

MINE = os.path.join(ROOT, "bwt-merge_b200", "bin", "bwt_merge_b200")


@pytest.fixture(scope="module")
def tools():
    bwtm_b200.build_library()
    from oracle.oracle import REF_DIR, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    return MINE, os.path.join(REF_DIR, "bwt_merge")


def run(tool, *args):
    return subprocess.run([tool] + list(args), capture_output=True, text=True, timeout=60)


def test_usage_is_the_reference_text(tools):
    mine, ref = run(tools[0]), run(tools[1])
    assert mine.returncode == ref.returncode == 0
    assert mine.stderr == ref.stderr and mine.stdout == ref.stdout


@pytest.mark.parametrize("args", [
    ("-i", "nosuchformat", "a", "b", "out"),                 # Invalid input format
    ("-o", "nosuchformat", "a", "b", "out"),                 # Invalid output format
    ("-i", "native,plain_default,sga", "a", "b", "out"),     # 3 formats for 2 inputs
    ("a", "out"),                                            # a single input: no output file
    ("-t", "4", "only_one"),
])
def test_option_errors_match_the_reference(tools, args):
    mine, ref = run(tools[0], *args), run(tools[1], *args)
    assert mine.returncode != 0 and ref.returncode != 0
    assert mine.stderr.strip().splitlines()[-1] == ref.stderr.strip().splitlines()[-1]


def test_options_may_be_glued_or_separate(tools, tmp_path):
    """-t4 and -t 4, options after the file names (GNU getopt permutes them, so does the table-driven parser)."""
    for args in (("-t4", "-iplain_default", "-oplain_default"), ("-t", "4", "-i", "plain_default", "-o", "plain_default")):
        for order in (lambda files: list(args) + files, lambda files: files[:1] + list(args) + files[1:]):
            res = run(tools[0], *order([str(tmp_path / "missing_a"), str(tmp_path / "missing_b"), str(tmp_path / "out")]))
            # parsing succeeded: the report names both inputs with the chosen format before anything is opened
            assert "missing_a (plain_default)" in res.stdout and "missing_b (plain_default)" in res.stdout
            assert "Threads:          4" in res.stdout or "Threads:" in res.stdout
