"""The named configurations of BASELINE.json at FULL size against the unmodified reference binary.

tests/golden/reference_merge_digests.json holds, per workload, the SHA-256 of the merged run-length bytes that
oracle/_ref/bwt_merge (the reference compiled from its own sources) produced:
  * config 1: inputs built on the CPU by the oracle's suffix sorter (oracle/make_reference_digests.py) -- so this
    also checks the GPU fixture builder, which must deliver the same inputs;
  * config 2: recorded on the GPU box by `bench.py --impl reference --record-digest` (inputs from bin/bwtm_fixture).
The run-length code is canonical (maximal runs, Run::write), so equal bytes <=> equal BWT.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import bwtm_b200
from bwtm_b200 import FMI, MergeParameters, synth
from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def merged_digest(config, sequence_blocks=0):
    sys_argv = sys.argv; sys.argv = ["bench.py", "--config", str(config)]
    try:
        args = bench.parse_args()
    finally:
        sys.argv = sys_argv
    known = bench.load_digests().get(bench.workload_key(args))
    if known is None:
        pytest.skip("no reference digest committed for config %d" % config)
    thr = synth.error_threshold(args.error)
    A = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_a, args.reads)])
    B = FMI.synthetic(args.genome, args.genome_seed, args.read_len, thr, [(args.seed_b, args.reads)])
    params = MergeParameters(); params.sequence_blocks = sequence_blocks
    M = FMI.merge(A, B, params)
    got = M.rle()
    result = {"sequences": M.sequences(), "bases": M.size(), "rle_bytes": len(got), "sha256": hashlib.sha256(got.tobytes()).hexdigest()}
    M.close()
    return result, known


def test_config1_full_size_equals_reference_binary():
    got, known = merged_digest(1)
    assert got == {k: known[k] for k in got}


def test_config1_in_five_search_batches_equals_reference_binary():
    got, known = merged_digest(1, sequence_blocks=5)
    assert got == {k: known[k] for k in got}


@pytest.mark.parametrize("wide", [False, True])
def test_config1_through_two_msd_partition_levels(monkeypatch, wide):
    """Config 1 has 24-bit keys: with the counting pass forced on (12 low bits) the 12 high bits take two MSD partition
    levels of 6 bits; with 64-bit keys as well. The result must still be the reference binary's."""
    monkeypatch.setenv("BWTM_LOCAL_SORT_MIN", "1"); monkeypatch.setenv("BWTM_LOCAL_SORT_DENSITY", "0")
    monkeypatch.setenv("BWTM_LOCAL_SORT_MAX_DENSITY", "0")
    if wide:
        monkeypatch.setenv("BWTM_FORCE_WIDE", "1")
    got, known = merged_digest(1)
    assert got == {k: known[k] for k in got}
    monkeypatch.setenv("BWTM_MSD", "0")        # the library's radix passes for the high bits: same bytes
    got, known = merged_digest(1)
    assert got == {k: known[k] for k in got}


def test_config2_full_size_equals_reference_binary():
    got, known = merged_digest(2)
    assert got == {k: known[k] for k in got}


def test_reference_binary_itself_on_config1(tmp_path):
    """Runs oracle/_ref/bwt_merge on files written by bin/bwtm_fixture (the route of bench.py's reference arm) and
    compares with the digest whose inputs came from the CPU sorter: the two input routes agree."""
    if not (os.path.exists(bench.REF_MERGE) and os.path.exists(bench.FIXTURE_TOOL)):
        pytest.skip("oracle/_ref or bin/bwtm_fixture not built")
    sys_argv = sys.argv; sys.argv = ["bench.py", "--config", "1"]
    try:
        args = bench.parse_args()
    finally:
        sys.argv = sys_argv
    known = bench.load_digests().get(bench.workload_key(args))
    if known is None:
        pytest.skip("no reference digest committed for config 1")
    a, b, out = (str(tmp_path / n) for n in ("A.native", "B.native", "out.native"))
    bench.write_fixture(args, a, [(args.seed_a, args.reads)])
    bench.write_fixture(args, b, [(args.seed_b, args.reads)])
    bench.run_reference_binary(a, b, out, os.cpu_count() or 1, str(tmp_path))
    digest = bench.native_file_digest(out)
    assert digest == {k: known[k] for k in digest}
