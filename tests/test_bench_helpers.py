"""bench.py's host-side helpers (no GPU): the digest of a native file is the SHA-256 of the reference's own BWT::data
bytes, the committed digests carry what bench.py compares, and both arms describe the workload identically."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from bwtm_b200 import synth
from conftest import ROOT, make_collection

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def parse(*argv):
    saved = sys.argv; sys.argv = ["bench.py"] + list(argv)
    try:
        return bench.parse_args()
    finally:
        sys.argv = saved


def test_native_file_digest_is_the_sha256_of_the_run_length_bytes(oracle, refhooks, tmp_path):
    from oracle.oracle import REF_DIR
    _, bwt = make_collection(oracle, 2000, 300, 50, 0.02, 42, 4)
    plain, native = str(tmp_path / "x.plain"), str(tmp_path / "x.native")
    synth.comps_to_chars(bwt).tofile(plain)
    subprocess.check_call([os.path.join(REF_DIR, "bwt_convert"), "-i", "plain_default", "-o", "native", plain, native], stdout=subprocess.DEVNULL)
    loaded = refhooks.load(native, "native")
    digest = bench.native_file_digest(native)
    assert digest["rle_bytes"] == loaded.bytes and digest["bases"] == loaded.size and digest["sequences"] == loaded.sequences
    assert digest["sha256"] == hashlib.sha256(loaded.rle().tobytes()).hexdigest()
    assert digest["sha256"] == hashlib.sha256(oracle.from_comps(bwt).rle().tobytes()).hexdigest()


def test_committed_digests_cover_the_named_two_input_configurations():
    digests = bench.load_digests()
    for config in (1, 2, 5):
        args = parse("--config", str(config))
        entry = digests.get(bench.workload_key(args))
        assert entry is not None, "no reference digest for config %d" % config
        n = args.reads * (args.read_len + 1)
        assert entry["bases"] == 2 * n and entry["sequences"] == 2 * args.reads and len(entry["sha256"]) == 64
        assert "oracle/_ref/bwt_merge" in entry["source"]


def test_both_arms_describe_the_workload_identically():
    args = parse("--config", "2")
    n = args.reads * (args.read_len + 1)
    ours = bench.config_dict(args, n, n, [1, 2, 3]); reference = bench.config_dict(parse("--config", "2", "--impl", "reference"), n, n, [1, 2, 3])
    assert ours == reference and json.dumps(ours) == json.dumps(reference)
    assert "2x10000000x150bp" in ours["workload"] and ours["inserted_bases"] == 1_510_000_000
