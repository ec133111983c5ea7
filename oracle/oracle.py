"""TEST INFRASTRUCTURE -- ctypes front-ends for the two checkers.  Not part of the product.

* ``Oracle``   : oracle/_build/libbwtm_oracle.so, the plain-C restatement (oracle/bwtm_oracle.c).
* ``RefHooks`` : oracle/_ref/libref_hooks.so, the UNMODIFIED reference classes compiled by
                 oracle/Makefile (present only where /root/reference was available at build time,
                 or where the prebuilt file travelled with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libbwtm_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HOOKS_SO = os.path.join(REF_DIR, "libref_hooks.so")
REFERENCE_SRC = "/root/reference"

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def build(ref=True):
    """Compile the C restatement and, when the reference sources are present, oracle/_ref (and, once the product
    library exists, the reference's own bwt_merge bound to it: target ref_b200)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_SRC):
        subprocess.check_call(["make", "-s", "-C", HERE, "-j8", "ref"])
        if os.path.exists(os.path.join(HERE, "..", "bwt-merge_b200", "lib", "libbwtm_b200.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref_b200"])


def ref_available():
    return os.path.exists(REF_HOOKS_SO) and os.path.exists(os.path.join(REF_DIR, "bwt_merge"))


def _ptr(a, t):
    return a.ctypes.data_as(t)


class OrcRun(C.Structure):
    _fields_ = [("pos", C.c_uint64), ("len", C.c_uint64)]


class OrcBytes(C.Structure):
    _fields_ = [("data", u8p), ("size", C.c_uint64), ("capacity", C.c_uint64)]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        self.L = L
        vp = C.c_void_p
        L.orc_bwt_from_rle.restype = vp; L.orc_bwt_from_rle.argtypes = [u8p, C.c_uint64]
        L.orc_bwt_from_comps.restype = vp; L.orc_bwt_from_comps.argtypes = [u8p, C.c_uint64]
        L.orc_bwt_free.argtypes = [vp]
        for f in ("bytes", "size", "sequences", "hash", "blocks"):
            getattr(L, "orc_bwt_" + f).restype = C.c_uint64
            getattr(L, "orc_bwt_" + f).argtypes = [vp]
        L.orc_bwt_rle.restype = u8p; L.orc_bwt_rle.argtypes = [vp]
        L.orc_bwt_counts.argtypes = [vp, u64p]
        L.orc_bwt_decode.argtypes = [vp, u8p]
        L.orc_bwt_samples.argtypes = [vp, u64p, u64p]
        L.orc_rank.restype = C.c_uint64; L.orc_rank.argtypes = [vp, C.c_uint64, C.c_uint8]
        L.orc_ranks.argtypes = [vp, C.c_uint64, u64p]
        L.orc_ranks_range.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, u64p]
        L.orc_inverse_select.argtypes = [vp, C.c_uint64, u64p, u8p]
        L.orc_access.restype = C.c_uint8; L.orc_access.argtypes = [vp, C.c_uint64]
        L.orc_find.argtypes = [vp, u8p, C.c_uint64, u64p, u64p]
        L.orc_count.restype = C.c_uint64; L.orc_count.argtypes = [vp, u8p, C.c_uint64]
        L.orc_build_ra_dfs.restype = C.c_uint64
        L.orc_build_ra_dfs.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.POINTER(C.POINTER(OrcRun))]
        L.orc_build_ra_walk.restype = C.c_uint64
        L.orc_build_ra_walk.argtypes = [vp, vp, C.c_uint64, C.c_uint64, u64p]
        L.orc_sort_compress.restype = C.c_uint64; L.orc_sort_compress.argtypes = [C.POINTER(OrcRun), C.c_uint64]
        L.orc_free.argtypes = [vp]
        L.orc_interleave.restype = vp; L.orc_interleave.argtypes = [vp, vp, C.POINTER(OrcRun), C.c_uint64]
        L.orc_merge.restype = vp; L.orc_merge.argtypes = [vp, vp, C.c_int]
        L.orc_build_bwt_from_reads.restype = C.c_int
        L.orc_build_bwt_from_reads.argtypes = [u8p, u64p, C.c_uint64, u8p]
        L.orc_run_write.argtypes = [C.POINTER(OrcBytes), C.c_uint8, C.c_uint64]
        L.orc_run_read.argtypes = [u8p, u64p, u8p, u64p]
        L.orc_bytes_init.argtypes = [C.POINTER(OrcBytes)]
        L.orc_bytes_free.argtypes = [C.POINTER(OrcBytes)]
        L.orc_bytes_push.argtypes = [C.POINTER(OrcBytes), C.c_uint8]
        L.orc_bytecode_write.argtypes = [C.POINTER(OrcBytes), C.c_uint64]
        L.orc_bytecode_read.restype = C.c_uint64; L.orc_bytecode_read.argtypes = [u8p, u64p]

    # -- codecs -------------------------------------------------------------
    def run_write(self, prefix, comp, length):
        """Append Run::write(comp, length) to ``prefix`` (bytes); returns the new byte string."""
        b = OrcBytes(); self.L.orc_bytes_init(C.byref(b))
        for x in prefix:
            self.L.orc_bytes_push(C.byref(b), x)
        self.L.orc_run_write(C.byref(b), comp, length)
        out = bytes(bytearray(b.data[i] for i in range(b.size)))
        self.L.orc_bytes_free(C.byref(b))
        return out

    def encode_runs(self, runs):
        """Run::write for a sequence of (comp, length) runs starting from an empty array."""
        b = OrcBytes(); self.L.orc_bytes_init(C.byref(b))
        for comp, length in runs:
            self.L.orc_run_write(C.byref(b), int(comp), int(length))
        out = np.ctypeslib.as_array(b.data, shape=(b.size,)).copy() if b.size else np.zeros(0, np.uint8)
        self.L.orc_bytes_free(C.byref(b))
        return out

    def decode_runs(self, rle):
        rle = np.ascontiguousarray(rle, dtype=np.uint8)
        padded = np.concatenate([rle, np.zeros(16, np.uint8)])
        i = C.c_uint64(0); comp = C.c_uint8(0); length = C.c_uint64(0)
        runs = []
        while i.value < len(rle):
            self.L.orc_run_read(_ptr(padded, u8p), C.byref(i), C.byref(comp), C.byref(length))
            runs.append((comp.value, length.value))
        return runs

    def bytecode_write(self, value):
        b = OrcBytes(); self.L.orc_bytes_init(C.byref(b))
        self.L.orc_bytecode_write(C.byref(b), value)
        out = bytes(bytearray(b.data[i] for i in range(b.size)))
        self.L.orc_bytes_free(C.byref(b))
        return out

    def bytecode_read(self, data, i=0):
        arr = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8).copy()
        pos = C.c_uint64(i)
        value = self.L.orc_bytecode_read(_ptr(arr, u8p), C.byref(pos))
        return value, pos.value

    # -- BWT objects ----------------------------------------------------------
    def from_rle(self, rle):
        rle = np.ascontiguousarray(rle, dtype=np.uint8)
        return OracleBWT(self, self.L.orc_bwt_from_rle(_ptr(rle, u8p), len(rle)))

    def from_comps(self, comps):
        comps = np.ascontiguousarray(comps, dtype=np.uint8)
        return OracleBWT(self, self.L.orc_bwt_from_comps(_ptr(comps, u8p), len(comps)))

    def bwt_of_reads(self, reads):
        """reads: list of uint8 comp arrays (values 1..5). Returns the BWT as a comp array."""
        starts = np.zeros(len(reads) + 1, dtype=np.uint64)
        starts[1:] = np.cumsum([len(r) for r in reads])
        comps = (np.concatenate(reads) if len(reads) else np.zeros(0, np.uint8)).astype(np.uint8)
        comps = np.ascontiguousarray(comps)
        out = np.zeros(int(starts[-1]) + len(reads), dtype=np.uint8)
        rc = self.L.orc_build_bwt_from_reads(_ptr(comps, u8p), _ptr(starts, u64p), len(reads), _ptr(out, u8p))
        assert rc == 0
        return out

    def build_ra_dfs(self, a, b, first=None, last=None):
        first = 0 if first is None else first
        last = b.sequences - 1 if last is None else last
        p = C.POINTER(OrcRun)()
        n = self.L.orc_build_ra_dfs(a.h, b.h, first, last, C.byref(p))
        arr = np.zeros((n, 2), dtype=np.uint64)
        if n:
            arr[:] = np.ctypeslib.as_array(C.cast(p, u64p), shape=(n, 2))
        self.L.orc_free(p)
        return arr

    def build_ra_walk(self, a, b, first=None, last=None):
        first = 0 if first is None else first
        last = b.sequences - 1 if last is None else last
        out = np.zeros(b.size + 1, dtype=np.uint64)
        n = self.L.orc_build_ra_walk(a.h, b.h, first, last, _ptr(out, u64p))
        return out[:n].copy()

    def sort_compress(self, runs):
        runs = np.ascontiguousarray(runs, dtype=np.uint64).reshape(-1, 2).copy()
        extra = np.zeros((len(runs) + 1, 2), dtype=np.uint64); extra[:len(runs)] = runs
        n = self.L.orc_sort_compress(C.cast(_ptr(extra, u64p), C.POINTER(OrcRun)), len(runs))
        return extra[:n].copy()

    def interleave(self, a, b, ra_runs):
        ra_runs = np.ascontiguousarray(ra_runs, dtype=np.uint64).reshape(-1, 2)
        h = self.L.orc_interleave(a.h, b.h, C.cast(_ptr(ra_runs, u64p), C.POINTER(OrcRun)), len(ra_runs))
        return OracleBWT(self, h)

    def merge(self, a, b, use_dfs=True):
        return OracleBWT(self, self.L.orc_merge(a.h, b.h, 1 if use_dfs else 0))


class OracleBWT:
    def __init__(self, oracle, handle):
        self.o = oracle; self.h = handle

    def __del__(self):
        try:
            self.o.L.orc_bwt_free(self.h)
        except Exception:
            pass

    @property
    def size(self): return self.o.L.orc_bwt_size(self.h)
    @property
    def sequences(self): return self.o.L.orc_bwt_sequences(self.h)
    @property
    def bytes(self): return self.o.L.orc_bwt_bytes(self.h)
    @property
    def blocks(self): return self.o.L.orc_bwt_blocks(self.h)

    def rle(self):
        n = self.bytes
        return np.ctypeslib.as_array(self.o.L.orc_bwt_rle(self.h), shape=(n,)).copy() if n else np.zeros(0, np.uint8)

    def counts(self):
        out = np.zeros(6, dtype=np.uint64); self.o.L.orc_bwt_counts(self.h, _ptr(out, u64p)); return out

    def C(self):
        out = np.zeros(7, dtype=np.uint64); out[1:] = np.cumsum(self.counts()); return out

    def decode(self):
        out = np.zeros(self.size, dtype=np.uint8); self.o.L.orc_bwt_decode(self.h, _ptr(out, u8p)); return out

    def hash(self): return self.o.L.orc_bwt_hash(self.h)

    def samples(self):
        k = self.blocks
        ends = np.zeros(k, dtype=np.uint64); cum = np.zeros((6, k), dtype=np.uint64)
        self.o.L.orc_bwt_samples(self.h, _ptr(ends, u64p), _ptr(cum, u64p))
        return ends, cum

    def rank(self, i, c): return self.o.L.orc_rank(self.h, int(i), int(c))

    def ranks(self, i):
        out = np.zeros(6, dtype=np.uint64); self.o.L.orc_ranks(self.h, int(i), _ptr(out, u64p)); return out

    def ranks_range(self, sp, ep):
        f = np.zeros(6, dtype=np.uint64); s = np.zeros(6, dtype=np.uint64)
        self.o.L.orc_ranks_range(self.h, int(sp), int(ep), _ptr(f, u64p), _ptr(s, u64p)); return f, s

    def inverse_select(self, i):
        r = C.c_uint64(0); c = C.c_uint8(0)
        self.o.L.orc_inverse_select(self.h, int(i), C.byref(r), C.byref(c)); return r.value, c.value

    def access(self, i): return self.o.L.orc_access(self.h, int(i))

    def find(self, comps):
        comps = np.ascontiguousarray(comps, dtype=np.uint8)
        sp = C.c_uint64(0); ep = C.c_uint64(0)
        self.o.L.orc_find(self.h, _ptr(comps, u8p), len(comps), C.byref(sp), C.byref(ep)); return sp.value, ep.value

    def count(self, comps):
        comps = np.ascontiguousarray(comps, dtype=np.uint8)
        return self.o.L.orc_count(self.h, _ptr(comps, u8p), len(comps))


class RefHooks:
    """The unmodified reference classes (FMI, BWT, Run, ByteCode) behind a C interface."""

    def __init__(self):
        if not os.path.exists(REF_HOOKS_SO):
            raise FileNotFoundError(REF_HOOKS_SO)
        L = C.CDLL(REF_HOOKS_SO)
        self.L = L
        vp = C.c_void_p
        L.ref_run_write.restype = C.c_uint64; L.ref_run_write.argtypes = [u8p, C.c_uint64, C.c_uint8, C.c_uint64]
        L.ref_run_read.argtypes = [u8p, C.c_uint64, u64p, u8p, u64p]
        L.ref_bytecode_write.restype = C.c_uint64; L.ref_bytecode_write.argtypes = [u8p, C.c_uint64]
        L.ref_bytecode_read.restype = C.c_uint64; L.ref_bytecode_read.argtypes = [u8p, C.c_uint64, u64p]
        L.ref_fmi_load.restype = vp; L.ref_fmi_load.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_fmi_free.argtypes = [vp]
        L.ref_fmi_serialize.argtypes = [vp, C.c_char_p, C.c_char_p]
        for f in ("size", "sequences", "bytes", "hash"):
            getattr(L, "ref_fmi_" + f).restype = C.c_uint64; getattr(L, "ref_fmi_" + f).argtypes = [vp]
        L.ref_fmi_C.argtypes = [vp, u64p]
        L.ref_fmi_rle.argtypes = [vp, u8p]
        L.ref_rank.restype = C.c_uint64; L.ref_rank.argtypes = [vp, C.c_uint64, C.c_uint8]
        L.ref_inverse_select.argtypes = [vp, C.c_uint64, u64p, u8p]
        L.ref_ranks.argtypes = [vp, C.c_uint64, u64p]
        L.ref_ranks_range.argtypes = [vp, C.c_uint64, C.c_uint64, u64p, u64p]
        L.ref_access.restype = C.c_uint8; L.ref_access.argtypes = [vp, C.c_uint64]
        L.ref_blocks.restype = C.c_uint64; L.ref_blocks.argtypes = [vp]
        L.ref_samples.argtypes = [vp, u64p, u64p]
        L.ref_find.argtypes = [vp, C.c_char_p, C.c_uint64, u64p, u64p]
        L.ref_merge.restype = vp; L.ref_merge.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_char_p]
        L.ref_fmi_copy.restype = vp; L.ref_fmi_copy.argtypes = [vp]
        L.ref_merge_params.restype = vp
        L.ref_merge_params.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p]

    def run_write(self, prefix, comp, length):
        buf = np.zeros(len(prefix) + 64, dtype=np.uint8)
        buf[:len(prefix)] = np.frombuffer(bytes(prefix), dtype=np.uint8)
        n = self.L.ref_run_write(_ptr(buf, u8p), len(prefix), comp, length)
        return bytes(buf[:n])

    def bytecode_write(self, value):
        buf = np.zeros(16, dtype=np.uint8)
        n = self.L.ref_bytecode_write(_ptr(buf, u8p), value)
        return bytes(buf[:n])

    def bytecode_read(self, data, i=0):
        arr = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8).copy()
        pos = C.c_uint64(i)
        v = self.L.ref_bytecode_read(_ptr(arr, u8p), len(arr), C.byref(pos))
        return v, pos.value

    def load(self, filename, fmt="plain_default"):
        return RefFMI(self, self.L.ref_fmi_load(filename.encode(), fmt.encode()))

    def merge(self, a, b, threads=2, sequence_blocks=4, temp_dir="/tmp"):
        return RefFMI(self, self.L.ref_merge(a.h, b.h, threads, sequence_blocks, temp_dir.encode()))


    def merge_params(self, a, b, threads, sequence_blocks=0, run_buffer_mb=0, thread_buffer_mb=0, merge_buffers=0,
                     temp_dir="/tmp"):
        """FMI(a, b, MergeParameters) with the reference defaults unless overridden (0 = default)."""
        return RefFMI(self, self.L.ref_merge_params(a.h, b.h, threads, sequence_blocks, run_buffer_mb,
                                                    thread_buffer_mb, merge_buffers, temp_dir.encode()))

    def copy(self, fmi):
        return RefFMI(self, self.L.ref_fmi_copy(fmi.h))


class RefFMI:
    def __init__(self, hooks, handle):
        self.k = hooks; self.h = handle

    def __del__(self):
        try:
            self.k.L.ref_fmi_free(self.h)
        except Exception:
            pass

    @property
    def size(self): return self.k.L.ref_fmi_size(self.h)
    @property
    def sequences(self): return self.k.L.ref_fmi_sequences(self.h)
    @property
    def bytes(self): return self.k.L.ref_fmi_bytes(self.h)
    @property
    def blocks(self): return self.k.L.ref_blocks(self.h)

    def hash(self): return self.k.L.ref_fmi_hash(self.h)

    def C(self):
        out = np.zeros(7, dtype=np.uint64); self.k.L.ref_fmi_C(self.h, _ptr(out, u64p)); return out

    def rle(self):
        out = np.zeros(self.bytes, dtype=np.uint8); self.k.L.ref_fmi_rle(self.h, _ptr(out, u8p)); return out

    def samples(self):
        k = self.blocks
        ends = np.zeros(k, dtype=np.uint64); cum = np.zeros((6, k), dtype=np.uint64)
        self.k.L.ref_samples(self.h, _ptr(ends, u64p), _ptr(cum, u64p)); return ends, cum

    def rank(self, i, c): return self.k.L.ref_rank(self.h, int(i), int(c))

    def ranks(self, i):
        out = np.zeros(6, dtype=np.uint64); self.k.L.ref_ranks(self.h, int(i), _ptr(out, u64p)); return out

    def ranks_range(self, sp, ep):
        f = np.zeros(6, dtype=np.uint64); s = np.zeros(6, dtype=np.uint64)
        self.k.L.ref_ranks_range(self.h, int(sp), int(ep), _ptr(f, u64p), _ptr(s, u64p)); return f, s

    def inverse_select(self, i):
        r = C.c_uint64(0); c = C.c_uint8(0)
        self.k.L.ref_inverse_select(self.h, int(i), C.byref(r), C.byref(c)); return r.value, c.value

    def access(self, i): return self.k.L.ref_access(self.h, int(i))

    def find(self, pattern):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        sp = C.c_uint64(0); ep = C.c_uint64(0)
        self.k.L.ref_find(self.h, pattern, len(pattern), C.byref(sp), C.byref(ep)); return sp.value, ep.value

    def serialize(self, filename, fmt="plain_default"):
        self.k.L.ref_fmi_serialize(self.h, filename.encode(), fmt.encode())
