"""ctypes binding of libbwtm_b200.so (include/bwtm.h) with the reference's vocabulary.

`FMI` mirrors the part of bwtmerge::FMI (fmi.h:86-230) that lies on the rank-array path:
size(), sequences(), LF(i), LF(i, comp), find(pattern), and the merging constructor
FMI(a, b, parameters) (fmi.h:107-110) as `FMI.merge(a, b, parameters)`.  Everything is computed by the
CUDA library; there is no CPU fallback and importing the oracle from here is forbidden.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libbwtm_b200.so")

SIGMA = 6
u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)

OK = 0
ERROR_NAMES = {-1: "BWTM_ERR_ARGUMENT", -2: "BWTM_ERR_CUDA", -3: "BWTM_ERR_MEMORY", -4: "BWTM_ERR_ALPHABET",
               -5: "BWTM_ERR_CAPACITY", -6: "BWTM_ERR_INTERNAL", -7: "BWTM_ERR_COMM"}

# Every symbol include/bwtm.h declares.
EXPORTS = [
    "bwtm_last_error", "bwtm_version", "bwtm_device_count", "bwtm_set_device", "bwtm_kernel_launches", "bwtm_memory_stats",
    "bwtm_index_create", "bwtm_index_create_pair", "bwtm_index_create_device", "bwtm_index_create_plain", "bwtm_index_create_runs", "bwtm_index_destroy", "bwtm_index_get_info",
    "bwtm_index_download", "bwtm_index_samples", "bwtm_index_extract", "bwtm_index_hash",
    "bwtm_index_build_pairs", "bwtm_rank", "bwtm_lf", "bwtm_lf2", "bwtm_count", "bwtm_merge", "bwtm_rank_array",
    "bwtm_shard_range", "bwtm_comm_unique_id", "bwtm_comm_create", "bwtm_comm_destroy", "bwtm_merge_distributed",
    "bwtm_tools_build_synthetic", "bwtm_tools_build_from_reads", "bwtm_tools_gather_bench", "bwtm_tools_chase_bench",
]


class BwtmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "BWTM_ERR"), code, message))
        self.code = code


class IndexInfo(C.Structure):
    _fields_ = [("sequences", C.c_uint64), ("bases", C.c_uint64), ("rle_bytes", C.c_uint64),
                ("counts", C.c_uint64 * SIGMA), ("C", C.c_uint64 * (SIGMA + 1)), ("device_bytes", C.c_uint64)]


class MergeOptions(C.Structure):
    """bwtm_merge_options: MergeParameters (fmi.h:45-80) plus device knobs."""
    _fields_ = [("run_buffer_size", C.c_uint64), ("thread_buffer_size", C.c_uint64), ("merge_buffers", C.c_uint64),
                ("threads", C.c_uint64), ("sequence_blocks", C.c_uint64), ("temp_dir", C.c_char_p),
                ("slab_symbols", C.c_uint64), ("keep_inputs", C.c_uint32), ("skip_index", C.c_uint32),
                ("host_output", C.c_void_p), ("host_output_capacity", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("search_seconds", C.c_double), ("sort_seconds", C.c_double), ("exchange_seconds", C.c_double),
                ("interleave_seconds", C.c_double), ("encode_seconds", C.c_double), ("index_seconds", C.c_double),
                ("total_seconds", C.c_double), ("ra_values", C.c_uint64), ("ra_runs", C.c_uint64),
                ("merged_runs", C.c_uint64), ("merged_bytes", C.c_uint64), ("walk_kernel_launches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("pair_index_seconds", C.c_double), ("walk_record_bytes", C.c_uint64),
                ("walk_table_bytes", C.c_uint64), ("search_batches", C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class ReadSegment(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("reads", C.c_uint64), ("first_read", C.c_uint64)]


def build_library(verbose=False):
    """Compile libbwtm_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", PKG_DIR, "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded CUDA library. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.bwtm_last_error.restype = C.c_char_p
    L.bwtm_version.restype = C.c_char_p
    L.bwtm_device_count.argtypes = [C.POINTER(C.c_int)]
    L.bwtm_set_device.argtypes = [C.c_int]
    L.bwtm_kernel_launches.restype = C.c_uint64
    L.bwtm_memory_stats.argtypes = [u64p, u64p, C.c_int]
    L.bwtm_index_create.argtypes = [u8p, C.c_uint64, u64p, C.POINTER(vp)]
    L.bwtm_index_create_pair.argtypes = [u8p, C.c_uint64, u64p, u8p, C.c_uint64, u64p, C.POINTER(vp), C.POINTER(vp)]
    L.bwtm_index_create_device.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.bwtm_index_create_plain.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.bwtm_index_create_runs.argtypes = [u8p, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(vp)]
    L.bwtm_index_destroy.argtypes = [vp]
    L.bwtm_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.bwtm_index_download.argtypes = [vp, u8p, C.c_uint64, u64p]
    L.bwtm_index_samples.argtypes = [vp, u64p, u64p, C.c_uint64]
    L.bwtm_index_extract.argtypes = [vp, C.c_uint64, C.c_uint64, u8p]
    L.bwtm_index_hash.argtypes = [vp, u64p]
    L.bwtm_rank.argtypes = [vp, u64p, u8p, C.c_uint64, u64p]
    L.bwtm_lf.argtypes = [vp, u64p, C.c_uint64, u64p, u8p]
    L.bwtm_index_build_pairs.argtypes = [vp]
    L.bwtm_lf2.argtypes = [vp, u64p, C.c_uint64, u64p, u64p, u8p]
    L.bwtm_count.argtypes = [vp, u8p, u64p, C.c_uint64, u8p, u64p]
    L.bwtm_merge.argtypes = [vp, vp, C.POINTER(MergeOptions), C.POINTER(vp), C.POINTER(Timings)]
    L.bwtm_rank_array.argtypes = [vp, vp, C.c_uint64, C.c_uint64, u64p, C.c_uint64, u64p]
    L.bwtm_shard_range.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, u64p, u64p]
    L.bwtm_comm_unique_id.argtypes = [u8p]
    L.bwtm_comm_create.argtypes = [u8p, C.c_int, C.c_int, C.POINTER(vp)]
    L.bwtm_comm_destroy.argtypes = [vp]
    L.bwtm_merge_distributed.argtypes = [vp, vp, vp, C.POINTER(MergeOptions), C.POINTER(vp), C.POINTER(Timings)]
    L.bwtm_tools_build_synthetic.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(ReadSegment),
                                             C.c_uint64, C.POINTER(vp)]
    L.bwtm_tools_build_from_reads.argtypes = [u8p, C.c_uint64, C.c_uint64, C.POINTER(vp)]
    L.bwtm_tools_gather_bench.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    L.bwtm_tools_chase_bench.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise BwtmError(rc, lib().bwtm_last_error().decode(errors="replace"))


def _p(a, t):
    return a.ctypes.data_as(t)


def kernel_launches():
    return lib().bwtm_kernel_launches()


def memory_stats(reset_peak=False):
    """(bytes in use, peak bytes since the last reset) of the library's device allocations."""
    used = C.c_uint64(0); peak = C.c_uint64(0)
    check(lib().bwtm_memory_stats(C.byref(used), C.byref(peak), 1 if reset_peak else 0))
    return used.value, peak.value


def set_device(device):
    check(lib().bwtm_set_device(device))


class MergeParameters:
    """bwtmerge::MergeParameters (fmi.h:45-80). The CPU buffer sizes are accepted and ignored."""
    RUN_BUFFER_SIZE = 8 * 1048576
    THREAD_BUFFER_SIZE = 256 * 1048576
    MERGE_BUFFERS = 6

    def __init__(self):
        self.run_buffer_size = self.RUN_BUFFER_SIZE
        self.thread_buffer_size = self.THREAD_BUFFER_SIZE
        self.merge_buffers = self.MERGE_BUFFERS
        self.threads = 1
        self.sequence_blocks = 0     # 0: batches chosen from the free device memory (include/bwtm.h)
        self.temp_dir = "."
        self.slab_symbols = 0
        self.skip_index = False
        self.host_output = None      # numpy uint8 array (ideally page-locked): receives the merged RLE bytes while encoding

    def setRB(self, mb): self.run_buffer_size = mb * 1048576 // 16
    def setTB(self, mb): self.thread_buffer_size = mb * 1048576
    def setMB(self, n): self.merge_buffers = n
    def setT(self, n): self.threads = n
    def setSB(self, n): self.sequence_blocks = n
    def setTemp(self, directory): self.temp_dir = directory

    def to_c(self, keep_inputs):
        o = MergeOptions()
        o.run_buffer_size = self.run_buffer_size; o.thread_buffer_size = self.thread_buffer_size
        o.merge_buffers = self.merge_buffers; o.threads = self.threads; o.sequence_blocks = self.sequence_blocks
        o.temp_dir = self.temp_dir.encode(); o.slab_symbols = self.slab_symbols
        o.keep_inputs = 1 if keep_inputs else 0; o.skip_index = 1 if self.skip_index else 0
        if self.host_output is not None:
            o.host_output = self.host_output.ctypes.data; o.host_output_capacity = self.host_output.nbytes
        return o


class FMI:
    """Device-resident FM-index of one BWT (bwtmerge::FMI + BWT, fmi.h:86-230, bwt.h:41-189)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.timings = None

    # -- construction ---------------------------------------------------------
    @classmethod
    def from_rle(cls, rle, expected_counts=None):
        """From the reference's run-length bytes (BWT::data). Runs K0."""
        rle = np.ascontiguousarray(rle, dtype=np.uint8)
        h = C.c_void_p()
        exp = None
        if expected_counts is not None:
            exp = _p(np.ascontiguousarray(expected_counts, dtype=np.uint64), u64p)
        check(lib().bwtm_index_create(_p(rle, u8p), len(rle), exp, C.byref(h)))
        return cls(h)

    RUNS_ROPEBWT, RUNS_SGA = 0, 1

    @classmethod
    def from_run_bytes(cls, runs, layout, slab_symbols=0):
        """From one run per byte (RopeData::read / SGAData::read, formats.cpp:286-310, 403-429), decoded on the device."""
        runs = np.ascontiguousarray(runs, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().bwtm_index_create_runs(_p(runs, u8p), len(runs), layout, slab_symbols, C.byref(h)))
        return cls(h)

    @classmethod
    def from_rle_pair(cls, rle_a, rle_b):
        """Both inputs of a merge: the second upload overlaps the first K0 (bwtm_index_create_pair)."""
        rle_a = np.ascontiguousarray(rle_a, dtype=np.uint8); rle_b = np.ascontiguousarray(rle_b, dtype=np.uint8)
        ha, hb = C.c_void_p(), C.c_void_p()
        check(lib().bwtm_index_create_pair(_p(rle_a, u8p), len(rle_a), None, _p(rle_b, u8p), len(rle_b), None, C.byref(ha), C.byref(hb)))
        return cls(ha), cls(hb)

    @classmethod
    def from_comps(cls, comps, slab_symbols=0):
        """From a plain comp sequence (PlainData::read, formats.cpp:133-161): device run detection + Run::write."""
        comps = np.ascontiguousarray(comps, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().bwtm_index_create_plain(_p(comps, u8p), len(comps), slab_symbols, C.byref(h)))
        return cls(h)

    @classmethod
    def from_reads(cls, reads):
        """Fixture: sort-based BWT construction of a reads x read_len comp matrix on the device."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().bwtm_tools_build_from_reads(_p(reads, u8p), reads.shape[0], reads.shape[1], C.byref(h)))
        return cls(h)

    @classmethod
    def synthetic(cls, genome_len, genome_seed, read_len, error_threshold, segments):
        """Fixture: BWT of synthetic reads; segments = [(read_seed, n_reads[, first_read]), ...]."""
        segs = (ReadSegment * len(segments))(*[ReadSegment(seg[0], seg[1], seg[2] if len(seg) > 2 else 0) for seg in segments])
        h = C.c_void_p()
        check(lib().bwtm_tools_build_synthetic(genome_len, genome_seed, read_len, error_threshold, segs, len(segments), C.byref(h)))
        return cls(h)

    @classmethod
    def merge(cls, a, b, parameters=None, keep_inputs=False):
        """FMI::FMI(FMI& a, FMI& b, MergeParameters) (fmi.cpp:336-369). Destroys a and b."""
        parameters = parameters or MergeParameters()
        opts = parameters.to_c(keep_inputs)
        out = C.c_void_p(); t = Timings()
        ha, hb = a._h, b._h
        if not keep_inputs:
            a._h = None; b._h = None
        rc = lib().bwtm_merge(ha, hb, C.byref(opts), C.byref(out), C.byref(t))
        check(rc)
        m = cls(out); m.timings = t
        return m

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.bwtm_index_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's accessors ---------------------------------------------
    def info(self):
        i = IndexInfo(); check(lib().bwtm_index_get_info(self._h, C.byref(i))); return i

    def size(self): return self.info().bases
    def sequences(self): return self.info().sequences
    def bytes(self): return self.info().rle_bytes
    def counts(self): return np.array(list(self.info().counts), dtype=np.uint64)
    def C(self): return np.array(list(self.info().C), dtype=np.uint64)

    def rle(self):
        n = self.bytes()
        out = np.zeros(n, dtype=np.uint8)
        got = C.c_uint64(0)
        check(lib().bwtm_index_download(self._h, _p(out, u8p), n, C.byref(got)))
        return out

    def download_into(self, out):
        got = C.c_uint64(0)
        check(lib().bwtm_index_download(self._h, _p(out, u8p), len(out), C.byref(got)))
        return got.value

    def samples(self):
        blocks = (self.bytes() + 63) // 64
        ends = np.zeros(blocks, dtype=np.uint64); cum = np.zeros((SIGMA, blocks), dtype=np.uint64)
        check(lib().bwtm_index_samples(self._h, _p(ends, u64p), _p(cum, u64p), blocks))
        return ends, cum

    def extract(self, first=0, count=None):
        count = self.size() - first if count is None else count
        out = np.zeros(count, dtype=np.uint8)
        check(lib().bwtm_index_extract(self._h, first, count, _p(out, u8p)))
        return out

    def hash(self):
        h = C.c_uint64(0); check(lib().bwtm_index_hash(self._h, C.byref(h))); return h.value

    def rank(self, positions, comps):
        """BWT::rank(i, c) for arrays of positions and comps."""
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        comps = np.ascontiguousarray(comps, dtype=np.uint8)
        out = np.zeros(len(positions), dtype=np.uint64)
        check(lib().bwtm_rank(self._h, _p(positions, u64p), _p(comps, u8p), len(positions), _p(out, u64p)))
        return out

    def LF(self, positions, comps=None):
        """FMI::LF(i) -> (positions, comps), or FMI::LF(i, c) -> positions when comps is given."""
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        if comps is not None:
            comps = np.ascontiguousarray(comps, dtype=np.uint8)
            return self.rank(positions, comps) + self.C()[comps.astype(np.int64)]
        out = np.zeros(len(positions), dtype=np.uint64); oc = np.zeros(len(positions), dtype=np.uint8)
        check(lib().bwtm_lf(self._h, _p(positions, u64p), len(positions), _p(out, u64p), _p(oc, u8p)))
        return out, oc

    def build_pairs(self):
        """Pair records for the two-step walk, built now instead of on first use as a merge input."""
        check(lib().bwtm_index_build_pairs(self._h)); return self

    def LF2(self, positions):
        """Two backward steps from the pair records: (LF(i), LF(LF(i)), BWT[i], BWT[LF(i)])."""
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        first = np.zeros(len(positions), dtype=np.uint64); second = np.zeros(len(positions), dtype=np.uint64)
        comps = np.zeros((len(positions), 2), dtype=np.uint8)
        check(lib().bwtm_lf2(self._h, _p(positions, u64p), len(positions), _p(first, u64p), _p(second, u64p), _p(comps, u8p)))
        return first, second, comps[:, 0], comps[:, 1]

    def count(self, patterns, char2comp=None):
        """Occurrences of each pattern (Range::length of FMI::find). patterns: list of uint8 arrays / bytes."""
        arrays = [np.frombuffer(p, dtype=np.uint8) if isinstance(p, (bytes, bytearray)) else np.asarray(p, dtype=np.uint8)
                  for p in patterns]
        offsets = np.zeros(len(arrays) + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([len(a) for a in arrays])
        flat = np.ascontiguousarray(np.concatenate(arrays) if arrays else np.zeros(0, np.uint8))
        if len(flat) == 0:
            flat = np.zeros(1, dtype=np.uint8)
        out = np.zeros(len(arrays), dtype=np.uint64)
        c2c = None
        if char2comp is not None:
            c2c = _p(np.ascontiguousarray(char2comp, dtype=np.uint8), u8p)
        check(lib().bwtm_count(self._h, _p(flat, u8p), _p(offsets, u64p), len(arrays), c2c, _p(out, u64p)))
        return out


def shard_range(total, rank, world):
    """[first, first + count) of `total` sequence ids owned by `rank` (bwtm_shard_range)."""
    first = C.c_uint64(0); count = C.c_uint64(0)
    check(lib().bwtm_shard_range(total, rank, world, C.byref(first), C.byref(count)))
    return first.value, count.value


COMM_ID_BYTES = 128


def comm_unique_id():
    buf = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
    check(lib().bwtm_comm_unique_id(_p(buf, u8p)))
    return buf


class Communicator:
    """One process per GPU (bwtm_comm): NCCL communicator owned by the library."""

    def __init__(self, unique_id, rank, world):
        unique_id = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert len(unique_id) == COMM_ID_BYTES
        self._h = C.c_void_p()
        self.rank, self.world = rank, world
        check(lib().bwtm_comm_create(_p(unique_id, u8p), rank, world, C.byref(self._h)))

    @classmethod
    def from_torch(cls, dist, rank, world):
        """Rank 0 creates the NCCL id; it is broadcast through the given torch.distributed module."""
        import torch
        ident = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            ident = torch.from_numpy(comm_unique_id().copy())
        if dist.get_backend() == "nccl":
            ident = ident.cuda(); dist.broadcast(ident, 0); ident = ident.cpu()
        else:
            dist.broadcast(ident, 0)
        return cls(ident.numpy(), rank, world)

    def merge(self, a, b, parameters=None, keep_inputs=False):
        """bwtm_merge_distributed: collective; every rank returns the complete merged FMI."""
        parameters = parameters or MergeParameters()
        opts = parameters.to_c(keep_inputs)
        out = C.c_void_p(); t = Timings()
        ha, hb = a._h, b._h
        if not keep_inputs:
            a._h = None; b._h = None
        check(lib().bwtm_merge_distributed(self._h, ha, hb, C.byref(opts), C.byref(out), C.byref(t)))
        m = FMI(out); m.timings = t
        return m

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.bwtm_comm_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rank_array(a, b, seq_first=0, seq_last=None):
    """Sorted RA values of b's sequences [seq_first, seq_last] with respect to a (buildRA + sort)."""
    seq_last = b.sequences() - 1 if seq_last is None else seq_last
    cap = b.size() + 1
    out = np.zeros(cap, dtype=np.uint64); n = C.c_uint64(0)
    check(lib().bwtm_rank_array(a._h, b._h, seq_first, seq_last, _p(out, u64p), cap, C.byref(n)))
    return out[:n.value]


def gather_bench(table_bytes, granule, n_loads, iterations=3):
    g = C.c_double(0)
    check(lib().bwtm_tools_gather_bench(table_bytes, granule, n_loads, iterations, C.byref(g)))
    return g.value


def chase_bench(table_bytes, granule, n_loads, threads_per_sm=1024, l2_fetch_granularity=0):
    """GB/s of dependent random loads of `granule` bytes: the rank/LF kernel's access pattern."""
    g = C.c_double(0)
    check(lib().bwtm_tools_chase_bench(table_bytes, granule, n_loads, threads_per_sm, l2_fetch_granularity, C.byref(g)))
    return g.value


DEFAULT_CHAR2COMP = np.full(256, 5, dtype=np.uint8)   # support.cpp:40-61
for _ch, _c in ((0, 0), (ord("$"), 0), (ord("A"), 1), (ord("a"), 1), (ord("C"), 2), (ord("c"), 2),
                (ord("G"), 3), (ord("g"), 3), (ord("T"), 4), (ord("t"), 4)):
    DEFAULT_CHAR2COMP[_ch] = _c
