/*
  bwtm.h -- C ABI of the B200-native rank-array path of BWT-merge.

  This is the drop-in boundary described in SURVEY.md 8(b).  The reference (jltsiren/bwt-merge)
  has no plugin/FFI interface; the narrowest seam is one C++ constructor,

      FMI::FMI(FMI& a, FMI& b, MergeParameters parameters)          fmi.h:107-110, fmi.cpp:336-369

  called from merge() (bwt_merge.cpp:287-299).  A host program keeps reading and writing BWT
  files exactly as the reference does and calls the functions below where the reference would
  construct the merged FMI, build rank/select samples, or answer -v pattern queries.
  INTEGRATION.md shows the binding a maintainer of the reference would add.

  Conventions
    * plain pointers and sizes only; host pointers unless a name ends in `_device`;
    * every function returns BWTM_OK (0) or a negative error code and never exits or throws;
      bwtm_last_error() returns a thread-local message for the last failure;
    * comp values are the reference's: 0 = $, 1..5 (support.h:227-229, support.cpp:40-63);
    * "RLE bytes" are the reference's run-length byte code (Run::write/Run::read, support.h:221-286)
      in 64-byte blocks, i.e. the contents of BWT::data (bwt.h:173) / the BlockArray payload of a
      native file;
    * there is no CPU fallback: without a CUDA device every compute entry point fails with
      BWTM_ERR_CUDA.
*/
#ifndef BWTM_H
#define BWTM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWTM_SIGMA 6

enum
{
  BWTM_OK            =  0,
  BWTM_ERR_ARGUMENT  = -1,  /* null pointer, bad size, malformed RLE data */
  BWTM_ERR_CUDA      = -2,  /* CUDA runtime error (including "no device") */
  BWTM_ERR_MEMORY    = -3,  /* device or host allocation failed */
  BWTM_ERR_ALPHABET  = -4,  /* fmi.cpp:338-342: cannot merge BWTs with different alphabets */
  BWTM_ERR_CAPACITY  = -5,  /* output buffer too small */
  BWTM_ERR_INTERNAL  = -6,  /* consistency check failed (e.g. rank array size != |B|) */
  BWTM_ERR_COMM      = -7   /* NCCL error */
};

/* Opaque device-resident index: RLE bytes + rank structure of one BWT (replaces BWT + its samples,
   bwt.h:172-178). */
typedef struct bwtm_index bwtm_index;

/* Opaque multi-GPU communicator (one process per GPU, NCCL). */
typedef struct bwtm_comm bwtm_comm;

typedef struct
{
  uint64_t sequences;              /* NativeHeader::sequences (formats.h:44-62) = count of comp 0 */
  uint64_t bases;                  /* NativeHeader::bases = BWT::size() */
  uint64_t rle_bytes;              /* BWT::bytes() */
  uint64_t counts[BWTM_SIGMA];     /* per-comp symbol counts (the `counts` of bwt.cpp:294) */
  uint64_t C[BWTM_SIGMA + 1];      /* Alphabet::C (support.cpp:84-91) */
  uint64_t device_bytes;           /* HBM held by this index */
} bwtm_index_info;

/* Mirrors MergeParameters (fmi.h:45-80). run_buffer_size, thread_buffer_size, merge_buffers and
   temp_dir configure CPU buffers that do not exist on the device: accepted and ignored. */
typedef struct
{
  uint64_t run_buffer_size;        /* -r, ignored */
  uint64_t thread_buffer_size;     /* -b, ignored */
  uint64_t merge_buffers;          /* -m, ignored */
  uint64_t threads;                /* -t, ignored (the device schedules its own walkers) */
  uint64_t sequence_blocks;        /* > 1: B's sequences are searched in that many batches, each sorted on its own and
                                      kept as a sorted run; the interleave merges the runs range by range. Bounds the
                                      work memory to |B| keys + two batch-sized buffers instead of 2 |B| keys (the
                                      counterpart of the reference's run/thread/merge buffers, fmi.cpp:164-257).
                                      1 = one batch; 0 = chosen from the free device memory. NOT the reference's -s,
                                      which counts CPU work units (4 per thread by default). */
  const char* temp_dir;            /* -d, ignored */
  /* device-side knobs (0 = default) */
  uint64_t slab_symbols;           /* merged positions interleaved per pass (default 2^30) */
  uint32_t keep_inputs;            /* 0: a and b are destroyed by bwtm_merge, as the reference does */
  uint32_t skip_index;             /* 1: do not build the rank structure of the result (RLE only) */
  /* Streaming download (single-GPU merge): if host_output is not NULL the merged RLE bytes are also copied into
     it, overlapped with the encoding of later parts (use page-locked memory). host_output_capacity must hold
     the whole result, else BWTM_ERR_CAPACITY. The byte count is bwtm_timings.merged_bytes. */
  uint8_t* host_output;
  uint64_t host_output_capacity;
} bwtm_merge_options;

/* Per-stage device timings of the last merge (CUDA events), the GPU counterpart of the
   VERBOSE_STATUS_INFO timers (fmi.cpp:360-364, bwt.cpp:300-313). */
typedef struct
{
  double search_seconds;           /* K1 rank/LF walk             ("RA built in", part 1) */
  double sort_seconds;             /* K2 radix sort of the RA     ("RA built in", part 2) */
  double exchange_seconds;         /* multi-GPU exchange of RA values by A-position range */
  double interleave_seconds;       /* K4 merge of A, B by RA      ("BWTs merged in") */
  double encode_seconds;           /* K3/K5 run detection + byte-exact Run::write */
  double index_seconds;            /* K0 on the result            ("rank/select built in") */
  double total_seconds;            /* host wall clock of bwtm_merge */
  uint64_t ra_values;              /* = |B| */
  uint64_t ra_runs;                /* distinct A positions (the reference's RA run count), 0 if not computed */
  uint64_t merged_runs;            /* maximal runs of the merged BWT */
  uint64_t merged_bytes;           /* RLE bytes of the merged BWT */
  uint64_t walk_kernel_launches;
  uint64_t kernel_launches;        /* all kernels launched by this library during the merge */
  /* How the search read the indexes: 64-byte records answer one backward step (two record reads per inserted
     base), 128-byte pair records answer two (one read per base). Pair records are built on first use of an index
     as a merge input and stay with it; the time spent building them inside this merge is pair_index_seconds. */
  double   pair_index_seconds;
  uint64_t walk_record_bytes;
  uint64_t walk_table_bytes;       /* bytes of the structures the walk reads at random (both indexes) */
  uint64_t search_batches;         /* batches B's sequences were searched in (options.sequence_blocks) */
} bwtm_timings;

/*----------------------------------------------------------------------------*/
/* Library and device */

const char* bwtm_last_error(void);
const char* bwtm_version(void);
int bwtm_device_count(int* count);
int bwtm_set_device(int device);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t bwtm_kernel_launches(void);
/* Device memory held by the library's allocations on the current device (indexes, work buffers): bytes in use now and
   their highest value since the last reset (reset_peak != 0 restarts the high-water mark at the current use). The
   counterpart of the reference's memoryUsage() report lines (fmi.cpp:348,363; utils.cpp:98-110). */
int bwtm_memory_stats(uint64_t* used_bytes, uint64_t* peak_bytes, int reset_peak);

/*----------------------------------------------------------------------------*/
/* Index: replaces BWT::build (bwt.cpp:476-512) and BWT::setHeader (bwt.cpp:468-474). */

/* Uploads `rle_bytes` bytes of run-length code and builds the device rank structure (kernel K0).
   If `expected_counts` is not NULL (6 values) the decoded per-comp counts must match it. */
int bwtm_index_create(const uint8_t* rle, uint64_t rle_bytes, const uint64_t* expected_counts,
                      bwtm_index** out);
/* Both inputs of a merge at once, the way FMI::FMI(FMI& a, FMI& b, ...) receives them (fmi.cpp:336): the upload
   of the second one runs while the rank structure of the first one is built. Same results as two calls of
   bwtm_index_create; on failure neither index is returned. */
int bwtm_index_create_pair(const uint8_t* rle_a, uint64_t rle_bytes_a, const uint64_t* expected_counts_a,
                           const uint8_t* rle_b, uint64_t rle_bytes_b, const uint64_t* expected_counts_b,
                           bwtm_index** out_a, bwtm_index** out_b);
/* Same, from RLE bytes that already live in device memory (copied). */
int bwtm_index_create_device(const void* rle_device, uint64_t rle_bytes, bwtm_index** out);
/* From a plain symbol sequence (one comp value per byte): the device counterpart of PlainData::read
   (formats.cpp:133-161): maximal runs (RunBuffer) -> Run::write -> samples. `slab_symbols` = symbols
   encoded per pass (0 = default). */
int bwtm_index_create_plain(const uint8_t* comps, uint64_t n, uint64_t slab_symbols, bwtm_index** out);
/* From one (comp, length) run per byte, lengths 1..31, decoded on the device: the counterpart of RopeData::read
   (formats.cpp:286-310; layout BWTM_RUNS_ROPEBWT: byte = length << 3 | comp, the bytes after the 4-byte tag) and of
   SGAData::read (formats.cpp:403-429; layout BWTM_RUNS_SGA: byte = comp << 5 | length, the bytes after the header).
   Consecutive runs of one symbol are joined, as the reference's RunBuffer does. A byte with length 0 or a comp
   value above 5 is refused (BWTM_ERR_ALPHABET). */
#define BWTM_RUNS_ROPEBWT 0
#define BWTM_RUNS_SGA     1
int bwtm_index_create_runs(const uint8_t* runs, uint64_t n_runs, int layout, uint64_t slab_symbols, bwtm_index** out);
int bwtm_index_destroy(bwtm_index* index);
int bwtm_index_get_info(const bwtm_index* index, bwtm_index_info* info);
/* Copies the RLE bytes to the host (what BlockArray::serialize, support.cpp:296-309, writes after
   its length field). */
int bwtm_index_download(const bwtm_index* index, uint8_t* out_rle, uint64_t capacity, uint64_t* rle_bytes);
/* Block samples as BWT::build computes them, for the native-format writer: for every 64-byte
   block k, the last sequence position of the block and the six cumulative counts through
   block k. `block_ends` has room for `blocks` values, `cumulative` for 6 * blocks (comp-major). */
int bwtm_index_samples(const bwtm_index* index, uint64_t* block_ends, uint64_t* cumulative, uint64_t blocks);
/* Decoded symbols (comp values) of positions [first, first + count). */
int bwtm_index_extract(const bwtm_index* index, uint64_t first, uint64_t count, uint8_t* out_comps);
/* FNV-1a over the decoded sequence, BWT::hash (bwt.cpp:538-549). */
int bwtm_index_hash(const bwtm_index* index, uint64_t* hash);

/*----------------------------------------------------------------------------*/
/* Queries: BWT::rank (bwt.cpp:318-341), FMI::LF(i) (fmi.h:147-150, bwt.cpp:445-464),
   FMI::find (fmi.h:195-209). Batched: n independent queries per call. */

int bwtm_rank(const bwtm_index* index, const uint64_t* positions, const uint8_t* comps, uint64_t n,
              uint64_t* out_ranks);
int bwtm_lf(const bwtm_index* index, const uint64_t* positions, uint64_t n,
            uint64_t* out_positions, uint8_t* out_comps);
/* Occurrence counts (Range::length of FMI::find) of n patterns stored back to back in `patterns`,
   pattern k = patterns[offsets[k] .. offsets[k+1]). If char2comp is not NULL (256 entries) the
   pattern bytes are characters and are mapped through it (Alphabet::char2comp), otherwise they are
   comp values. This is the per-pattern value verifyFMI adds to `results` (bwt_merge.cpp:253-254). */
int bwtm_count(const bwtm_index* index, const uint8_t* patterns, const uint64_t* offsets, uint64_t n,
               const uint8_t* char2comp, uint64_t* out_counts);

/* Builds the pair records of an index now (2 bytes per symbol on top of the basic rank records) instead of on its
   first use as a merge input; a no-op when they exist. bwtm_merge builds them itself when the inserted collection is
   large enough to pay for it; a caller that merges the same index repeatedly, or distributes a merge over several
   GPUs (where every rank would build them for 1/G of the search), can do it once up front. */
int bwtm_index_build_pairs(bwtm_index* index);
/* Two backward steps at once, from the pair records (built on first use): for every position i, comps[2k] =
   BWT[i], comps[2k+1] = BWT[LF(i)], first = LF(i), second = LF(LF(i)) (0 where the step starts at an endmarker).
   Equals two applications of bwtm_lf; the rank-array search uses it to read one record per two inserted bases. */
int bwtm_lf2(bwtm_index* index, const uint64_t* positions, uint64_t n,
             uint64_t* out_first, uint64_t* out_second, uint8_t* out_comps);

/*----------------------------------------------------------------------------*/
/* Merge: replaces FMI::FMI(FMI& a, FMI& b, MergeParameters) (fmi.cpp:336-369):
   buildRA (fmi.cpp:272-334) -> sort/compress (fmi.cpp:220-257, support.h:415-453) ->
   interleave (bwt.cpp:194-314).  The result is BWT(sequences of a followed by sequences of b).
   `options` and `timings` may be NULL. Unless options->keep_inputs is set, a and b are destroyed
   (even on failure), as the reference destroys its inputs (fmi.h:107-109). */
int bwtm_merge(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
               bwtm_index** out, bwtm_timings* timings);

/* The sorted rank array of the sequences [seq_first, seq_last] of b with respect to a: one value
   per suffix, the multiset buildRA emits (fmi.cpp:290) after sorting (support.h:421). For tests and
   diagnostics. `capacity` values fit in out_sorted; *n_values receives the count. */
int bwtm_rank_array(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                    uint64_t* out_sorted, uint64_t capacity, uint64_t* n_values);

/*----------------------------------------------------------------------------*/
/* Multi-GPU (one process per GPU). Both indexes are replicated on every rank, the
   sequences of b are split across ranks, RA values are exchanged by A-position range with one NCCL
   all-to-all and every rank interleaves a contiguous slice of the merged BWT. */

/* Contiguous block [first, first + count) of `total` items owned by `rank` of `world`: the split of B's
   sequence ids across GPUs (one ParallelLoop block per rank, utils.cpp:169-197). Pure host arithmetic. */
int bwtm_shard_range(uint64_t total, uint32_t rank, uint32_t world, uint64_t* first, uint64_t* count);

#define BWTM_COMM_ID_BYTES 128
int bwtm_comm_unique_id(uint8_t id[BWTM_COMM_ID_BYTES]);               /* rank 0; broadcast it by any means */
int bwtm_comm_create(const uint8_t id[BWTM_COMM_ID_BYTES], int rank, int world, bwtm_comm** out);
int bwtm_comm_destroy(bwtm_comm* comm);
/* Collective: every rank calls it with its replicas of a and b. On return every rank holds the complete merged index:
   the run-length bytes of all slices gathered, the rank structure built from the plane chunks the ranks store into each
   other's record windows (or decoded from the bytes when peer windows are not available).
   It fails together: if one rank fails locally (allocation, capacity, invalid input), every rank returns an error from
   the same call and the communicator stays usable.
   options->sequence_blocks > 1 (the same value on every rank): the ranks search their share of b in that many batches,
   keep them as sorted runs, exchange the runs piece by piece and merge them range by range while interleaving -- for
   inputs whose rank-array values do not fit twice into a GPU. 0 and 1: one batch (no automatic choice here). */
int bwtm_merge_distributed(bwtm_comm* comm, bwtm_index* a, bwtm_index* b,
                           const bwtm_merge_options* options, bwtm_index** out, bwtm_timings* timings);

/*----------------------------------------------------------------------------*/
/* Fixture tools (not in the reference, which only merges: README.md:21). Device-side synthetic
   reads and a sort-based builder of the multi-string BWT, used to make
   inputs of benchmark size and to cross-check merges at sizes the CPU oracle cannot reach. */

/* BWT of `reads` reads of `read_len` bases sampled from a synthetic genome; returns a new index.
   first_read lets callers build BWT(A ++ B) directly: reads [0, n) of seed s1 followed by reads of
   seed s2 are described by two segments. */
typedef struct
{
  uint64_t seed;        /* read seed */
  uint64_t reads;       /* number of reads in this segment */
  uint64_t first_read;  /* index of the segment's first read within the seed's read stream */
} bwtm_read_segment;

int bwtm_tools_build_synthetic(uint64_t genome_len, uint64_t genome_seed, uint64_t read_len,
                               uint64_t error_threshold, const bwtm_read_segment* segments,
                               uint64_t n_segments, bwtm_index** out);
/* BWT of explicit reads: comp values 1..5, `reads` x `read_len` row-major on the host. */
int bwtm_tools_build_from_reads(const uint8_t* read_comps, uint64_t reads, uint64_t read_len,
                                bwtm_index** out);
/* Random-access HBM microbenchmark: `n_loads` aligned loads of `granule` bytes (32, 64 or 128) at
   pseudo-random offsets of a `table_bytes` table; returns achieved GB/s. The denominator for the
   rank/LF kernel's random-sector roofline. */
int bwtm_tools_gather_bench(uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iterations,
                            double* gbytes_per_second);

/* The same with the access pattern of the rank/LF kernel: walkers follow chains of DEPENDENT record reads
   (the next address depends on the loaded data); 4 lanes share a walker and read its record with one
   instruction (2 lanes for 32-byte records); `threads_per_sm` resident threads per SM (multiple of 256, at
   most 2048). The result divided by `granule` is the random record (= line request) rate. If l2_fetch_granularity is 32, 64 or 128, cudaLimitMaxL2FetchGranularity is set to
   it first (0 = leave unchanged). */
int bwtm_tools_chase_bench(uint64_t table_bytes, uint32_t granule, uint64_t n_loads, uint32_t threads_per_sm,
                           uint32_t l2_fetch_granularity, double* gbytes_per_second);

#ifdef __cplusplus
}
#endif

#endif /* BWTM_H */
