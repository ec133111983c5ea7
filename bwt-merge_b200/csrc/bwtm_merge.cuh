// Internal interfaces between the stages of the merge (K1 walk, K2 sort, K4 interleave, K5 encode).
#pragma once

#include "bwtm_internal.cuh"

namespace bwtm
{

// Optional by-product of the walk: counts[(value >> shift) & (bins - 1)] += 1 for every emitted value, the histogram
// of the first partition level of the sort (collected in shared memory while the values are flushed).
struct WalkHistogram
{
  unsigned long long* counts;   // bins zeroed counters on the device, or null
  int                 shift;
  unsigned int        bins;     // power of two, at most 1024
  // Alternatively the histogram of ALL the partitioned bits, counted with global reductions (the walk waits for DRAM,
  // its L2 has atomic throughput to spare): fine_counts[value >> fine_shift] += 1. Then `counts` is not used.
  unsigned long long* fine_counts;
  int                 fine_shift;
};

// K1. Appends one RA value per suffix of the sequences [seq_first, seq_last] of b to d_out (unordered).
template<class KeyT>
int walk_sequences(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                   KeyT* d_out, uint64_t capacity, uint64_t* emitted, cudaStream_t stream, const WalkHistogram* histogram = nullptr);

// K1 without waiting: enqueues the walk on `stream`; RA values are appended at *cursor, which consecutive
// launches may share. `counters` (walk_counters_bytes() zeroed bytes) is private to the launch.
template<class KeyT>
int walk_sequences_async(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                         KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor,
                         int max_blocks_per_sm, cudaStream_t stream, const WalkHistogram* histogram = nullptr);
// K1, two-step form (bwtm_pairs.cu): both indexes carry pair records.
template<class KeyT>
int walk_pairs_async(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                     KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor, cudaStream_t stream,
                     const WalkHistogram* histogram = nullptr);
// Decides how b is walked against a and builds what that needs: pair records on both (two backward steps per
// record read) when it pays for the `walked_bases` symbols of b this GPU searches and fits, else nothing (single-step walk on the basic records). Fills the walk_* fields
// and pair_index_seconds of `timings`.
int prepare_walk(bwtm_index* a, bwtm_index* b, uint64_t walked_bases, cudaStream_t stream, bwtm_timings* timings);
bool walk_uses_pairs(const bwtm_index* a, const bwtm_index* b);
uint64_t walk_counters_bytes();
int walk_counters_check(const void* host_copy);   // non-zero: the output buffer overflowed

// K2. Radix sort on the low `bits` bits; *sorted points into d_keys or d_alt.
template<class KeyT>
int sort_keys(KeyT* d_keys, KeyT* d_alt, uint64_t n, int bits, KeyT** sorted, cudaStream_t stream, uint64_t key_limit = 0,   // key_limit: exclusive bound of the key values, when known
              const WalkHistogram* walked = nullptr);   // histogram delivered by the walk (sort_plan_histogram)
// True when sort_keys(n, bits, key_limit) partitions the high bits itself; then `plan` describes the histogram the walk
// can deliver (counts / fine_counts are left null: the caller allocates bins, or *fine_bins, zeroed counters).
bool sort_plan_histogram(uint64_t n, int bits, uint64_t key_limit, WalkHistogram* plan, uint64_t* fine_bins);

// Sequential state of the byte encoder that crosses slabs (and GPU slices).
struct EncodeControl
{
  unsigned long long out_size;      // bytes written so far (the `array.size()` of Run::write, support.h:267)
  unsigned long long slab_base;     // out_size at the start of the slab's parallel part
  unsigned long long carry_len;     // pending maximal run not yet written (RunBuffer state, utils.h:121-142)
  unsigned int       carry_sym;
  unsigned int       start;         // first run of the slab encoded by the parallel part
  unsigned long long count;         // number of runs encoded by the parallel part
  unsigned long long n_short;       // of which shorter than MAX_RUN
  unsigned long long n_long;
  unsigned long long long_bytes;    // bytes produced by the long runs of the slab
  unsigned long long runs_total;    // maximal runs emitted so far
};

// K4 + K3 + K5 over the merged positions [begin, end): symbols of a and b interleaved according to
// the sorted RA values `keys` (keys[j - key_base] is the RA value of b's position j), run detection
// and the byte-exact writer.  Appends to *d_out (grown when needed); the encoder state lives in
// d_control (device) and crosses calls.  `finish` flushes the pending run.
// Optional host destination that receives finished output bytes while later slabs are still being encoded.
struct HostSink
{
  uint8_t*     ptr;
  uint64_t     capacity;
  uint64_t     copied;     // bytes already queued for copying
  cudaStream_t stream;     // non-blocking copy stream
  cudaEvent_t  ready;
  bool         device_visible;   // page-locked host memory: copied by a kernel instead of the DMA engine
};

struct OutputBuffer
{
  uint8_t* ptr;
  uint64_t capacity;
  uint64_t origin;     // global byte offset of ptr[0]: a GPU slice of the output starts where the previous slice ended
  HostSink* sink;      // may be null
  bool borrowed = false;   // ptr belongs to someone else (a peer window): never freed or grown here
  uint8_t* at_origin() const { return ptr - origin; }   // kernels index this with global offsets
};

// Queues the copy of all finished bytes (below the writer's current offset) to the sink.
int flush_to_host(OutputBuffer* out, const EncodeControl* d_control, cudaStream_t stream);

// CUDA-event stopwatch on one stream.
struct EventTimer
{
  cudaEvent_t begin, end;
  cudaStream_t stream;
  bool ok;
  explicit EventTimer(cudaStream_t s) : stream(s)
  {
    ok = (cudaEventCreate(&begin) == cudaSuccess && cudaEventCreate(&end) == cudaSuccess);
  }
  ~EventTimer() { if(ok) { cudaEventDestroy(begin); cudaEventDestroy(end); } }
  void start() { if(ok) { cudaEventRecord(begin, stream); } }
  float stop()
  {
    if(!ok) { return 0.0f; }
    cudaEventRecord(end, stream); cudaEventSynchronize(end);
    float ms = 0.0f; cudaEventElapsedTime(&ms, begin, end); return ms;
  }
};

// K3 + K5 on a stream of symbol slabs: run detection and the byte-exact Run::write.
// What the sequential glue (enc_head) needs to know about a slab: its number of maximal runs, and the first
// and the last of them.
struct SlabEnds
{
  unsigned int runs;
  unsigned int last_start;      // slab position where the last run starts
  unsigned int first_sym, first_len, last_sym, last_len;
};

struct SlabEncoder
{
  uint64_t max_symbols;
  uint64_t long_capacity;                       // long runs the transducer arrays can hold
  const uint4* slab_planes;                     // input of the last detect(), read again by emit()
  uint64_t slab_symbols;
  DeviceBuffer tile_count, tile_first, tile_last, next_first, group_first, class_base, ends, long_len, long_shorts, tile_bytes, tile_entry,
               long_offset, checkpoints, placed, cub_temp;
  uint64_t detected_runs;                       // result of detect(): maximal runs of the slab
  uint64_t part_count, part_short, part_long;   // its parallel part, runs [1, m - 1)
  int init(uint64_t max_symbols, cudaStream_t stream);
  int reserve_long(uint64_t long_runs);
  // K3: maximal runs of the slab given as plane chunks; independent of the encoder state.
  int detect(const uint4* d_planes, uint64_t symbols, cudaStream_t stream);
  // K5: writes the runs found by detect() continuing from the state in d_control. write = advance + emit:
  // advance() moves the writer state past this slab (sequential, tiny), emit() writes the bytes.
  int advance(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream);
  int emit(OutputBuffer* out, cudaStream_t stream);
  int write(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream);
  int encode(const uint4* d_planes, uint64_t symbols, OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream);
  int finish(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream);
};

// The run detection counts items in 32-bit ints: a slab holds fewer than 2^31 symbols.
constexpr uint64_t MAX_SLAB_SYMBOLS = (1ull << 31) - (1ull << 16);
uint64_t clamp_slab(uint64_t slab_symbols, uint64_t total, bool allow_large = false);
int ensure_capacity(OutputBuffer* out, uint64_t needed, uint64_t valid_bytes, cudaStream_t stream);

// Symbols of a complete sequence (device, one comp per byte) -> index. Used by the fixture builder.
int index_from_symbols(const uint8_t* d_symbols, uint64_t n, uint64_t slab_symbols, cudaStream_t stream, bwtm_index** out);

// Wraps encoded RLE bytes (out->ptr[0 .. rle_bytes)) into an index; frees the buffer. counts may be NULL unless skip_index.
int finish_index(OutputBuffer* out, uint64_t rle_bytes, const uint64_t* counts, uint64_t sequences, bool skip_index,
                 cudaStream_t stream, bwtm_index** result, DeviceBuffer* filled_records = nullptr, uint64_t size = 0);
int bit_length_host(uint64_t v);

int merge_local(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
                bwtm_index** result, bwtm_timings* timings);

// K4 for one slab: merged symbols of positions [p0, p1) as plane chunks, 16 bytes per 32 positions, chunk 0 at
// d_merged[0] ((p1 - p0) / 32 rounded up chunks; tile_j: (p1 - p0) / 4096 + 2 entries).
template<class KeyT>
int interleave_slab(const bwtm_index* a, const bwtm_index* b, const KeyT* d_keys, uint64_t key_base, uint64_t key_count,
                    uint64_t p0, uint64_t p1, uint4* d_merged, uint64_t* d_tile_j, cudaStream_t stream,
                    unsigned long long* d_distinct_keys = nullptr);
uint64_t interleave_tile_size();

template<class KeyT>
int interleave_range(const bwtm_index* a, const bwtm_index* b, const KeyT* d_keys, uint64_t key_base, uint64_t key_count,
                     uint64_t begin, uint64_t end, uint64_t slab_symbols,
                     OutputBuffer* out, EncodeControl* d_control, bool finish,
                     float* interleave_ms, float* encode_ms, cudaStream_t stream, unsigned long long* d_distinct_keys = nullptr,
                     uint4* d_result_records = nullptr);

} // namespace bwtm
