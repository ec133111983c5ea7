// K1: the rank-array search (single-step form), and K2: the sort of its values (MSD partition of the high key
// bits + counting of the low bits; the library's radix sort for small inputs).
//
// Replaces buildRA (fmi.cpp:272-334).  The reference walks the reverse trie of B depth first and
// emits (rank in A, number of suffixes) per trie node; 87-96 % of its nodes are singletons
// (SURVEY.md section 0), so the device does what the reference does for a singleton node
// (fmi.cpp:296-303) for every suffix: one walker per sequence of B steps backwards with
//     (c, b') = LF_B(b)          FMI::LF(i),    fmi.h:147-150  -> BWT::inverse_select, bwt.cpp:445-464
//     a'      = LF_A(a, c)       FMI::LF(i, c), fmi.h:152-155  -> BWT::rank,           bwt.cpp:318-341
// starting from (a, b) = (sequences of A, sequence id) (fmi.cpp:286) and emits `a` for every suffix.
// The emitted multiset equals the reference's RA after run expansion.
//
// Four lanes = one walker (k1_walk_coop below); a warp refills finished walkers from a global sequence
// counter, so lanes stay busy for any mix of sequence lengths.  Every step is two 64-byte record reads
// (B and A) at unrelated addresses: the kernel is bound by HBM random line requests and hides the latency
// with occupancy.  RA values are staged per warp in shared memory and appended to the output in coalesced
// chunks claimed with one atomic per chunk.  When both indexes carry pair records, the two-step form in
// bwtm_pairs.cu (one 128-byte record per side and TWO steps) replaces it.
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "bwtm_internal.cuh"
#include "bwtm_merge.cuh"

namespace bwtm
{

constexpr int WALK_THREADS = 256;
constexpr int WALK_WARPS   = WALK_THREADS / 32;
constexpr int WALK_STAGE   = 256;   // staged RA values per warp

struct WalkCounters
{
  unsigned long long next_sequence;   // relative to seq_begin, per launch
  unsigned long long emitted;         // output cursor (may be shared by consecutive launches: `cursor`)
  int                overflow;
};

// K1, cooperative form.  Measured on B200 (profiles/r01_random_line_ceiling_coop_chase.txt): dependent random
// reads are limited by the number of 128-byte LINE REQUESTS, about 39.4 G lines/s, whatever the record size
// (32, 64 or 128 bytes), and a record read by four consecutive 16-byte loads of one thread costs two
// requests.  So four lanes share one walker: lane `sub` loads chunk `sub` of the record, the whole 64-byte
// record is one instruction and one line request, the in-record rank is a popcount per lane plus two
// shuffles, and the header counter is fetched from the lanes that hold its words.  Loads carry the
// L2::64B hint: without it every record miss fetched a full 128-byte line from HBM (346 GB of DRAM reads
// for 193 GB of records, profiles/r01_k1_walk_v1_ncu_raw.csv).
constexpr int COOP_LANES = 4;

__device__ __forceinline__ uint4 load_chunk_64B(const uint4* p)
{
  uint4 r;
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// In-record part of a cooperative rank: (count of comp before `pos` inside the record) + (the record's
// header counter of comp). `q` is this lane's chunk; the caller adds the superblock and C counters.
__device__ __forceinline__ uint32_t coop_record_rank(const uint4& q, uint32_t offset, uint32_t comp,
                                                     unsigned mask, int sub, int group_base)
{
  int k = (int)offset - 32 * sub;
  k = (k < 0 ? 0 : k);
  uint32_t count = __popc(match_mask(q, comp) & low_mask(k));
  count += __shfl_xor_sync(mask, count, 1);
  count += __shfl_xor_sync(mask, count, 2);
  uint32_t s = 25u * (comp - 1u);
  uint32_t w = s >> 5, shift = s & 31u;
  uint32_t lo = __shfl_sync(mask, q.w, group_base + (int)w);
  uint32_t hi = __shfl_sync(mask, q.w, group_base + (int)(w < 3 ? w + 1 : 3));
  return (__funnelshift_r(lo, hi, shift) & FIELD_MASK) + count;
}

// PosT = uint32_t when both BWTs are shorter than 2^32 (halves the address and rank arithmetic).
template<class KeyT, class PosT>
__global__ void __launch_bounds__(WALK_THREADS, 8)
k1_walk_coop(DeviceIndex a, DeviceIndex b, uint64_t seq_begin, uint64_t seq_end,
             KeyT* __restrict__ out, uint64_t capacity, WalkCounters* counters, unsigned long long* cursor, WalkHistogram histogram)
{
  __shared__ KeyT stage_all[WALK_WARPS][WALK_STAGE];
  __shared__ PosT c_a[8], c_b[8];
  __shared__ unsigned int digit_counts[1024];

#pragma unroll
  for(int c = 0; c <= SIGMA; c++) { if(threadIdx.x == c) { c_a[c] = (PosT)a.C[c]; c_b[c] = (PosT)b.C[c]; } }
  for(unsigned int d = threadIdx.x; d < 1024; d += WALK_THREADS) { digit_counts[d] = 0; }
  __syncthreads();
  const bool counting = (histogram.counts != nullptr), counting_fine = (histogram.fine_counts != nullptr);
  const unsigned int digit_mask = histogram.bins - 1;

  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned LEADERS = 0x11111111u;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (COOP_LANES - 1);
  const int group_base = lane & ~(COOP_LANES - 1);
  const unsigned leaders_below = LEADERS & ((1u << group_base) - 1u);
  KeyT* stage = stage_all[threadIdx.x >> 5];
  const PosT first_rank = (PosT)a.sequences;
  const uint4* __restrict__ records_a = a.records + sub;
  const uint4* __restrict__ records_b = b.records + sub;

  uint32_t fill = 0;          // warp-uniform
  bool exhausted = false;     // warp-uniform
  bool alive = false;         // uniform within a group
  PosT pos_a = 0, pos_b = 0;

  while(true)
  {
    unsigned need = __ballot_sync(FULL, !alive) & LEADERS;
    if(need != 0 && !exhausted)
    {
      unsigned long long base = 0;
      int wanted = __popc(need);
      if(lane == 0) { base = atomicAdd(&(counters->next_sequence), (unsigned long long)wanted); }
      base = __shfl_sync(FULL, base, 0);
      uint64_t first = seq_begin + base;
      if(!alive)
      {
        uint64_t mine = first + __popc(need & leaders_below);
        if(mine < seq_end) { alive = true; pos_b = (PosT)mine; pos_a = first_rank; }
      }
      if(first + wanted >= seq_end) { exhausted = true; }
    }

    unsigned active = __ballot_sync(FULL, alive);
    if(active == 0) { break; }

    // Both record reads of this step are issued first: neither address depends on the other's data.
    uint4 qb = make_uint4(0, 0, 0, 0), qa = make_uint4(0, 0, 0, 0);
    if(alive)
    {
      qb = load_chunk_64B(records_b + 4 * (size_t)(pos_b >> RECORD_SHIFT));
      qa = load_chunk_64B(records_a + 4 * (size_t)(pos_a >> RECORD_SHIFT));
    }

    // Emit the rank of the current suffix (fmi.cpp:290 with a run of length 1) while the loads fly.
    if(alive && sub == 0) { stage[fill + __popc(active & leaders_below)] = (KeyT)pos_a; }
    fill += __popc(active & LEADERS);
    if(fill > WALK_STAGE - 8)
    {
      __syncwarp();
      unsigned long long base = 0;
      if(lane == 0) { base = atomicAdd(cursor, (unsigned long long)fill); }
      base = __shfl_sync(FULL, base, 0);
      if(base + fill <= capacity)
      {
        for(uint32_t k = lane; k < fill; k += 32)
        {
          KeyT value = stage[k];
          out[base + k] = value;
          if(counting) { atomicAdd(&digit_counts[(unsigned int)(value >> histogram.shift) & digit_mask], 1u); }
          if(counting_fine) { atomicAdd(histogram.fine_counts + (uint64_t)(value >> histogram.fine_shift), 1ull); }
        }
      }
      else
      {
        if(lane == 0) { counters->overflow = 1; }
        alive = false; exhausted = true;
      }
      __syncwarp();
      fill = 0;
      active = __ballot_sync(FULL, alive);
      if(active == 0) { break; }
    }

    if(alive)
    {
      // (c, b') = LF_B(b): FMI::LF(i), fmi.h:147-150;  a' = LF_A(a, c): FMI::LF(i, c), fmi.h:152-155
      uint32_t offset_b = (uint32_t)pos_b & (RECORD_SYMBOLS - 1), offset_a = (uint32_t)pos_a & (RECORD_SYMBOLS - 1);
      uint32_t t = offset_b & 31u;
      uint32_t mine = ((qb.x >> t) & 1u) | (((qb.y >> t) & 1u) << 1) | (((qb.z >> t) & 1u) << 2);
      uint32_t comp = __shfl_sync(active, mine, group_base + (int)(offset_b >> 5));
      uint32_t safe = (comp == 0 ? 1u : comp);
      PosT super_b = (PosT)__ldg(b.super + (size_t)(pos_b >> SUPER_SHIFT) * SUPER_STRIDE + safe);
      PosT super_a = (PosT)__ldg(a.super + (size_t)(pos_a >> SUPER_SHIFT) * SUPER_STRIDE + safe);
      PosT next_b = c_b[safe] + super_b + coop_record_rank(qb, offset_b, safe, active, sub, group_base);
      PosT next_a = c_a[safe] + super_a + coop_record_rank(qa, offset_a, safe, active, sub, group_base);
      if(comp == 0) { alive = false; }
      else { pos_b = next_b; pos_a = next_a; }
    }
  }

  if(fill > 0)
  {
    __syncwarp();
    unsigned long long base = 0;
    if(lane == 0) { base = atomicAdd(cursor, (unsigned long long)fill); }
    base = __shfl_sync(FULL, base, 0);
    if(base + fill <= capacity)
    {
      for(uint32_t k = lane; k < fill; k += 32)
        {
          KeyT value = stage[k];
          out[base + k] = value;
          if(counting) { atomicAdd(&digit_counts[(unsigned int)(value >> histogram.shift) & digit_mask], 1u); }
          if(counting_fine) { atomicAdd(histogram.fine_counts + (uint64_t)(value >> histogram.fine_shift), 1ull); }
        }
    }
    else if(lane == 0) { counters->overflow = 1; }
  }
  if(counting)
  {
    __syncthreads();
    for(unsigned int d = threadIdx.x; d < histogram.bins; d += WALK_THREADS)
    {
      if(digit_counts[d] != 0) { atomicAdd(histogram.counts + d, (unsigned long long)digit_counts[d]); }
    }
  }
}

template<class KeyT, class PosT>
static int launch_coop(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                       KeyT* d_out, uint64_t capacity, WalkCounters* counters, unsigned long long* cursor,
                       int sms, int max_blocks_per_sm, cudaStream_t stream, WalkHistogram histogram)
{
  int per_sm = 0;
  BWTM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_walk_coop<KeyT, PosT>, WALK_THREADS, 0));
  if(per_sm < 1) { per_sm = 1; }
  if(max_blocks_per_sm > 0 && per_sm > max_blocks_per_sm) { per_sm = max_blocks_per_sm; }
  uint64_t sequences = seq_last + 1 - seq_first;
  uint64_t blocks = std::min((uint64_t)sms * per_sm, div_up(sequences, WALK_THREADS / COOP_LANES));
  k1_walk_coop<KeyT, PosT><<<(unsigned)blocks, WALK_THREADS, 0, stream>>>(
    device_view(a), device_view(b), seq_first, seq_last + 1, d_out, capacity, counters, cursor, histogram);
  return BWTM_OK;
}

// Enqueues one walk over the sequences [seq_first, seq_last] on `stream` without waiting for it.
// `counters` must be zeroed; RA values are appended at *cursor (shared by consecutive launches).
template<class KeyT>
int walk_sequences_async(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                         KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor,
                         int max_blocks_per_sm, cudaStream_t stream, const WalkHistogram* histogram)
{
  WalkHistogram counting = { nullptr, 0, 1, nullptr, 0 };
  if(histogram != nullptr && (histogram->fine_counts != nullptr || (histogram->counts != nullptr && histogram->bins <= 1024))) { counting = *histogram; if(counting.counts == nullptr) { counting.bins = 1; } }
  int device = 0, sms = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  BWTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if(walk_uses_pairs(a, b))   // two backward steps per record read (bwtm_pairs.cu)
  {
    return walk_pairs_async<KeyT>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, stream, &counting);
  }
  if(a->size < 0xFFFFFFFFull && b->size < 0xFFFFFFFFull && getenv("BWTM_FORCE_WIDE") == nullptr)
  {
    BWTM_TRY((launch_coop<KeyT, uint32_t>(a, b, seq_first, seq_last, d_out, capacity, static_cast<WalkCounters*>(counters), cursor, sms, max_blocks_per_sm, stream, counting)));
  }
  else
  {
    BWTM_TRY((launch_coop<KeyT, uint64_t>(a, b, seq_first, seq_last, d_out, capacity, static_cast<WalkCounters*>(counters), cursor, sms, max_blocks_per_sm, stream, counting)));
  }
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

template int walk_sequences_async<uint32_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint32_t*, uint64_t, void*, unsigned long long*, int, cudaStream_t, const WalkHistogram*);
template int walk_sequences_async<uint64_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint64_t*, uint64_t, void*, unsigned long long*, int, cudaStream_t, const WalkHistogram*);

uint64_t walk_counters_bytes() { return sizeof(WalkCounters); }

int walk_counters_check(const void* host_copy)
{
  const WalkCounters* c = static_cast<const WalkCounters*>(host_copy);
  return c->overflow;
}

template<class KeyT>
int walk_sequences(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                   KeyT* d_out, uint64_t capacity, uint64_t* emitted, cudaStream_t stream, const WalkHistogram* histogram)
{
  DeviceBuffer counters; BWTM_TRY(counters.allocate(sizeof(WalkCounters)));
  BWTM_CUDA(cudaMemsetAsync(counters.ptr, 0, sizeof(WalkCounters), stream));

  unsigned long long* cursor = &(counters.as<WalkCounters>()->emitted);
  BWTM_TRY(walk_sequences_async<KeyT>(a, b, seq_first, seq_last, d_out, capacity, counters.ptr, cursor, 0, stream, histogram));

  WalkCounters host;
  BWTM_CUDA(cudaMemcpyAsync(&host, counters.ptr, sizeof(WalkCounters), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  *emitted = host.emitted;
  if(host.overflow) { set_error("rank array buffer too small: %llu values for capacity %llu",
                                (unsigned long long)host.emitted, (unsigned long long)capacity); return BWTM_ERR_CAPACITY; }
  return BWTM_OK;
}

template int walk_sequences<uint32_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint32_t*, uint64_t, uint64_t*, cudaStream_t, const WalkHistogram*);
template int walk_sequences<uint64_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint64_t*, uint64_t, uint64_t*, cudaStream_t, const WalkHistogram*);

// K2: sort of the RA values (support.h:421 sequentialSort + the merge cascade fmi.cpp:220-257).
// Only the low `bits` bits are sorted.  The result is in `d_keys` or `d_alt`; returns which.
//
// cub's onesweep kernel is latency-bound on B200 (18 % of DRAM bandwidth, 37 % occupancy: profiles/
// r01_cub_sort_tuning_sweep*.txt); for 4-byte keys a larger tile (384 threads x 24 keys instead of the
// library's 384 x 19) is the best of the swept configurations (29.8 vs 33.4 ms for 1.51 G keys).
struct TunedSortHub
{
  using KeyT = uint32_t; using ValueT = cub::NullType; using OffsetT = unsigned long long;
  static constexpr bool KEYS_ONLY = true;
  using DominantT = KeyT;
  struct Policy : cub::ChainedPolicy<1000, Policy, Policy>
  {
    static constexpr bool ONESWEEP = true;
    static constexpr int ONESWEEP_RADIX_BITS = 8;
    static constexpr int PRIMARY_RADIX_BITS = 7, SINGLE_TILE_RADIX_BITS = 6, SEGMENTED_RADIX_BITS = 6;
    using HistogramPolicy    = cub::AgentRadixSortHistogramPolicy<128, 16, 1, KeyT, ONESWEEP_RADIX_BITS>;
    using ExclusiveSumPolicy = cub::AgentRadixSortExclusiveSumPolicy<256, ONESWEEP_RADIX_BITS>;
    using OnesweepPolicy = cub::AgentRadixSortOnesweepPolicy<384, 24, DominantT, 1, cub::RADIX_RANK_MATCH_EARLY_COUNTS_ANY,
                                                             cub::BLOCK_SCAN_RAKING_MEMOIZE, cub::RADIX_SORT_STORE_DIRECT, ONESWEEP_RADIX_BITS>;
    // Not launched when ONESWEEP is set, but the dispatch instantiates them.
    using ScanPolicy = cub::AgentScanPolicy<512, 23, OffsetT, cub::BLOCK_LOAD_WARP_TRANSPOSE, cub::LOAD_DEFAULT, cub::BLOCK_STORE_WARP_TRANSPOSE, cub::BLOCK_SCAN_RAKING_MEMOIZE>;
    using DownsweepPolicy = cub::AgentRadixSortDownsweepPolicy<512, 23, DominantT, cub::BLOCK_LOAD_TRANSPOSE, cub::LOAD_DEFAULT, cub::RADIX_RANK_MATCH, cub::BLOCK_SCAN_WARP_SCANS, PRIMARY_RADIX_BITS>;
    using AltDownsweepPolicy = cub::AgentRadixSortDownsweepPolicy<256, 47, DominantT, cub::BLOCK_LOAD_TRANSPOSE, cub::LOAD_DEFAULT, cub::RADIX_RANK_MEMOIZE, cub::BLOCK_SCAN_WARP_SCANS, PRIMARY_RADIX_BITS - 1>;
    using UpsweepPolicy    = cub::AgentRadixSortUpsweepPolicy<256, 23, DominantT, cub::LOAD_DEFAULT, PRIMARY_RADIX_BITS>;
    using AltUpsweepPolicy = cub::AgentRadixSortUpsweepPolicy<256, 47, DominantT, cub::LOAD_DEFAULT, PRIMARY_RADIX_BITS - 1>;
    using SingleTilePolicy = cub::AgentRadixSortDownsweepPolicy<256, 19, DominantT, cub::BLOCK_LOAD_DIRECT, cub::LOAD_LDG, cub::RADIX_RANK_MEMOIZE, cub::BLOCK_SCAN_WARP_SCANS, SINGLE_TILE_RADIX_BITS>;
    using SegmentedPolicy = cub::AgentRadixSortDownsweepPolicy<192, 39, DominantT, cub::BLOCK_LOAD_TRANSPOSE, cub::LOAD_DEFAULT, cub::RADIX_RANK_MEMOIZE, cub::BLOCK_SCAN_WARP_SCANS, SEGMENTED_RADIX_BITS>;
    using AltSegmentedPolicy = cub::AgentRadixSortDownsweepPolicy<384, 11, DominantT, cub::BLOCK_LOAD_TRANSPOSE, cub::LOAD_DEFAULT, cub::RADIX_RANK_MEMOIZE, cub::BLOCK_SCAN_WARP_SCANS, SEGMENTED_RADIX_BITS - 1>;
  };
  using MaxPolicy = Policy;
};

static cudaError_t radix_sort_dispatch(void* temp, size_t& bytes, cub::DoubleBuffer<uint32_t>& keys, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream)
{
  cub::DoubleBuffer<cub::NullType> values;
  return cub::DispatchRadixSort<false, uint32_t, cub::NullType, unsigned long long, TunedSortHub>::Dispatch(
    temp, bytes, keys, values, (unsigned long long)n, begin_bit, end_bit, true, stream);
}

static cudaError_t radix_sort_dispatch(void* temp, size_t& bytes, cub::DoubleBuffer<uint64_t>& keys, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream)
{
  return cub::DeviceRadixSort::SortKeys(temp, bytes, keys, (int64_t)n, begin_bit, end_bit, stream);
}

template<class KeyT>
static int radix_sort_bits(cub::DoubleBuffer<KeyT>& buffers, uint64_t n, int begin_bit, int end_bit, cudaStream_t stream)
{
  size_t temp_bytes = 0;
  BWTM_CUDA(radix_sort_dispatch(nullptr, temp_bytes, buffers, n, begin_bit, end_bit, stream));
  DeviceBuffer temp; BWTM_TRY(temp.allocate(temp_bytes));
  BWTM_CUDA(radix_sort_dispatch(temp.ptr, temp_bytes, buffers, n, begin_bit, end_bit, stream));
  count_launch((uint64_t)(2 + (end_bit - begin_bit + 7) / 8));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  return BWTM_OK;
}

// The low bits of the keys are not radix-sorted. RA values are spread over the positions of A about as
// evenly as the inserted suffixes are, so after sorting by the high bits every range of 2^local_bits A
// positions holds a moderate number of keys that only have to be COUNTED: a CTA builds the histogram of
// its range in shared memory and writes every value as many times as it occurred. That replaces two
// passes over the keys (of four for a 31-bit key) by one that is bound by plain streaming.
constexpr int LOCAL_THREADS = 1024;

__device__ __forceinline__ uint32_t padded(uint32_t v) { return v + (v >> 5); }   // a thread's 32 counters on 32 banks

// offsets[r] = number of keys whose high part is below r, r in [0, ranges].
template<class KeyT>
__global__ void range_offsets(const KeyT* __restrict__ keys, uint64_t n, int local_bits, uint64_t ranges, unsigned long long* __restrict__ offsets)
{
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(r > ranges) { return; }
  uint64_t lo = 0, hi = n;
  while(lo < hi)
  {
    uint64_t mid = lo + (hi - lo) / 2;
    if(((uint64_t)keys[mid] >> local_bits) < r) { lo = mid + 1; } else { hi = mid; }
  }
  offsets[r] = lo;
}

// Ranges with more than `limit` keys (every inserted sequence contributes the value |sequences of A|, so there
// always is one) are listed as (begin, end) pairs after the counter in heavy[0].
constexpr int MAX_HEAVY_RANGES = 64;

__global__ void range_heavy(const unsigned long long* __restrict__ offsets, uint64_t ranges, unsigned long long small_keys,
                            unsigned long long limit, unsigned long long* __restrict__ heavy)
{
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= ranges) { return; }
  unsigned long long begin = offsets[r], end = offsets[r + 1];
  if(end - begin > small_keys && end - begin <= limit) { atomicAdd(heavy + 1 + 2 * MAX_HEAVY_RANGES, 1ull); }   // wide-counter ranges
  if(end - begin <= limit) { return; }
  unsigned long long slot = atomicAdd(heavy, 1ull);
  if(slot < (unsigned long long)MAX_HEAVY_RANGES) { heavy[1 + 2 * slot] = begin; heavy[2 + 2 * slot] = end; }
}

// Ranges of at most SMALL_RANGE_KEYS keys (nearly all of them): 16-bit counters, two per word. The sorted low
// parts are laid out in shared memory without a loop over the copies of a value: the owner of a value
// writes it (plus one) at the first output slot of the value, and at the start of every 32-slot row the
// value reaches into; a row then is one ballot and one shuffle away from its final contents.
constexpr uint32_t SMALL_RANGE_KEYS = 65535;
constexpr uint32_t SMALL_RANGE_ROWS = SMALL_RANGE_KEYS / 32 + 1;

__device__ __forceinline__ uint32_t padded_word(uint32_t w) { return w + (w >> 4); }   // a thread's 16 words on 16 banks, odd stride

__device__ __forceinline__ void place_value(uint16_t* staged, uint16_t* row_first, uint32_t value, uint32_t count, uint32_t& position)
{
  if(count == 0) { return; }
  staged[position] = (uint16_t)(value + 1);
  const uint32_t last_row = (position + count - 1) >> 5;
  uint32_t row = (position + 31) >> 5;
  if(row <= last_row) { row_first[row] = (uint16_t)(value + 1); }   // a value of a few copies starts at most one row
#pragma unroll 1
  for(row++; row <= last_row; row++) { row_first[row] = (uint16_t)(value + 1); }
  position += count;
}

template<class KeyT>
__global__ void __launch_bounds__(LOCAL_THREADS)
local_counting_sort_small(const KeyT* __restrict__ in, KeyT* __restrict__ out, const unsigned long long* __restrict__ offsets, int local_bits,
                          unsigned long long small_keys)
{
  extern __shared__ uint32_t shared_words[];
  __shared__ uint32_t warp_sums[LOCAL_THREADS / 32];
  const uint64_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
  if(hi == lo || hi - lo > small_keys) { return; }
  const uint32_t values = 1u << local_bits, words = values / 2, words_per_thread = words / LOCAL_THREADS;
  uint32_t* counters = shared_words;                                                  // padded_word(words) words
  uint32_t* staged_words = shared_words + padded_word(words);
  uint16_t* staged = reinterpret_cast<uint16_t*>(staged_words);                       // SMALL_RANGE_KEYS + 1 entries
  uint16_t* row_first = staged + (SMALL_RANGE_KEYS + 1);                              // SMALL_RANGE_ROWS entries
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t total = (uint32_t)(hi - lo);
  for(uint32_t w = tid; w < padded_word(words); w += LOCAL_THREADS) { counters[w] = 0; }
  for(uint32_t w = tid; w < (total + 1) / 2; w += LOCAL_THREADS) { staged_words[w] = 0; }
  __syncthreads();
  // independent loads first, then the shared-memory atomics
  const KeyT* mine_in = in + lo + tid;
  uint32_t k = tid;
  for(; k + 7 * LOCAL_THREADS < total; k += 8 * LOCAL_THREADS, mine_in += 8 * LOCAL_THREADS)   // full batches: no bounds checks
  {
    uint32_t low[8];
#pragma unroll
    for(int u = 0; u < 8; u++) { low[u] = (uint32_t)mine_in[u * LOCAL_THREADS] & (values - 1u); }
#pragma unroll
    for(int u = 0; u < 8; u++) { atomicAdd(&counters[padded_word(low[u] >> 1)], 1u << (16 * (low[u] & 1u))); }
  }
  {
    uint32_t low[8];
#pragma unroll
    for(int u = 0; u < 8; u++) { low[u] = (k + u * LOCAL_THREADS < total ? (uint32_t)mine_in[u * LOCAL_THREADS] & (values - 1u) : 0xFFFFFFFFu); }
#pragma unroll
    for(int u = 0; u < 8; u++) { if(low[u] != 0xFFFFFFFFu) { atomicAdd(&counters[padded_word(low[u] >> 1)], 1u << (16 * (low[u] & 1u))); } }
  }
  __syncthreads();

  const uint32_t first_word = tid * words_per_thread;
  uint32_t sum = 0;
  for(uint32_t i = 0; i < words_per_thread; i++) { uint32_t pair = counters[padded_word(first_word + i)]; sum += (pair & 0xFFFFu) + (pair >> 16); }
  uint32_t inclusive = sum;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= (uint32_t)offset) { inclusive += other; }
  }
  if(lane == 31) { warp_sums[warp] = inclusive; }
  __syncthreads();
  if(warp == 0)
  {
    uint32_t w = warp_sums[lane], scanned = w;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      uint32_t other = __shfl_up_sync(0xFFFFFFFFu, scanned, offset);
      if(lane >= (uint32_t)offset) { scanned += other; }
    }
    warp_sums[lane] = scanned - w;
  }
  __syncthreads();
  uint32_t position = warp_sums[warp] + inclusive - sum;
  for(uint32_t i = 0; i < words_per_thread; i++)
  {
    uint32_t pair = counters[padded_word(first_word + i)];
    uint32_t value = 2 * (first_word + i);
    place_value(staged, row_first, value, pair & 0xFFFFu, position);
    place_value(staged, row_first, value + 1, pair >> 16, position);
  }
  __syncthreads();

  const KeyT high = (KeyT)blockIdx.x << local_bits;
  const uint32_t rows = (total + 31) >> 5;
  for(uint32_t row = warp; row < rows; row += LOCAL_THREADS / 32)
  {
    const uint32_t o = 32 * row + lane;
    uint32_t x = (o < total ? (uint32_t)staged[o] : 0u);
    if(lane == 0) { x = row_first[row]; }
    const uint32_t heads = __ballot_sync(0xFFFFFFFFu, x != 0) & ((2u << lane) - 1u);   // bit 0 is always set
    const uint32_t mine = __shfl_sync(0xFFFFFFFFu, x, 31 - __clz(heads));
    if(o < total) { out[lo + o] = high | (KeyT)(mine - 1u); }
  }
}

template<class KeyT>
__global__ void __launch_bounds__(LOCAL_THREADS)
local_counting_sort(const KeyT* __restrict__ in, KeyT* __restrict__ out, const unsigned long long* __restrict__ offsets, int local_bits,
                    unsigned long long small_keys, unsigned long long limit)
{
  extern __shared__ uint32_t counters[];   // padded(1 << local_bits) words
  __shared__ uint32_t warp_sums[LOCAL_THREADS / 32];
  const uint64_t lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
  if(hi == lo || hi - lo <= small_keys || hi - lo > limit) { return; }   // small ranges: local_counting_sort_small
  const uint32_t values = 1u << local_bits, per_thread = values / LOCAL_THREADS;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for(uint32_t v = tid; v < padded(values); v += LOCAL_THREADS) { counters[v] = 0; }
  __syncthreads();
  for(uint64_t k = lo + tid; k < hi; k += LOCAL_THREADS) { atomicAdd(&counters[padded((uint32_t)in[k] & (values - 1u))], 1u); }
  __syncthreads();

  // counts -> exclusive prefixes, in place: thread t owns the values [t per_thread, (t + 1) per_thread)
  const uint32_t first = tid * per_thread;
  uint32_t sum = 0;
  for(uint32_t i = 0; i < per_thread; i++) { sum += counters[padded(first + i)]; }
  uint32_t inclusive = sum;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= (uint32_t)offset) { inclusive += other; }
  }
  if(lane == 31) { warp_sums[warp] = inclusive; }
  __syncthreads();
  if(warp == 0)
  {
    uint32_t w = warp_sums[lane], scanned = w;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      uint32_t other = __shfl_up_sync(0xFFFFFFFFu, scanned, offset);
      if(lane >= (uint32_t)offset) { scanned += other; }
    }
    warp_sums[lane] = scanned - w;
  }
  __syncthreads();
  uint32_t running = warp_sums[warp] + inclusive - sum;
  for(uint32_t i = 0; i < per_thread; i++)
  {
    uint32_t count = counters[padded(first + i)];
    counters[padded(first + i)] = running;
    running += count;
  }
  __syncthreads();

  // output position o holds the last value whose prefix is <= o
  const KeyT high = (KeyT)blockIdx.x << local_bits;
  const uint64_t total = hi - lo;
  for(uint64_t o = tid; o < total; o += LOCAL_THREADS)
  {
    uint32_t a = 0, b = values;
    while(b - a > 1)
    {
      uint32_t middle = (a + b) >> 1;
      if((uint64_t)counters[padded(middle)] <= o) { a = middle; } else { b = middle; }
    }
    out[lo + o] = high | (KeyT)a;
  }
}

//------------------------------------------------------------------------------
// High key bits: most-significant-digit partition passes.
//
// The counting pass above finishes the low bits of every range on its own, so the high bits only have to be
// PARTITIONED, not sorted stably: an MSD pass may place the keys of a bucket in any order. That removes what makes
// a general radix pass expensive (the chained scan that gives every tile its place in every bucket): a CTA
// histograms its tile, claims room in each bucket with one atomic add per non-empty bin, and writes its keys bucket
// by bucket from shared memory. Two passes of up to 10 bits replace the library's onesweep passes
// (latency-bound at 18 % of DRAM bandwidth on B200, profiles/r01_cub_sort_tuning_sweep*.txt).

constexpr int MSD_THREADS  = 512;
constexpr int MSD_ITEMS    = 16;
constexpr int MSD_TILE     = MSD_THREADS * MSD_ITEMS;   // 8192 keys
constexpr int MSD_MAX_BINS = 1024;

// Segment s of a level holds the keys [bounds[s], bounds[s + 1]) and is cut into tiles; tile_first[s] is the index of
// its first tile (tile_first[segments] = number of tiles). Returns false for a CTA beyond the last tile.
__device__ __forceinline__ bool msd_locate_tile(const unsigned long long* __restrict__ bounds, const unsigned int* __restrict__ tile_first,
                                                unsigned int segments, unsigned int tile, unsigned int& segment, uint64_t& begin, uint64_t& end)
{
  if(tile >= tile_first[segments]) { return false; }
  unsigned int lo = 0, hi = segments;   // last segment with tile_first <= tile
  while(hi - lo > 1)
  {
    unsigned int mid = (lo + hi) >> 1;
    if(tile_first[mid] <= tile) { lo = mid; } else { hi = mid; }
  }
  segment = lo;
  begin = bounds[lo] + (uint64_t)(tile - tile_first[lo]) * MSD_TILE;
  end = bounds[lo + 1];
  if(end > begin + MSD_TILE) { end = begin + MSD_TILE; }
  return true;
}

// counts[segment * bins + digit] += number of keys of the tile with that digit.
template<class KeyT>
__global__ void __launch_bounds__(MSD_THREADS)
msd_histogram(const KeyT* __restrict__ keys, const unsigned long long* __restrict__ bounds, const unsigned int* __restrict__ tile_first,
              unsigned int segments, int shift, unsigned int bins, unsigned long long* __restrict__ counts)
{
  __shared__ unsigned int histogram[MSD_MAX_BINS];
  __shared__ unsigned int s_segment; __shared__ unsigned long long s_begin, s_end; __shared__ int s_valid;
  if(threadIdx.x == 0)
  {
    unsigned int segment = 0; uint64_t begin = 0, end = 0;
    s_valid = msd_locate_tile(bounds, tile_first, segments, blockIdx.x, segment, begin, end) ? 1 : 0;
    s_segment = segment; s_begin = begin; s_end = end;
  }
  for(unsigned int d = threadIdx.x; d < bins; d += MSD_THREADS) { histogram[d] = 0; }
  __syncthreads();
  if(!s_valid) { return; }
  const uint64_t begin = s_begin, end = s_end;
  const unsigned int mask = bins - 1;
  for(uint64_t k = begin + threadIdx.x; k < end; k += MSD_THREADS) { atomicAdd(&histogram[(unsigned int)(keys[k] >> shift) & mask], 1u); }
  __syncthreads();
  for(unsigned int d = threadIdx.x; d < bins; d += MSD_THREADS)
  {
    if(histogram[d] != 0) { atomicAdd(&counts[(uint64_t)s_segment * bins + d], (unsigned long long)histogram[d]); }
  }
}

// One CTA per segment: counts -> cursors (absolute output index of the next key of every bin) and the segment table
// of the next level (sub-segment segment * bins + digit): bounds and tiles.
__global__ void __launch_bounds__(MSD_MAX_BINS)
msd_cursors(const unsigned long long* __restrict__ counts, const unsigned long long* __restrict__ bounds, unsigned int bins,
            unsigned long long* __restrict__ cursors, unsigned long long* __restrict__ next_bounds, unsigned int* __restrict__ next_tiles)
{
  __shared__ unsigned long long warp_totals[MSD_MAX_BINS / 32];
  const unsigned int d = threadIdx.x, lane = d & 31, warp = d >> 5;
  const uint64_t slot = (uint64_t)blockIdx.x * bins + d;
  unsigned long long mine = (d < bins ? counts[slot] : 0), inclusive = mine;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    unsigned long long other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= (unsigned int)offset) { inclusive += other; }
  }
  if(lane == 31) { warp_totals[warp] = inclusive; }
  __syncthreads();
  if(warp == 0)
  {
    unsigned long long value = (lane < blockDim.x / 32 ? warp_totals[lane] : 0), scanned = value;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      unsigned long long other = __shfl_up_sync(0xFFFFFFFFu, scanned, offset);
      if(lane >= (unsigned int)offset) { scanned += other; }
    }
    if(lane < blockDim.x / 32) { warp_totals[lane] = scanned - value; }
  }
  __syncthreads();
  if(d >= bins) { return; }
  unsigned long long start = bounds[blockIdx.x] + warp_totals[warp] + inclusive - mine;
  cursors[slot] = start;
  next_bounds[slot] = start;
  next_tiles[slot] = (unsigned int)((mine + MSD_TILE - 1) / MSD_TILE);   // scanned by msd_tile_scan
  if(blockIdx.x == gridDim.x - 1 && d == bins - 1) { next_bounds[slot + 1] = start + mine; }
}

// In-place exclusive scan of the tile counts of a level (entries + 1 values; the last one becomes the total).
__global__ void __launch_bounds__(1024)
msd_tile_scan(unsigned int* __restrict__ tiles, uint64_t entries)
{
  __shared__ unsigned int warp_totals[32];
  __shared__ unsigned int carry;
  if(threadIdx.x == 0) { carry = 0; }
  __syncthreads();
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for(uint64_t base = 0; base < entries; base += 1024)
  {
    uint64_t k = base + threadIdx.x;
    unsigned int mine = (k < entries ? tiles[k] : 0), inclusive = mine;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      unsigned int other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
      if(lane >= (unsigned int)offset) { inclusive += other; }
    }
    if(lane == 31) { warp_totals[warp] = inclusive; }
    __syncthreads();
    if(warp == 0)
    {
      unsigned int value = warp_totals[lane], scanned = value;
#pragma unroll
      for(int offset = 1; offset < 32; offset <<= 1)
      {
        unsigned int other = __shfl_up_sync(0xFFFFFFFFu, scanned, offset);
        if(lane >= (unsigned int)offset) { scanned += other; }
      }
      warp_totals[lane] = scanned - value;
    }
    __syncthreads();
    unsigned int exclusive = carry + warp_totals[warp] + inclusive - mine;
    if(k < entries) { tiles[k] = exclusive; }
    __syncthreads();
    if(threadIdx.x == 1023) { carry = exclusive + mine; }
    __syncthreads();
  }
  if(threadIdx.x == 0) { tiles[entries] = carry; }
}

// The partition pass: keys of a tile go to their buckets in `out`, in any order within a bucket.
template<class KeyT>
__global__ void __launch_bounds__(MSD_THREADS, 3)
msd_scatter(const KeyT* __restrict__ in, KeyT* __restrict__ out, const unsigned long long* __restrict__ bounds,
            const unsigned int* __restrict__ tile_first, unsigned int segments, int shift, unsigned int bins,
            unsigned long long* __restrict__ cursors)
{
  extern __shared__ __align__(16) unsigned char msd_shared[];
  KeyT* staged = reinterpret_cast<KeyT*>(msd_shared);   // MSD_TILE keys
  __shared__ unsigned int histogram[MSD_MAX_BINS];      // count, then the bin's next free slot in `staged`
  __shared__ unsigned long long target[MSD_MAX_BINS];   // output index of the bin's first key minus its first slot
  __shared__ unsigned int warp_totals[MSD_THREADS / 32];
  __shared__ unsigned int s_segment; __shared__ unsigned long long s_begin, s_end; __shared__ int s_valid;
  if(threadIdx.x == 0)
  {
    unsigned int segment = 0; uint64_t begin = 0, end = 0;
    s_valid = msd_locate_tile(bounds, tile_first, segments, blockIdx.x, segment, begin, end) ? 1 : 0;
    s_segment = segment; s_begin = begin; s_end = end;
  }
  for(unsigned int d = threadIdx.x; d < bins; d += MSD_THREADS) { histogram[d] = 0; }
  __syncthreads();
  if(!s_valid) { return; }
  const uint64_t begin = s_begin;
  const unsigned int count = (unsigned int)(s_end - s_begin), mask = bins - 1;
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  KeyT keys[MSD_ITEMS];
#pragma unroll
  for(int i = 0; i < MSD_ITEMS; i++)
  {
    unsigned int k = threadIdx.x + i * MSD_THREADS;
    keys[i] = (k < count ? in[begin + k] : (KeyT)0);
  }
#pragma unroll
  for(int i = 0; i < MSD_ITEMS; i++)
  {
    unsigned int k = threadIdx.x + i * MSD_THREADS;
    if(k < count) { atomicAdd(&histogram[(unsigned int)(keys[i] >> shift) & mask], 1u); }
  }
  __syncthreads();

  // Room in the buckets (one atomic per non-empty bin) and the bins' places in the staging area: exclusive scan of
  // the counts, two bins per thread (bins <= 2 * MSD_THREADS).
  {
    unsigned int d0 = 2 * threadIdx.x, d1 = d0 + 1;
    unsigned int c0 = (d0 < bins ? histogram[d0] : 0), c1 = (d1 < bins ? histogram[d1] : 0);
    unsigned int sum = c0 + c1, inclusive = sum;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      unsigned int other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
      if(lane >= (unsigned int)offset) { inclusive += other; }
    }
    if(lane == 31) { warp_totals[warp] = inclusive; }
    __syncthreads();
    if(warp == 0)
    {
      unsigned int value = (lane < MSD_THREADS / 32 ? warp_totals[lane] : 0), scanned = value;
#pragma unroll
      for(int offset = 1; offset < 32; offset <<= 1)
      {
        unsigned int other = __shfl_up_sync(0xFFFFFFFFu, scanned, offset);
        if(lane >= (unsigned int)offset) { scanned += other; }
      }
      if(lane < MSD_THREADS / 32) { warp_totals[lane] = scanned - value; }
    }
    __syncthreads();
    unsigned int first0 = warp_totals[warp] + inclusive - sum, first1 = first0 + c0;
    unsigned long long* row = cursors + (uint64_t)s_segment * bins;
    if(d0 < bins) { histogram[d0] = first0; if(c0 != 0) { target[d0] = atomicAdd(row + d0, (unsigned long long)c0) - first0; } }
    if(d1 < bins) { histogram[d1] = first1; if(c1 != 0) { target[d1] = atomicAdd(row + d1, (unsigned long long)c1) - first1; } }
  }
  __syncthreads();
#pragma unroll
  for(int i = 0; i < MSD_ITEMS; i++)
  {
    unsigned int k = threadIdx.x + i * MSD_THREADS;
    if(k < count) { staged[atomicAdd(&histogram[(unsigned int)(keys[i] >> shift) & mask], 1u)] = keys[i]; }   // the bin's next slot
  }
  __syncthreads();
  for(unsigned int k = threadIdx.x; k < count; k += MSD_THREADS)
  {
    KeyT key = staged[k];
    out[target[(unsigned int)(key >> shift) & mask] + k] = key;
  }
}

// Partitions keys[0, n) by the bits [low_bit, high_bit) with one or two MSD passes (at most 20 bits). On return
// *where holds the partitioned keys (d_keys or d_alt) and d_offsets[r] (r = 0 .. 2^(high_bit - low_bit)) is the index
// of the first key whose bits are >= r: the ranges of the counting pass.
// sums[row] = sum of the `width` counters of a row (level-1 counts from the histogram of all partitioned bits).
__global__ void __launch_bounds__(256)
msd_row_sums(const unsigned long long* __restrict__ fine, unsigned int width, unsigned long long* __restrict__ sums)
{
  __shared__ unsigned long long warp_totals[8];
  unsigned long long mine = 0;
  for(unsigned int d = threadIdx.x; d < width; d += 256) { mine += fine[(uint64_t)blockIdx.x * width + d]; }
#pragma unroll
  for(int offset = 16; offset > 0; offset >>= 1) { mine += __shfl_xor_sync(0xFFFFFFFFu, mine, offset); }
  if((threadIdx.x & 31) == 0) { warp_totals[threadIdx.x >> 5] = mine; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    unsigned long long total = 0;
    for(int w = 0; w < 8; w++) { total += warp_totals[w]; }
    sums[blockIdx.x] = total;
  }
}

static void msd_geometry(int low_bit, int high_bit, int* levels, int* bits1, int* bits2)
{
  const int total_bits = high_bit - low_bit;
  *levels = (total_bits > 10 ? 2 : 1);
  *bits1 = (*levels == 1 ? total_bits : (total_bits + 1) / 2);
  *bits2 = total_bits - *bits1;
}

template<class KeyT>
static int msd_partition(KeyT* d_keys, KeyT* d_alt, uint64_t n, int low_bit, int high_bit, unsigned long long* d_offsets,
                         KeyT** where, cudaStream_t stream, const WalkHistogram* walked)
{
  const unsigned long long* level1_counts = (walked != nullptr ? walked->counts : nullptr);
  const unsigned long long* fine_counts = (walked != nullptr ? walked->fine_counts : nullptr);
  const int total_bits = high_bit - low_bit;
  int levels = 1, bits1 = total_bits, bits2 = 0;
  msd_geometry(low_bit, high_bit, &levels, &bits1, &bits2);
  const unsigned int bins1 = 1u << bits1, bins2 = 1u << bits2;
  const uint64_t ranges = 1ull << total_bits;

  // level tables: bounds (entries + 1), tile_first (entries + 1), counts / cursors (entries x bins)
  DeviceBuffer bounds1, tiles1, counts1, cursors1, bounds2, tiles2, counts2;
  BWTM_TRY(bounds1.allocate(2 * sizeof(unsigned long long))); BWTM_TRY(tiles1.allocate(2 * sizeof(unsigned int)));
  BWTM_TRY(counts1.allocate(bins1 * sizeof(unsigned long long))); BWTM_TRY(cursors1.allocate(bins1 * sizeof(unsigned long long)));
  BWTM_TRY(bounds2.allocate((bins1 + 1) * sizeof(unsigned long long))); BWTM_TRY(tiles2.allocate((bins1 + 1) * sizeof(unsigned int)));
  const unsigned long long host_bounds[2] = { 0, n };
  const unsigned int tiles_level1 = (unsigned int)div_up(n, MSD_TILE);
  const unsigned int host_tiles[2] = { 0, tiles_level1 };
  BWTM_CUDA(cudaMemcpyAsync(bounds1.ptr, host_bounds, sizeof(host_bounds), cudaMemcpyHostToDevice, stream));
  BWTM_CUDA(cudaMemcpyAsync(tiles1.ptr, host_tiles, sizeof(host_tiles), cudaMemcpyHostToDevice, stream));

  // level 1 (the most significant bits); its histogram may have been collected by the walk
  const int shift1 = low_bit + bits2;
  if(fine_counts != nullptr)   // all partitioned bits were counted by the walk: level 1 is the sum of its rows
  {
    msd_row_sums<<<bins1, 256, 0, stream>>>(fine_counts, bins2, counts1.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
  }
  else if(level1_counts != nullptr)
  {
    BWTM_CUDA(cudaMemcpyAsync(counts1.ptr, level1_counts, bins1 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
  }
  else
  {
    BWTM_CUDA(cudaMemsetAsync(counts1.ptr, 0, bins1 * sizeof(unsigned long long), stream));
    msd_histogram<KeyT><<<tiles_level1, MSD_THREADS, 0, stream>>>(d_keys, bounds1.as<unsigned long long>(), tiles1.as<unsigned int>(), 1, shift1, bins1,
                                                                   counts1.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
  }
  unsigned long long* level1_bounds = (levels == 1 ? d_offsets : bounds2.as<unsigned long long>());
  msd_cursors<<<1, std::max(32u, bins1), 0, stream>>>(counts1.as<unsigned long long>(), bounds1.as<unsigned long long>(), bins1,
                                                      cursors1.as<unsigned long long>(), level1_bounds, tiles2.as<unsigned int>());
  BWTM_LAUNCH_CHECK();
  const size_t staged_bytes = (size_t)MSD_TILE * sizeof(KeyT);
  BWTM_CUDA(cudaFuncSetAttribute(msd_scatter<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_bytes));
  msd_scatter<KeyT><<<tiles_level1, MSD_THREADS, staged_bytes, stream>>>(d_keys, d_alt, bounds1.as<unsigned long long>(), tiles1.as<unsigned int>(), 1, shift1, bins1,
                                                               cursors1.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  if(levels == 1) { *where = d_alt; return BWTM_OK; }

  // level 2: every bucket of level 1 is a segment
  msd_tile_scan<<<1, 1024, 0, stream>>>(tiles2.as<unsigned int>(), bins1);
  BWTM_LAUNCH_CHECK();
  const unsigned int tiles_level2 = tiles_level1 + bins1;   // upper bound: every segment wastes less than one tile
  const unsigned long long* level2_counts = fine_counts;
  if(level2_counts == nullptr)
  {
    BWTM_TRY(counts2.allocate(ranges * sizeof(unsigned long long)));
    BWTM_CUDA(cudaMemsetAsync(counts2.ptr, 0, ranges * sizeof(unsigned long long), stream));
    msd_histogram<KeyT><<<tiles_level2, MSD_THREADS, 0, stream>>>(d_alt, bounds2.as<unsigned long long>(), tiles2.as<unsigned int>(), bins1, low_bit, bins2,
                                                                   counts2.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
    level2_counts = counts2.as<unsigned long long>();
  }
  // The cursors of level 2 are written over the counts; their initial values are the range offsets.
  DeviceBuffer cursors2, unused_tiles;
  BWTM_TRY(cursors2.allocate(ranges * sizeof(unsigned long long))); BWTM_TRY(unused_tiles.allocate((ranges + 1) * sizeof(unsigned int)));
  msd_cursors<<<bins1, std::max(32u, bins2), 0, stream>>>(level2_counts, bounds2.as<unsigned long long>(), bins2,
                                                          cursors2.as<unsigned long long>(), d_offsets, unused_tiles.as<unsigned int>());
  BWTM_LAUNCH_CHECK();
  msd_scatter<KeyT><<<tiles_level2, MSD_THREADS, staged_bytes, stream>>>(d_alt, d_keys, bounds2.as<unsigned long long>(), tiles2.as<unsigned int>(), bins1, low_bit, bins2,
                                                               cursors2.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  *where = d_keys;
  return BWTM_OK;
}

static uint64_t env_number(const char* name, uint64_t fallback)
{
  const char* text = getenv(name);
  return (text == nullptr ? fallback : (uint64_t)strtoull(text, nullptr, 10));
}

// How sort_keys treats n keys of `bits` bits below key_limit: plain radix sort of everything, or the high bits
// partitioned / radix-sorted and the low `local_bits` bits counted range by range.
struct SortPlan
{
  bool     counting;      // the counting route
  bool     msd;           // ... with MSD partition passes for the high bits
  int      local_bits;
  uint64_t key_limit;
};

static SortPlan sort_plan(uint64_t n, int bits, uint64_t key_limit)
{
  SortPlan plan;
  if(key_limit == 0 || (bits < 64 && key_limit > (1ull << bits))) { key_limit = (bits < 64 ? 1ull << bits : ~0ull); }
  plan.key_limit = key_limit;
  // BWTM_LOCAL_SORT_MIN: smallest input that takes the counting route (tests force it with 1).
  const uint64_t local_min = env_number("BWTM_LOCAL_SORT_MIN", 1ull << 22);
  plan.local_bits = std::min(15, std::max(12, bits - 16));
  const int local_bits = plan.local_bits;
  // The counting pass costs about the same per range whatever the range holds: it pays off when the ranges
  // are well filled (a rank of a multi-GPU merge sorts only its share of the keys over the same positions).
  const uint64_t expected_ranges = (bits > local_bits ? std::min<uint64_t>(1ull << std::min(bits - local_bits, 40), ((key_limit - 1) >> local_bits) + 1) : 1);
  const uint64_t local_density = env_number("BWTM_LOCAL_SORT_DENSITY", 8192);
  const uint64_t max_density = env_number("BWTM_LOCAL_SORT_MAX_DENSITY", 1ull << 19);   // 0: no upper bound (tests)
  plan.counting = !(n < local_min || bits <= local_bits || bits - local_bits > 22 || n / expected_ranges < local_density ||
                    (max_density > 0 && n / expected_ranges > max_density));   // nearly every range would be a heavy one (a small A under a large B)
  plan.msd = (plan.counting && bits - local_bits <= 20 && env_number("BWTM_MSD", 1) != 0);
  return plan;
}

bool sort_plan_histogram(uint64_t n, int bits, uint64_t key_limit, WalkHistogram* histogram, uint64_t* fine_bins)
{
  SortPlan plan = sort_plan(n, bits, key_limit);
  histogram->counts = nullptr; histogram->fine_counts = nullptr; histogram->shift = 0; histogram->bins = 1; histogram->fine_shift = 0;
  *fine_bins = 0;
  if(!plan.msd) { return false; }
  int levels = 1, bits1 = 0, bits2 = 0;
  msd_geometry(plan.local_bits, bits, &levels, &bits1, &bits2);
  histogram->shift = plan.local_bits + bits2; histogram->bins = 1u << bits1;
  // By default only the first level's digit is counted by the walk (in shared memory). BWTM_FINE_HISTOGRAM=1: all
  // partitioned bits, with global reductions -- measured: +4 ms in the walk for -1.4 ms in the sort (config 2), so off.
  if(levels == 2 && env_number("BWTM_FINE_HISTOGRAM", 0) != 0) { histogram->fine_shift = plan.local_bits; *fine_bins = 1ull << (bits - plan.local_bits); }
  return true;
}

template<class KeyT>
int sort_keys(KeyT* d_keys, KeyT* d_alt, uint64_t n, int bits, KeyT** sorted, cudaStream_t stream, uint64_t key_limit,
              const WalkHistogram* walked)
{
  const SortPlan plan = sort_plan(n, bits, key_limit);
  key_limit = plan.key_limit;
  cub::DoubleBuffer<KeyT> buffers(d_keys, d_alt);
  // BWTM_LOCAL_SORT_LIMIT: most keys one range may hold before the plain radix sort takes over.
  const uint64_t local_limit = env_number("BWTM_LOCAL_SORT_LIMIT", 1ull << 19);
  const uint64_t small_keys = std::min<uint64_t>(env_number("BWTM_LOCAL_SORT_SMALL", SMALL_RANGE_KEYS), SMALL_RANGE_KEYS);   // tests: 0
  const int local_bits = plan.local_bits;
  if(!plan.counting)
  {
  BWTM_TRY(radix_sort_bits<KeyT>(buffers, n, 0, bits, stream));
    *sorted = buffers.Current();
    return BWTM_OK;
  }

  const uint64_t ranges = std::min<uint64_t>(1ull << (bits - local_bits), ((key_limit - 1) >> local_bits) + 1);
  DeviceBuffer offsets; BWTM_TRY(offsets.allocate(((1ull << (bits - local_bits)) + 1) * sizeof(unsigned long long)));
  DeviceBuffer heavy; BWTM_TRY(heavy.allocate((2 + 2 * MAX_HEAVY_RANGES) * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemsetAsync(heavy.ptr, 0, (2 + 2 * MAX_HEAVY_RANGES) * sizeof(unsigned long long), stream));
  // The high bits: MSD partition passes (they also deliver the range offsets), or the library's radix passes
  // (BWTM_MSD=0, and keys with more than 20 high bits).
  if(plan.msd)
  {
    KeyT* where = nullptr;
    BWTM_TRY(msd_partition<KeyT>(d_keys, d_alt, n, local_bits, bits, offsets.as<unsigned long long>(), &where, stream, walked));
    if(where != buffers.Current()) { buffers.selector ^= 1; }
  }
  else
  {
    BWTM_TRY(radix_sort_bits<KeyT>(buffers, n, local_bits, bits, stream));
    range_offsets<KeyT><<<(unsigned)div_up(ranges + 1, 256), 256, 0, stream>>>(buffers.Current(), n, local_bits, ranges, offsets.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
  }
  range_heavy<<<(unsigned)div_up(ranges, 256), 256, 0, stream>>>(offsets.as<unsigned long long>(), ranges, small_keys, local_limit, heavy.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  const uint32_t half_words = 1u << (local_bits - 1);
  const size_t small_bytes = (size_t)(half_words + (half_words >> 4)) * sizeof(uint32_t) + (size_t)(SMALL_RANGE_KEYS + 1 + SMALL_RANGE_ROWS) * sizeof(uint16_t);
  BWTM_CUDA(cudaFuncSetAttribute(local_counting_sort_small<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_bytes));
  local_counting_sort_small<KeyT><<<(unsigned)ranges, LOCAL_THREADS, small_bytes, stream>>>(buffers.Current(), buffers.Alternate(),
                                                                                           offsets.as<unsigned long long>(), local_bits, small_keys);
  BWTM_LAUNCH_CHECK();
  unsigned long long heavy_host[2 + 2 * MAX_HEAVY_RANGES];
  BWTM_CUDA(cudaMemcpyAsync(heavy_host, heavy.ptr, sizeof(heavy_host), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  if(heavy_host[1 + 2 * MAX_HEAVY_RANGES] > 0 && heavy_host[0] <= (unsigned long long)MAX_HEAVY_RANGES)
  {
    const size_t shared_bytes = (size_t)((1u << local_bits) + (1u << (local_bits - 5))) * sizeof(uint32_t);
    BWTM_CUDA(cudaFuncSetAttribute(local_counting_sort<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared_bytes));
    local_counting_sort<KeyT><<<(unsigned)ranges, LOCAL_THREADS, shared_bytes, stream>>>(buffers.Current(), buffers.Alternate(), offsets.as<unsigned long long>(),
                                                                                         local_bits, small_keys, local_limit);
    BWTM_LAUNCH_CHECK();
  }
  if(heavy_host[0] > (unsigned long long)MAX_HEAVY_RANGES)   // the keys pile up in many ranges: plain radix sort of everything
  {
    BWTM_TRY(radix_sort_bits<KeyT>(buffers, n, 0, bits, stream));
    *sorted = buffers.Current();
    return BWTM_OK;
  }
  // The few heavy ranges (typically long stretches of one value) get a radix sort of their low bits.
  DeviceBuffer temp;
  for(unsigned long long k = 0; k < heavy_host[0]; k++)
  {
    unsigned long long begin = heavy_host[1 + 2 * k], count = heavy_host[2 + 2 * k] - begin;
    size_t temp_bytes = 0;
    BWTM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, buffers.Current() + begin, buffers.Alternate() + begin, (int64_t)count, 0, local_bits, stream));
    if(temp_bytes > temp.bytes) { BWTM_TRY(temp.allocate(temp_bytes)); }
    BWTM_CUDA(cub::DeviceRadixSort::SortKeys(temp.ptr, temp_bytes, buffers.Current() + begin, buffers.Alternate() + begin, (int64_t)count, 0, local_bits, stream));
    count_launch((uint64_t)(2 + (local_bits + 7) / 8));
  }
  BWTM_CUDA(cudaStreamSynchronize(stream));
  *sorted = buffers.Alternate();
  return BWTM_OK;
}

template int sort_keys<uint32_t>(uint32_t*, uint32_t*, uint64_t, int, uint32_t**, cudaStream_t, uint64_t, const WalkHistogram*);
template int sort_keys<uint64_t>(uint64_t*, uint64_t*, uint64_t, int, uint64_t**, cudaStream_t, uint64_t, const WalkHistogram*);

} // namespace bwtm
