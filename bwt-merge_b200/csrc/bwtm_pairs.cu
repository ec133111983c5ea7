// Two-step LF: the pair records and the rank-array walk that uses them.
//
// buildRA (fmi.cpp:272-334) spends its time in two dependent random reads per inserted base:
//     (c, b') = LF_B(b)      FMI::LF(i),    fmi.h:147-150 -> BWT::inverse_select, bwt.cpp:445-464
//     a'      = LF_A(a, c)   FMI::LF(i, c), fmi.h:152-155 -> BWT::rank,           bwt.cpp:318-341
// and on B200 such reads are limited by the NUMBER of line requests (about 39 G/s), not by their size:
// 32-, 64- and 128-byte records cost the same (profiles/r01_random_line_ceiling_coop_chase.txt).  So a record
// twice as large that answers TWO backward steps halves the time of the walk.
//
// For every position i of a BWT let c1 = BWT[i] and c2 = BWT[LF(i)] (c2 = 0 when c1 is the endmarker).
// With pairrank(i, c1 c2) = #{k < i : BWT[k] = c1 and BWT[LF(k)] = c2},
//     LF(LF(i, c1), c2) = C[c2] + rank(C[c1], c2) + pairrank(i, c1 c2)
// because the positions below LF(i, c1) = C[c1] + rank(i, c1) that hold c2 are those below C[c1] (a constant
// per pair) plus the images of the k < i with BWT[k] = c1.  The intermediate position LF(i, c1), which the rank
// array needs as well, is the ordinary single-symbol rank.
//
// Pair record: 128 bytes per 64 positions, read by four lanes with two 16-byte loads each (lane l owns words
// 8 l .. 8 l + 7):
//     words  0..2   bit planes of c1, positions 0..31       words  8..10  the same, positions 32..63
//     words  3..5   bit planes of c2, positions 0..31       words 11..13  the same, positions 32..63
//     words  6, 7, 14, 15   128-bit field of five 25-bit counters: #c1 in [start of the 2^25 superblock, 64 r)
//     words 16..31  twenty-five 20-bit counters: #(c1 c2) in [start of the 2^20 pair superblock, 64 r), both >= 1
// plus one row of 32 u64 per pair superblock: entry 5 (c1 - 1) + (c2 - 1) = C[c2] + rank(C[c1], c2) + #(c1 c2)
// before the superblock, i.e. the complete two-step LF value up to the superblock.  Single counts use the
// superblock table of the basic records.
//
// Building it is a streaming pass: for a fixed c the k-th occurrence of c maps to C[c] + k, so a thread that
// owns 32 consecutive positions reads its c2 values from (at most a few) consecutive chunks per symbol.
#include <algorithm>
#include <cstring>

#include "bwtm_merge.cuh"

namespace bwtm
{

constexpr int PAIR_SHIFT        = 6;                      // positions per pair record = 64
constexpr int PAIR_WORDS        = 32;
constexpr int PAIR_SUPER_SHIFT  = 20;                     // positions per pair superblock
constexpr int PAIR_SUPER_STRIDE = 32;                     // u64 per row
constexpr int PAIR_RECORDS_PER_SUPER = 1 << (PAIR_SUPER_SHIFT - PAIR_SHIFT);   // 16384
constexpr uint32_t PAIR_FIELD_MASK = (1u << 20) - 1;

//------------------------------------------------------------------------------
// Builder

// Bit deposit ("expand", the inverse of compress; Hacker's Delight 7-5): the low popc(mask) bits of x go to the set
// positions of mask, in order. The five move masks depend on the mask only and serve all three bit planes.
struct DepositNetwork { uint32_t move[5]; uint32_t mask; };

__device__ __forceinline__ DepositNetwork deposit_network(uint32_t m)
{
  DepositNetwork net; net.mask = m;
  uint32_t mk = ~m << 1;
#pragma unroll
  for(int i = 0; i < 5; i++)
  {
    uint32_t mp = mk ^ (mk << 1); mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
    uint32_t mv = mp & m;
    net.move[i] = mv;
    m = (m ^ mv) | (mv >> (1u << i));
    mk &= ~mp;
  }
  return net;
}

__device__ __forceinline__ uint32_t deposit(const DepositNetwork& net, uint32_t x)
{
#pragma unroll
  for(int i = 4; i >= 0; i--) { uint32_t mv = net.move[i]; x = (x & ~mv) | ((x << (1u << i)) & mv); }
  return x & net.mask;
}

// One thread per 32-position chunk of the basic records: c1 planes copied, c2 planes by following LF. For a symbol
// c the occurrences in the chunk map to CONSECUTIVE positions j, j + 1, ... (LF is order-preserving per symbol), so
// their c2 values are a window of up to 32 bits of each plane starting at j, and a bit deposit puts them at the
// chunk positions that hold c. No loop over occurrences: every thread does the same work.
__global__ void __launch_bounds__(256)
pairs_gather(DeviceIndex idx, uint64_t n_chunks, uint32_t* __restrict__ pair_words)
{
  const unsigned FULL = 0xFFFFFFFFu;
  uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, sub = lane & 3, group_base = lane & ~3;
  // The grid covers whole records (4 chunks): threads beyond the end take part in the shuffles with empty chunks.
  uint4 q = make_uint4(0, 0, 0, 0);
  const uint64_t basic_chunks = ((idx.size >> RECORD_SHIFT) + 1) * 4;
  if(chunk < basic_chunks) { q = __ldg(idx.records + chunk); }
  uint32_t h0 = __shfl_sync(FULL, q.w, group_base), h1 = __shfl_sync(FULL, q.w, group_base + 1);
  uint32_t h2 = __shfl_sync(FULL, q.w, group_base + 2), h3 = __shfl_sync(FULL, q.w, group_base + 3);

  // Occurrences of every symbol in the earlier chunks of the record: one byte per symbol, scanned over the 4 lanes.
  uint64_t packed = 0;
#pragma unroll
  for(uint32_t c = 1; c < SIGMA; c++) { packed |= (uint64_t)__popc(match_mask(q, c)) << (8 * (c - 1)); }
  uint64_t before = packed;
  {
    uint64_t other = __shfl_up_sync(FULL, before, 1, 4); if(sub >= 1) { before += other; }
    other = __shfl_up_sync(FULL, before, 2, 4);          if(sub >= 2) { before += other; }
  }
  before -= packed;
  if(chunk >= n_chunks) { return; }

  const uint64_t* super_row = idx.super + ((chunk >> 2) >> SUPER_RECORD_SHIFT) * SUPER_STRIDE;
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for(uint32_t c = 1; c < SIGMA; c++)
  {
    const uint32_t m = match_mask(q, c);
    if(m == 0) { continue; }
    const uint64_t j = idx.C[c] + __ldg(super_row + c) + header_field(h0, h1, h2, h3, c) + ((before >> (8 * (c - 1))) & 0xFFu);
    const uint32_t shift = (uint32_t)j & 31u;
    const uint4 lo = __ldg(idx.records + (j >> 5));
    uint4 hi = make_uint4(0, 0, 0, 0);
    if(shift + __popc(m) > 32) { hi = __ldg(idx.records + (j >> 5) + 1); }
    const DepositNetwork net = deposit_network(m);
    t0 |= deposit(net, __funnelshift_r(lo.x, hi.x, shift));
    t1 |= deposit(net, __funnelshift_r(lo.y, hi.y, shift));
    t2 |= deposit(net, __funnelshift_r(lo.z, hi.z, shift));
  }
  uint32_t* words = pair_words + (chunk >> 1) * PAIR_WORDS + (chunk & 1) * 8;
  *reinterpret_cast<uint4*>(words) = make_uint4(q.x, q.y, q.z, t0);
  *reinterpret_cast<uint2*>(words + 4) = make_uint2(t1, t2);
}

// The 25 pair counts of one pair record (both halves), added to counts[].
__device__ __forceinline__ void pair_record_counts(const uint32_t* __restrict__ words, uint32_t counts[25])
{
#pragma unroll
  for(int half = 0; half < 2; half++)
  {
    uint4 lo = *reinterpret_cast<const uint4*>(words + 8 * half);
    uint2 hi = *reinterpret_cast<const uint2*>(words + 8 * half + 4);
    uint4 first = make_uint4(lo.x, lo.y, lo.z, 0), second = make_uint4(lo.w, hi.x, hi.y, 0);
    uint32_t m2[SIGMA];
#pragma unroll
    for(uint32_t c = 1; c < SIGMA; c++) { m2[c] = match_mask(second, c); }
#pragma unroll
    for(uint32_t c1 = 1; c1 < SIGMA; c1++)
    {
      uint32_t m1 = match_mask(first, c1);
#pragma unroll
      for(uint32_t c2 = 1; c2 < SIGMA; c2++) { counts[5 * (c1 - 1) + (c2 - 1)] += __popc(m1 & m2[c2]); }
    }
  }
}

constexpr int PAIR_FILL_THREADS = 512;
constexpr int PAIR_FILL_WARPS   = PAIR_FILL_THREADS / 32;                           // 16
constexpr int PAIR_FILL_ROUNDS  = PAIR_RECORDS_PER_SUPER / PAIR_FILL_THREADS;       // 32 rounds of 32 records per warp

// Counters of every pair record, one CTA per pair superblock: the 25 pair counters relative to the superblock, the
// five single counters (relative to the 2^25 superblock) from the header of the basic record, and the superblock's
// totals for the table. Warp w owns the records [1024 w, 1024 w + 1024) of the superblock; in a round its lanes take
// 32 consecutive records, so that loads and stores of a round cover 4 KB of contiguous memory. Two phases: totals of
// the warps first (their exclusive scan gives every warp its base), then the same counts again, scanned over the
// lanes of every round (two 16-bit classes per word), packed and stored.
__global__ void __launch_bounds__(PAIR_FILL_THREADS)
pairs_fill(DeviceIndex idx, uint32_t* __restrict__ pair_words, uint64_t n_pair_records, unsigned long long* __restrict__ totals)
{
  __shared__ uint32_t warp_totals[25][PAIR_FILL_WARPS];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warp_first = (uint64_t)blockIdx.x * PAIR_RECORDS_PER_SUPER + (uint64_t)warp * (32 * PAIR_FILL_ROUNDS);

  uint32_t base[25];
#pragma unroll
  for(int k = 0; k < 25; k++) { base[k] = 0; }
  for(int round = 0; round < PAIR_FILL_ROUNDS; round++)
  {
    uint64_t record = warp_first + 32 * round + lane;
    if(record < n_pair_records) { pair_record_counts(pair_words + record * PAIR_WORDS, base); }
  }
#pragma unroll
  for(int k = 0; k < 25; k++)
  {
    uint32_t v = base[k];
#pragma unroll
    for(int offset = 16; offset > 0; offset >>= 1) { v += __shfl_xor_sync(0xFFFFFFFFu, v, offset); }
    if(lane == 0) { warp_totals[k][warp] = v; }
  }
  __syncthreads();
  if(threadIdx.x < 25)   // exclusive scan over the warps, class by class; the superblock's total goes to the table builder
  {
    uint32_t running = 0;
    for(int w = 0; w < PAIR_FILL_WARPS; w++) { uint32_t v = warp_totals[threadIdx.x][w]; warp_totals[threadIdx.x][w] = running; running += v; }
    totals[(uint64_t)blockIdx.x * 25 + threadIdx.x] = running;
  }
  __syncthreads();
#pragma unroll
  for(int k = 0; k < 25; k++) { base[k] = warp_totals[k][warp]; }

  for(int round = 0; round < PAIR_FILL_ROUNDS; round++)
  {
    const uint64_t record = warp_first + 32 * round + lane;
    const bool valid = (record < n_pair_records);
    uint32_t* words = pair_words + record * PAIR_WORDS;
    uint32_t counts[25];
#pragma unroll
    for(int k = 0; k < 25; k++) { counts[k] = 0; }
    if(valid) { pair_record_counts(words, counts); }
    // inclusive scan over the lanes, two classes per word (a round holds at most 32 * 64 = 2048 of a class)
    uint32_t scanned[13];
#pragma unroll
    for(int i = 0; i < 13; i++) { scanned[i] = counts[2 * i] | (i < 12 ? counts[2 * i + 1] << 16 : 0u); }
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
#pragma unroll
      for(int i = 0; i < 13; i++)
      {
        uint32_t other = __shfl_up_sync(0xFFFFFFFFu, scanned[i], offset);
        if(lane >= (uint32_t)offset) { scanned[i] += other; }
      }
    }
    if(valid)
    {
      // 25 x 20 bits -> words 16..31
      uint32_t out[16];
      uint64_t accumulator = 0; int bits = 0, word = 0;
#pragma unroll
      for(int k = 0; k < 25; k++)
      {
        uint32_t inclusive = (k & 1 ? scanned[k >> 1] >> 16 : scanned[k >> 1] & 0xFFFFu);
        uint32_t prefix = base[k] + inclusive - counts[k];
        accumulator |= (uint64_t)(prefix & PAIR_FIELD_MASK) << bits; bits += 20;
        if(bits >= 32) { out[word++] = (uint32_t)accumulator; accumulator >>= 32; bits -= 32; }
      }
      out[word] = (uint32_t)accumulator;   // bits 480..499 end in word 15
      uint4* destination = reinterpret_cast<uint4*>(words + 16);
      destination[0] = make_uint4(out[0], out[1], out[2], out[3]);    destination[1] = make_uint4(out[4], out[5], out[6], out[7]);
      destination[2] = make_uint4(out[8], out[9], out[10], out[11]);  destination[3] = make_uint4(out[12], out[13], out[14], out[15]);
      // single counters: header of the basic record (+ its first half for the odd pair record)
      const uint4* basic = idx.records + 4 * (record >> 1);
      uint4 q0 = __ldg(basic), q1 = __ldg(basic + 1), q2 = __ldg(basic + 2), q3 = __ldg(basic + 3);
      uint64_t field[SIGMA];
#pragma unroll
      for(uint32_t c = 1; c < SIGMA; c++)
      {
        field[c] = header_field(q0.w, q1.w, q2.w, q3.w, c);
        if(record & 1) { field[c] += __popc(match_mask(q0, c)) + __popc(match_mask(q1, c)); }
      }
      uint64_t lo = field[1] | (field[2] << 25) | (field[3] << 50);
      uint64_t hi = (field[3] >> 14) | (field[4] << 11) | (field[5] << 36);
      *reinterpret_cast<uint2*>(words + 6) = make_uint2((uint32_t)lo, (uint32_t)(lo >> 32));
      *reinterpret_cast<uint2*>(words + 14) = make_uint2((uint32_t)hi, (uint32_t)(hi >> 32));
    }
    // the round's totals (lane 31 holds them) move the warp's base
#pragma unroll
    for(int i = 0; i < 13; i++)
    {
      uint32_t total = __shfl_sync(0xFFFFFFFFu, scanned[i], 31);
      base[2 * i] += total & 0xFFFFu;
      if(i < 12) { base[2 * i + 1] += total >> 16; }
    }
  }
}

// One block: the superblock table (32 u64 per row). Thread k < 25 scans its pair class over the superblocks and
// adds the constant part of the two-step LF value, C[c2] + rank(C[c1], c2).
__global__ void pairs_super(DeviceIndex idx, const unsigned long long* __restrict__ totals, uint64_t n_pair_super, uint64_t* __restrict__ super2)
{
  uint32_t k = threadIdx.x;
  if(k >= PAIR_SUPER_STRIDE) { return; }
  if(k >= 25)
  {
    for(uint64_t sb = 0; sb < n_pair_super; sb++) { super2[sb * PAIR_SUPER_STRIDE + k] = 0; }
    return;
  }
  uint32_t c1 = k / 5 + 1, c2 = k % 5 + 1;
  uint64_t running = idx.C[c2] + rank_nonzero(idx, idx.C[c1], c2);
  for(uint64_t sb = 0; sb < n_pair_super; sb++)
  {
    super2[sb * PAIR_SUPER_STRIDE + k] = running;
    running += totals[sb * 25 + k];
  }
}

uint64_t pair_index_bytes(uint64_t size)
{
  uint64_t n_pair_records = (size >> PAIR_SHIFT) + 1;
  uint64_t n_pair_super = ((n_pair_records - 1) >> (PAIR_SUPER_SHIFT - PAIR_SHIFT)) + 1;
  return n_pair_records * PAIR_WORDS * sizeof(uint32_t) + n_pair_super * PAIR_SUPER_STRIDE * sizeof(uint64_t);
}

// Builds the pair records of `index` (kept with the index until it is destroyed). No-op when they exist.
int ensure_pair_index(bwtm_index* index, cudaStream_t stream, bool built_ahead)
{
  if(index->d_pairs != nullptr) { return BWTM_OK; }
  if(index->d_records == nullptr) { set_error("the index has no rank structure"); return BWTM_ERR_ARGUMENT; }
  const uint64_t n_pair_records = (index->size >> PAIR_SHIFT) + 1;
  const uint64_t n_pair_super = ((n_pair_records - 1) >> (PAIR_SUPER_SHIFT - PAIR_SHIFT)) + 1;
  DeviceBuffer pairs, super2, totals;
  // From the default pool, not from the pool of the other index buffers: the walk reads these gigabytes at random
  // and is sensitive to how they are mapped (config 2: 41.0 ms here, 44.0 ms from the second pool, 39.5 ms from
  // cudaMalloc, whose cost per index the end-to-end path would not get back; profiles/r02_pool_placement.txt). They are
  // built before a merge allocates its work buffers, so they do not cut up the blocks those are reused from.
  // bwtm_index_build_pairs (built_ahead) takes the plain allocation: it is paid once, outside any merge.
  if(built_ahead) { BWTM_TRY(device_alloc_plain(&pairs.ptr, n_pair_records * PAIR_WORDS * sizeof(uint32_t))); pairs.bytes = n_pair_records * PAIR_WORDS * sizeof(uint32_t); }
  else { BWTM_TRY(pairs.allocate(n_pair_records * PAIR_WORDS * sizeof(uint32_t))); }
  BWTM_TRY(super2.allocate(n_pair_super * PAIR_SUPER_STRIDE * sizeof(uint64_t)));
  BWTM_TRY(totals.allocate(n_pair_super * 25 * sizeof(unsigned long long)));
  const uint64_t n_chunks = 2 * n_pair_records;
  const uint64_t grid_chunks = div_up(n_chunks, 4) * 4;
  DeviceIndex view = device_view(index);
  pairs_gather<<<(unsigned)div_up(grid_chunks, 256), 256, 0, stream>>>(view, n_chunks, pairs.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  pairs_fill<<<(unsigned)n_pair_super, PAIR_FILL_THREADS, 0, stream>>>(view, pairs.as<uint32_t>(), n_pair_records, totals.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  pairs_super<<<1, PAIR_SUPER_STRIDE, 0, stream>>>(view, totals.as<unsigned long long>(), n_pair_super, super2.as<uint64_t>());
  BWTM_LAUNCH_CHECK();
  index->pair_bytes = pairs.bytes + super2.bytes;
  index->device_bytes += index->pair_bytes;
  index->n_pair_records = n_pair_records; index->n_pair_super = n_pair_super;
  index->d_pairs = static_cast<uint4*>(pairs.detach());
  index->d_pair_super = static_cast<uint64_t*>(super2.detach());
  return BWTM_OK;
}

void release_pair_index(bwtm_index* index)
{
  if(index->d_pairs == nullptr) { return; }
  device_free(index->d_pairs); device_free(index->d_pair_super);
  index->d_pairs = nullptr; index->d_pair_super = nullptr;
  index->device_bytes -= index->pair_bytes; index->pair_bytes = 0;
}

// BWTM_WALK=single: never build or use pair records; BWTM_WALK=pairs: always (tests); default: when it pays.
static int walk_policy()
{
  const char* env = getenv("BWTM_WALK");
  if(env == nullptr) { return 0; }
  if(env[0] == 's' || env[0] == '1') { return -1; }
  if(env[0] == 'p' || env[0] == '2') { return 1; }
  return 0;
}

bool walk_uses_pairs(const bwtm_index* a, const bwtm_index* b)
{
  return (a->d_pairs != nullptr && b->d_pairs != nullptr && walk_policy() >= 0);
}

// The decision of prepare_walk for indexes of n_a and n_b symbols of which `build` symbols still lack pair records,
// when this GPU walks `walked_bases` of b.
static bool pairs_pay(uint64_t n_a, uint64_t n_b, uint64_t build, uint64_t walked_bases, uint64_t basic_bytes)
{
  const int policy = walk_policy();
  if(policy != 0) { return (policy > 0); }
  // Building the records is a streaming pass over an index (a few ps per symbol); the walk saves about four times
  // that per inserted base. An index that already has them only costs the other side's pass.
  if(build > 4 * walked_bases) { return false; }
  // 2 bytes per symbol on top of the basic records: not when that would crowd out the rank array itself (two key
  // buffers for the bases this GPU walks) or take more than a third of the device.
  const uint64_t total_bytes = device_total_bytes();
  const uint64_t pair_bytes = pair_index_bytes(n_a) + pair_index_bytes(n_b);
  const uint64_t key_bytes = 2 * walked_bases * (n_a < 0xFFFFFFFFull ? 4 : 8);
  if(total_bytes > 0 && (pair_bytes > total_bytes / 3 || pair_bytes + key_bytes + 2 * basic_bytes > total_bytes - total_bytes / 8)) { return false; }
  return true;
}

// For bwtm_index_create_pair: the pair records of the first input, built while the second one is still uploading,
// when a merge of the two would build them anyway. Failure is not an error (the merge decides again).
void build_pairs_ahead(bwtm_index* a, uint64_t expected_b_size, cudaStream_t stream)
{
  if(a->d_pairs != nullptr || expected_b_size == 0) { return; }
  const uint64_t basic_bytes = a->device_bytes + (uint64_t)((double)a->device_bytes * ((double)expected_b_size / (double)std::max<uint64_t>(a->size, 1)));
  if(!pairs_pay(a->size, expected_b_size, a->size + expected_b_size, expected_b_size, basic_bytes)) { return; }
  if(ensure_pair_index(a, stream) != BWTM_OK) { release_pair_index(a); cudaGetLastError(); }
}

int prepare_walk(bwtm_index* a, bwtm_index* b, uint64_t walked_bases, cudaStream_t stream, bwtm_timings* timings)
{
  const int policy = walk_policy();
  const uint64_t build = (a->d_pairs == nullptr ? a->size : 0) + (b->d_pairs == nullptr ? b->size : 0);
  bool want = pairs_pay(a->size, b->size, build, walked_bases, a->device_bytes + b->device_bytes - a->pair_bytes - b->pair_bytes);
  if(want)
  {
    EventTimer timer(stream);
    const bool building = (a->d_pairs == nullptr || b->d_pairs == nullptr);
    if(building) { timer.start(); }
    int rc = ensure_pair_index(a, stream);
    if(rc == BWTM_OK) { rc = ensure_pair_index(b, stream); }
    if(building) { timings->pair_index_seconds = timer.stop() * 1e-3; }
    if(rc == BWTM_ERR_MEMORY && policy == 0) { release_pair_index(a); release_pair_index(b); cudaGetLastError(); }   // single-step walk instead
    else if(rc != BWTM_OK) { return rc; }
  }
  if(walk_uses_pairs(a, b))
  {
    timings->walk_record_bytes = 128;
    timings->walk_table_bytes = a->pair_bytes + b->pair_bytes;
  }
  else
  {
    timings->walk_record_bytes = 64;
    timings->walk_table_bytes = (a->n_records + b->n_records) * 64;
  }
  return BWTM_OK;
}

//------------------------------------------------------------------------------
// Device view and the two-step LF

struct PairView
{
  const uint4*    records;   // 8 uint4 per pair record
  const uint64_t* super1;    // superblock table of the basic records: absolute count of comp c before the superblock
  const uint64_t* super2;    // PAIR_SUPER_STRIDE u64 per pair superblock
  uint64_t        size, sequences;
  uint64_t        C[SIGMA + 1];
};

static PairView pair_view(const bwtm_index* index)
{
  PairView v;
  v.records = index->d_pairs; v.super1 = index->d_super; v.super2 = index->d_pair_super;
  v.size = index->size; v.sequences = index->sequences;
  for(int c = 0; c <= SIGMA; c++) { v.C[c] = index->C[c]; }
  return v;
}

// Reference form of the two-step LF on the pair records, one thread per query (diagnostics and tests):
// c1 = BWT[i], c2 = BWT[LF(i)], first = LF(i), second = LF(LF(i)); positions are 0 where undefined.
__global__ void query_lf2(PairView idx, const uint64_t* __restrict__ positions, uint64_t n,
                          uint64_t* __restrict__ first, uint64_t* __restrict__ second, uint8_t* __restrict__ comps)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) { return; }
  uint64_t i = positions[k];
  if(i >= idx.size) { first[k] = 0; second[k] = 0; comps[2 * k] = 0; comps[2 * k + 1] = 0; return; }
  const uint32_t* words = reinterpret_cast<const uint32_t*>(idx.records) + (i >> PAIR_SHIFT) * PAIR_WORDS;
  uint32_t offset = (uint32_t)i & 63u, half = offset >> 5, t = offset & 31u;
  const uint32_t* planes = words + 8 * half;
  uint32_t c1 = ((planes[0] >> t) & 1u) | (((planes[1] >> t) & 1u) << 1) | (((planes[2] >> t) & 1u) << 2);
  uint32_t c2 = ((planes[3] >> t) & 1u) | (((planes[4] >> t) & 1u) << 1) | (((planes[5] >> t) & 1u) << 2);
  comps[2 * k] = (uint8_t)c1; comps[2 * k + 1] = (uint8_t)c2;
  first[k] = 0; second[k] = 0;
  if(c1 == 0) { return; }
  uint32_t single = 0, pair = 0;
  for(uint32_t h = 0; h <= half; h++)
  {
    const uint32_t* p = words + 8 * h;
    uint32_t limit = (h < half ? 0xFFFFFFFFu : low_mask((int)t));
    uint32_t m1 = match_mask(make_uint4(p[0], p[1], p[2], 0), c1) & limit;
    single += __popc(m1);
    if(c2 != 0) { pair += __popc(m1 & match_mask(make_uint4(p[3], p[4], p[5], 0), c2)); }
  }
  first[k] = idx.C[c1] + idx.super1[(i >> SUPER_SHIFT) * SUPER_STRIDE + c1] + header_field(words[6], words[7], words[14], words[15], c1) + single;
  if(c2 == 0) { return; }
  uint32_t field = 5 * (c1 - 1) + (c2 - 1), bit = 20 * field;
  uint64_t both = (uint64_t)words[16 + (bit >> 5)] | ((uint64_t)words[16 + min((bit >> 5) + 1, 15u)] << 32);
  second[k] = idx.super2[(i >> PAIR_SUPER_SHIFT) * PAIR_SUPER_STRIDE + field] + (uint32_t)((both >> (bit & 31u)) & PAIR_FIELD_MASK) + pair;
}

//------------------------------------------------------------------------------
// K1, two-step form

constexpr int PW_THREADS = 256;
constexpr int PW_WARPS   = PW_THREADS / 32;
constexpr int PW_STAGE   = 256;     // staged RA values per warp
constexpr int PW_LANES   = 4;       // lanes per walker

struct PairWalkCounters
{
  unsigned long long next_sequence;
  unsigned long long emitted;
  int                overflow;
};

__device__ __forceinline__ uint4 load_pair_chunk(const uint4* p)
{
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Complement masks of a comp value: plane ^ mask is the match of one bit plane.
__device__ __forceinline__ void complement_masks(uint32_t c, uint32_t& n0, uint32_t& n1, uint32_t& n2)
{
  n0 = (c & 1u) ? 0u : 0xFFFFFFFFu; n1 = (c & 2u) ? 0u : 0xFFFFFFFFu; n2 = (c & 4u) ? 0u : 0xFFFFFFFFu;
}

// MIN_CTAS resident CTAs per SM bound the registers (5: 51, 6: 42, 7: 36): the kernel waits for DRAM, so more walkers
// in flight are worth more than registers (4 CTAs: 45.5 ms, 5: 39.9 ms on config 2; BWTM_WALK_CTAS selects 5, 6 or 7).
template<class KeyT, class PosT, int MIN_CTAS>
__global__ void __launch_bounds__(PW_THREADS, MIN_CTAS)
k1_walk_pairs(PairView a, PairView b, uint64_t seq_begin, uint64_t seq_end,
              KeyT* __restrict__ out, uint64_t capacity, PairWalkCounters* counters, unsigned long long* cursor, WalkHistogram histogram)
{
  __shared__ KeyT stage_all[PW_WARPS][PW_STAGE];
  __shared__ unsigned int digit_counts[1024];
  __shared__ __align__(16) uint32_t scratch_all[PW_WARPS][32 / PW_LANES][2][16];   // pair counters of A and B per walker
  __shared__ PosT c_a[8];

#pragma unroll
  for(int c = 0; c <= SIGMA; c++) { if(threadIdx.x == c) { c_a[c] = (PosT)a.C[c]; } }
  for(unsigned int d = threadIdx.x; d < 1024; d += PW_THREADS) { digit_counts[d] = 0; }
  __syncthreads();
  const bool counting = (histogram.counts != nullptr), counting_fine = (histogram.fine_counts != nullptr);
  const unsigned int digit_mask = histogram.bins - 1;

  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned LEADERS = 0x11111111u;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (PW_LANES - 1);
  const int group_base = lane & ~(PW_LANES - 1);
  const unsigned leaders_below = LEADERS & ((1u << group_base) - 1u);
  KeyT* stage = stage_all[threadIdx.x >> 5];
  uint32_t* scratch_a = scratch_all[threadIdx.x >> 5][lane >> 2][0];
  uint32_t* scratch_b = scratch_all[threadIdx.x >> 5][lane >> 2][1];
  const PosT first_rank = (PosT)a.sequences;
  const uint4* __restrict__ records_a = a.records + 2 * sub;
  const uint4* __restrict__ records_b = b.records + 2 * sub;

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

    // The pair records of both sides: neither address depends on the other's data.
    uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0, a0 = b0, a1 = b0;
    if(alive)
    {
      const uint4* pb = records_b + 8 * (size_t)(pos_b >> PAIR_SHIFT);
      const uint4* pa = records_a + 8 * (size_t)(pos_a >> PAIR_SHIFT);
      b0 = load_pair_chunk(pb); b1 = load_pair_chunk(pb + 1);
      a0 = load_pair_chunk(pa); a1 = load_pair_chunk(pa + 1);
    }

    // Rank of the current suffix (fmi.cpp:290 with a run of length 1), staged while the loads fly. An iteration
    // stages at most two values per walker.
    if(fill > PW_STAGE - 2 * (32 / PW_LANES))
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
    if(alive && sub == 0) { stage[fill + __popc(active & leaders_below)] = (KeyT)pos_a; }
    fill += __popc(active & LEADERS);

    // Lanes 2 and 3 hold the pair counters: they go to the walker's scratch, where any lane can index them.
    if(sub >= 2)
    {
      uint4* sa = reinterpret_cast<uint4*>(scratch_a + 8 * (sub - 2)); sa[0] = a0; sa[1] = a1;
      uint4* sb = reinterpret_cast<uint4*>(scratch_b + 8 * (sub - 2)); sb[0] = b0; sb[1] = b1;
    }

    const uint32_t offset_b = (uint32_t)pos_b & 63u, offset_a = (uint32_t)pos_a & 63u;
    // (c1, c2) = (BWT_B[b], BWT_B[LF_B(b)]) from the lane that holds b's half of the record
    uint32_t cc;
    {
      uint32_t t = offset_b & 31u;
      uint32_t mine = ((b0.x >> t) & 1u) | (((b0.y >> t) & 1u) << 1) | (((b0.z >> t) & 1u) << 2)
                    | (((b0.w >> t) & 1u) << 3) | (((b1.x >> t) & 1u) << 4) | (((b1.y >> t) & 1u) << 5);
      cc = __shfl_sync(FULL, mine, group_base + (int)(offset_b >> 5));
    }
    const uint32_t c1 = cc & 7u, c2 = cc >> 3;
    const uint32_t s1 = (c1 == 0 ? 1u : c1), s2 = (c2 == 0 ? 1u : c2);   // safe indices for finished walkers
    const uint32_t field = 5u * (s1 - 1u) + (s2 - 1u);

    // Superblock parts of the three LF values (L2-resident tables).
    PosT super_a1 = 0, super_a2 = 0, super_b2 = 0;
    if(alive)
    {
      super_a1 = (PosT)__ldg(a.super1 + (size_t)(pos_a >> SUPER_SHIFT) * SUPER_STRIDE + s1);
      super_a2 = (PosT)__ldg(a.super2 + (size_t)(pos_a >> PAIR_SUPER_SHIFT) * PAIR_SUPER_STRIDE + field);
      super_b2 = (PosT)__ldg(b.super2 + (size_t)(pos_b >> PAIR_SUPER_SHIFT) * PAIR_SUPER_STRIDE + field);
    }

    // In-record counts: lanes 0 and 1 hold 32 positions each.
    uint32_t packed = 0;
    if(sub < 2)
    {
      uint32_t n0, n1, n2, o0, o1, o2;
      complement_masks(s1, n0, n1, n2); complement_masks(s2, o0, o1, o2);
      int ka = (int)offset_a - 32 * sub; ka = (ka < 0 ? 0 : ka);
      int kb = (int)offset_b - 32 * sub; kb = (kb < 0 ? 0 : kb);
      uint32_t ma1 = (a0.x ^ n0) & (a0.y ^ n1) & (a0.z ^ n2) & low_mask(ka);
      uint32_t ma2 = ma1 & (a0.w ^ o0) & (a1.x ^ o1) & (a1.y ^ o2);
      uint32_t mb2 = (b0.x ^ n0) & (b0.y ^ n1) & (b0.z ^ n2) & low_mask(kb) & (b0.w ^ o0) & (b1.x ^ o1) & (b1.y ^ o2);
      packed = __popc(ma1) | (__popc(ma2) << 8) | (__popc(mb2) << 16);
    }
    packed += __shfl_xor_sync(FULL, packed, 1);
    packed += __shfl_xor_sync(FULL, packed, 2);

    // Single counter of A: 25-bit field of the 128-bit field in words 6, 7 (lane 0) and 14, 15 (lane 1).
    uint32_t single_a;
    {
      uint32_t s = 25u * (s1 - 1u), w = s >> 5, shift = s & 31u, w_next = (w < 3 ? w + 1 : 3);
      uint32_t lo = __shfl_sync(FULL, (w & 1u) ? a1.w : a1.z, group_base + (int)(w >> 1));
      uint32_t hi = __shfl_sync(FULL, (w_next & 1u) ? a1.w : a1.z, group_base + (int)(w_next >> 1));
      single_a = __funnelshift_r(lo, hi, shift) & FIELD_MASK;
    }
    // Pair counters of A and B from the scratch.
    __syncwarp();
    uint32_t pair_a, pair_b;
    {
      uint32_t bit = 20u * field, w = bit >> 5, shift = bit & 31u, w_next = (w < 15 ? w + 1 : 15);
      pair_a = __funnelshift_r(scratch_a[w], scratch_a[w_next], shift) & PAIR_FIELD_MASK;
      pair_b = __funnelshift_r(scratch_b[w], scratch_b[w_next], shift) & PAIR_FIELD_MASK;
    }
    __syncwarp();   // the scratch is rewritten in the next iteration

    const PosT middle_a = c_a[s1] + super_a1 + single_a + (packed & 0xFFu);             // LF_A(a, c1)
    const PosT next_a = super_a2 + pair_a + ((packed >> 8) & 0xFFu);                    // LF_A(LF_A(a, c1), c2)
    const PosT next_b = super_b2 + pair_b + (packed >> 16);                             // LF_B(LF_B(b))

    // c1 = $: the sequence is finished. Otherwise the suffix that starts with c1 has rank middle_a.
    if(c1 == 0) { alive = false; }
    unsigned second = __ballot_sync(FULL, alive);
    if(alive && sub == 0) { stage[fill + __popc(second & leaders_below)] = (KeyT)middle_a; }
    fill += __popc(second & LEADERS);
    if(c2 == 0) { alive = false; }
    if(alive) { pos_a = next_a; pos_b = next_b; }
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
    for(unsigned int d = threadIdx.x; d < histogram.bins; d += PW_THREADS)
    {
      if(digit_counts[d] != 0) { atomicAdd(histogram.counts + d, (unsigned long long)digit_counts[d]); }
    }
  }
}

template<class KeyT, class PosT, int MIN_CTAS>
static int launch_pairs_with(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                             KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor, int sms, cudaStream_t stream,
                             WalkHistogram histogram)
{
  int per_sm = 0;
  BWTM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_walk_pairs<KeyT, PosT, MIN_CTAS>, PW_THREADS, 0));
  if(per_sm < 1) { per_sm = 1; }
  uint64_t sequences = seq_last + 1 - seq_first;
  uint64_t blocks = std::min((uint64_t)sms * per_sm, div_up(sequences, PW_THREADS / PW_LANES));
  k1_walk_pairs<KeyT, PosT, MIN_CTAS><<<(unsigned)blocks, PW_THREADS, 0, stream>>>(
    pair_view(a), pair_view(b), seq_first, seq_last + 1, d_out, capacity, static_cast<PairWalkCounters*>(counters), cursor, histogram);
  return BWTM_OK;
}

template<class KeyT, class PosT>
static int launch_pairs(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                        KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor, int sms, cudaStream_t stream,
                        WalkHistogram histogram)
{
  int ctas = 6;
  if(const char* env = getenv("BWTM_WALK_CTAS")) { ctas = atoi(env); }
  if(ctas <= 5) { return launch_pairs_with<KeyT, PosT, 5>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, sms, stream, histogram); }
  if(ctas >= 7) { return launch_pairs_with<KeyT, PosT, 7>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, sms, stream, histogram); }
  return launch_pairs_with<KeyT, PosT, 6>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, sms, stream, histogram);
}

// Enqueues the two-step walk over the sequences [seq_first, seq_last]. Both indexes must have pair records.
// `counters` is walk_counters_bytes() zeroed bytes (same layout as the single-step walk's counters).
template<class KeyT>
int walk_pairs_async(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                     KeyT* d_out, uint64_t capacity, void* counters, unsigned long long* cursor, cudaStream_t stream,
                     const WalkHistogram* histogram)
{
  WalkHistogram counting = { nullptr, 0, 1, nullptr, 0 };
  if(histogram != nullptr && (histogram->fine_counts != nullptr || (histogram->counts != nullptr && histogram->bins <= 1024))) { counting = *histogram; if(counting.counts == nullptr) { counting.bins = 1; } }
  static_assert(sizeof(PairWalkCounters) == 24, "counter layout shared with the single-step walk");
  if(a->d_pairs == nullptr || b->d_pairs == nullptr) { set_error("pair records missing"); return BWTM_ERR_INTERNAL; }
  int device = 0, sms = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  BWTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if(a->size < 0xFFFFFFFFull && b->size < 0xFFFFFFFFull && getenv("BWTM_FORCE_WIDE") == nullptr)
  {
    BWTM_TRY((launch_pairs<KeyT, uint32_t>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, sms, stream, counting)));
  }
  else
  {
    BWTM_TRY((launch_pairs<KeyT, uint64_t>(a, b, seq_first, seq_last, d_out, capacity, counters, cursor, sms, stream, counting)));
  }
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

template int walk_pairs_async<uint32_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint32_t*, uint64_t, void*, unsigned long long*, cudaStream_t, const WalkHistogram*);
template int walk_pairs_async<uint64_t>(const bwtm_index*, const bwtm_index*, uint64_t, uint64_t, uint64_t*, uint64_t, void*, unsigned long long*, cudaStream_t, const WalkHistogram*);

} // namespace bwtm

//------------------------------------------------------------------------------
// C ABI: diagnostics of the pair records

using namespace bwtm;

extern "C"
{

int bwtm_index_build_pairs(bwtm_index* index)
{
  if(index == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(ensure_pair_index(index, 0, true));
  BWTM_CUDA(cudaStreamSynchronize(0));
  return BWTM_OK;
}

int bwtm_lf2(bwtm_index* index, const uint64_t* positions, uint64_t n, uint64_t* out_first, uint64_t* out_second, uint8_t* out_comps)
{
  if(index == nullptr || positions == nullptr || out_first == nullptr || out_second == nullptr || out_comps == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  if(n == 0) { return BWTM_OK; }
  BWTM_TRY(ensure_pair_index(index, 0));
  DeviceBuffer pos, first, second, comps;
  BWTM_TRY(pos.allocate(n * 8)); BWTM_TRY(first.allocate(n * 8)); BWTM_TRY(second.allocate(n * 8)); BWTM_TRY(comps.allocate(2 * n));
  BWTM_CUDA(cudaMemcpy(pos.ptr, positions, n * 8, cudaMemcpyHostToDevice));
  query_lf2<<<(unsigned)div_up(n, 256), 256>>>(pair_view(index), pos.as<uint64_t>(), n, first.as<uint64_t>(), second.as<uint64_t>(), comps.as<uint8_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(out_first, first.ptr, n * 8, cudaMemcpyDeviceToHost));
  BWTM_CUDA(cudaMemcpy(out_second, second.ptr, n * 8, cudaMemcpyDeviceToHost));
  BWTM_CUDA(cudaMemcpy(out_comps, comps.ptr, 2 * n, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

} // extern "C"
