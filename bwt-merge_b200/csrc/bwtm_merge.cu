// K4 interleave, K3 run detection, K5 byte-exact encoder, and the merge entry point.
//
// Replaces mergeBWT (bwt.cpp:215-282) and its RunBuffer + Run::write output path:
//   * the merged sequence M has B[j] at position j + RA[j] and A[i] at position i + #{j : RA[j] <= i}
//     ("before B[j] come RA[j] symbols of A", bwt.cpp:234-261), RA being the sorted rank array;
//   * M is cut into tiles of TILE positions; a merge-path search on the diagonal gives every tile
//     its first A and B index, a shared-memory bitmap marks which positions of the tile come from B,
//     and every thread fetches 16 consecutive symbols from the position-addressed records;
//   * maximal runs of M (what the reference's RunBuffer produces, utils.h:121-142) are found with a
//     device run-length encode;
//   * Run::write (support.h:256-282) is sequential through the output offset modulo 64 only for
//     runs of length >= 42: runs shorter than that always take one byte.  The long runs are
//     compacted, the writer is evaluated for all 64 entry offsets per tile of long runs (a 64-state
//     transducer), the tile maps are composed by a scan, and every run is then written at its
//     exact byte offset.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include <cub/cub.cuh>

#include "bwtm_merge.cuh"
#include "bwtm_batches.cuh"

namespace bwtm
{

constexpr int TILE        = 4096;
constexpr int IL_THREADS  = 256;
constexpr int PER_THREAD  = TILE / IL_THREADS;   // 16
constexpr int LONG_TILE   = 2048;                // long runs per transducer tile (at most 16 bytes each: checkpoints fit 16 bits)
constexpr int LONG_SUB    = 32;                  // long runs per checkpointed sub-tile
constexpr int SCAN_CHUNK  = 128;                 // tile maps staged in shared memory per step of the tile scan

//------------------------------------------------------------------------------
// K4

template<class KeyT>
__global__ void k4_partition(const KeyT* __restrict__ keys, uint64_t key_base, uint64_t key_count,
                             uint64_t begin, uint64_t end, uint64_t tiles, uint64_t* __restrict__ tile_j)
{
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(t > tiles) { return; }
  uint64_t d = begin + t * TILE;
  if(d > end) { d = end; }
  // Number of B symbols placed before merged position d: j + RA[j] is strictly increasing in j.
  uint64_t lo = 0, hi = key_count;
  while(lo < hi)
  {
    uint64_t mid = lo + (hi - lo) / 2;
    if(key_base + mid + (uint64_t)keys[mid] < d) { lo = mid + 1; } else { hi = mid; }
  }
  tile_j[t] = key_base + lo;
}

// Bits [pos, pos + 16) of the three planes of `idx`, packed 16 bits apart; only the chunks that hold the
// `count` symbols actually needed are touched.
__device__ __forceinline__ uint64_t plane_window(const DeviceIndex& idx, uint64_t pos, uint32_t count)
{
  if(count == 0) { return 0; }
  uint64_t chunk = pos >> 5; uint32_t shift = (uint32_t)(pos & 31u);
  uint4 lo = __ldg(idx.records + chunk);
  uint4 hi = make_uint4(0, 0, 0, 0);
  if(shift + count > 32) { hi = __ldg(idx.records + chunk + 1); }
  uint64_t w0 = __funnelshift_r(lo.x, hi.x, shift) & 0xFFFFu;
  uint64_t w1 = __funnelshift_r(lo.y, hi.y, shift) & 0xFFFFu;
  uint64_t w2 = __funnelshift_r(lo.z, hi.z, shift) & 0xFFFFu;
  return w0 | (w1 << 16) | (w2 << 32);
}

template<class KeyT>
__global__ void __launch_bounds__(IL_THREADS)
k4_interleave(DeviceIndex a, DeviceIndex b, const KeyT* __restrict__ keys, uint64_t key_base,
              const uint64_t* __restrict__ tile_j, uint64_t begin, uint64_t end, uint4* __restrict__ merged,
              unsigned long long* __restrict__ distinct_keys)
{
  __shared__ uint32_t bitmap[TILE / 32];
  __shared__ uint32_t prefix[TILE / 32];
  __shared__ uint32_t distinct;
  if(threadIdx.x == 0) { distinct = 0; }

  const int tid = threadIdx.x;
  uint64_t d0 = begin + (uint64_t)blockIdx.x * TILE;
  uint64_t d1 = (d0 + TILE < end ? d0 + TILE : end);
  uint64_t j0 = tile_j[blockIdx.x], j1 = tile_j[blockIdx.x + 1];
  uint64_t i0 = d0 - j0;

  if(tid < TILE / 32) { bitmap[tid] = 0; }
  __syncthreads();
  uint32_t new_values = 0;   // RA values that differ from their predecessor: the reference's RA run count
  {
    // Every thread takes a contiguous share of the tile's keys: their merged positions j + RA[j] increase, so the
    // bits of a share fall into a few consecutive words that are collected in a register and written once each
    // (one atomic per key made the sorted keys of a warp collide on the same word: 292 M bank conflicts per 2^30
    // positions, profiles/r02_k4k5_ncu.txt).
    const uint32_t total = (uint32_t)(j1 - j0), share = (total + IL_THREADS - 1) / IL_THREADS;
    const uint32_t first = min(tid * share, total), last = min(first + share, total);
    if(first < last)
    {
      const KeyT* mine = keys + (j0 - key_base);
      KeyT previous = (j0 + first > key_base ? mine[(int64_t)first - 1] : (KeyT)0);
      bool no_previous = (j0 + first == key_base);
      uint32_t word = 0xFFFFFFFFu, collected = 0;
      for(uint32_t k = first; k < last; k++)
      {
        KeyT key = mine[k];
        uint32_t q = (uint32_t)(j0 + k + (uint64_t)key - d0);
        if((q >> 5) != word)
        {
          if(collected != 0) { atomicOr(&bitmap[word], collected); }
          word = q >> 5; collected = 0;
        }
        collected |= 1u << (q & 31u);
        new_values += (no_previous || key != previous) ? 1u : 0u;
        previous = key; no_previous = false;
      }
      if(collected != 0) { atomicOr(&bitmap[word], collected); }
    }
  }
  if(distinct_keys != nullptr && new_values != 0) { atomicAdd(&distinct, new_values); }
  __syncthreads();
  if(distinct_keys != nullptr && tid == 0 && distinct != 0) { atomicAdd(distinct_keys, (unsigned long long)distinct); }
  if(tid < 32)
  {
    uint32_t local[4], sum = 0;
#pragma unroll
    for(int w = 0; w < 4; w++) { local[w] = sum; sum += __popc(bitmap[tid * 4 + w]); }
    uint32_t inclusive = sum;
#pragma unroll
    for(int offset = 1; offset < 32; offset <<= 1)
    {
      uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
      if(tid >= offset) { inclusive += v; }
    }
    uint32_t exclusive = inclusive - sum;
#pragma unroll
    for(int w = 0; w < 4; w++) { prefix[tid * 4 + w] = exclusive + local[w]; }
  }
  __syncthreads();

  uint32_t p0 = tid * PER_THREAD;
  uint32_t word = p0 >> 5, shift = p0 & 31u;
  uint32_t bits = bitmap[word];
  uint32_t before = prefix[word] + __popc(bits & low_mask((int)shift));
  uint32_t flags = (bits >> shift) & 0xFFFFu;
  uint64_t jb = j0 + before, ia = i0 + p0 - before;
  int valid = 0;
  if(d0 + p0 < d1) { uint64_t left = d1 - d0 - p0; valid = (left < (uint64_t)PER_THREAD ? (int)left : PER_THREAD); }

  // The thread's symbols come from consecutive positions of a (flag 0) and of b (flag 1): a 16-bit window of
  // each source's planes is taken and the bits are dealt out in flag order. The result is the plane-chunk
  // layout of the rank records (bwtm_common.cuh), so the merged sequence is never stored one symbol per byte.
  uint32_t from_b = __popc(flags & low_mask(valid));
  uint64_t window_a = plane_window(a, ia, (uint32_t)valid - from_b), window_b = plane_window(b, jb, from_b);
  uint64_t dealt = 0;
#pragma unroll
  for(int s = 0; s < PER_THREAD; s++)
  {
    uint64_t take_b = (flags >> s) & 1u;
    uint64_t source = (take_b ? window_b : window_a);
    dealt |= (source & 0x0000000100010001ull) << s;
    window_b >>= take_b; window_a >>= (1u - take_b);
  }
  if(valid < PER_THREAD) { dealt &= 0x0000000100010001ull * (uint64_t)low_mask(valid); }
  // Two neighbouring threads hold the halves of one 32-position chunk.
  uint64_t upper = __shfl_down_sync(0xFFFFFFFFu, dealt, 1);
  if((tid & 1) == 0 && valid > 0)
  {
    uint4 chunk;
    chunk.x = (uint32_t)(dealt & 0xFFFFu)         | ((uint32_t)(upper & 0xFFFFu) << 16);
    chunk.y = (uint32_t)((dealt >> 16) & 0xFFFFu) | ((uint32_t)((upper >> 16) & 0xFFFFu) << 16);
    chunk.z = (uint32_t)((dealt >> 32) & 0xFFFFu) | ((uint32_t)((upper >> 32) & 0xFFFFu) << 16);
    chunk.w = 0;
    merged[((d0 - begin) >> 5) + (tid >> 1)] = chunk;
  }
}

//------------------------------------------------------------------------------
// K5: Run::write (support.h:256-282)

__device__ __forceinline__ uint32_t bytecode_length(uint64_t v)   // bytes ByteCode::write emits (support.h:203-212)
{
  uint32_t n = 1;
  while(v > 0x7Fu) { v >>= 7; n++; }
  return n;
}

// Bytes emitted for a run of length >= MAX_RUN that starts at output offset `state` (mod 64).
__device__ __forceinline__ uint32_t long_run_bytes(uint64_t length, uint32_t state)
{
  uint32_t bytes = 0;
  while(length > 0)
  {
    if(length < (uint64_t)MAX_RUN) { bytes++; break; }
    uint32_t remaining = RLE_BLOCK - state;
    uint32_t basic = (remaining > 1 ? MAX_RUN : MAX_RUN - 1);
    length -= basic; bytes++; state = (state + 1) & 63u; remaining--;
    if(remaining > 0)
    {
      uint64_t extension = length;
      uint32_t nb = bytecode_length(extension);
      if(nb > remaining) { extension = (1ull << (7 * remaining)) - 1; nb = remaining; }  // bit_length(length) > 7 * remaining
      length -= extension; bytes += nb; state = (state + nb) & 63u;
    }
  }
  return bytes;
}

// Writes the run at absolute output offset `offset`; returns the bytes written.
// The byte of output offset x goes to out[x - origin] (origin != 0: a staging buffer for part of the output).
__device__ __forceinline__ uint32_t write_run(uint8_t* __restrict__ out, uint64_t offset, uint32_t comp, uint64_t length, uint64_t origin = 0)
{
  uint64_t pos = offset;
  while(length > 0)
  {
    if(length < (uint64_t)MAX_RUN) { out[pos++ - origin] = (uint8_t)(comp + SIGMA * (length - 1)); break; }
    uint32_t remaining = RLE_BLOCK - (uint32_t)(pos & 63u);
    uint32_t basic = (remaining > 1 ? MAX_RUN : MAX_RUN - 1);
    out[pos++ - origin] = (uint8_t)(comp + SIGMA * (basic - 1)); length -= basic; remaining--;
    if(remaining > 0)
    {
      uint64_t extension = length;
      if(bytecode_length(extension) > remaining) { extension = (1ull << (7 * remaining)) - 1; }
      length -= extension;
      while(extension > 0x7Fu) { out[pos++ - origin] = (uint8_t)((extension & 0x7Fu) | 0x80u); extension >>= 7; }
      out[pos++ - origin] = (uint8_t)extension;
    }
  }
  return (uint32_t)(pos - offset);
}

// Sequential glue between slabs (and GPU slices): the RunBuffer state (utils.h:121-142). The pending run
// absorbs the slab's first run when the symbols agree; it is written as soon as the slab shows that it
// has ended; the slab's last run becomes the new pending run. The runs in between, [1, m - 1), never
// depend on the pending run: they are the parallel part.
__global__ void enc_head(EncodeControl* ctl, const SlabEnds* __restrict__ ends, uint8_t* __restrict__ out, int finish)
{
  if(blockIdx.x != 0 || threadIdx.x != 0) { return; }
  ctl->start = 1; ctl->count = 0; ctl->n_short = 0; ctl->n_long = 0; ctl->long_bytes = 0;
  if(finish)
  {
    if(ctl->carry_len > 0)
    {
      ctl->out_size += write_run(out, ctl->out_size, ctl->carry_sym, ctl->carry_len);
      ctl->runs_total++; ctl->carry_len = 0;
    }
    ctl->slab_base = ctl->out_size;
    return;
  }
  const uint64_t m = ends->runs;
  if(m > 0)
  {
    if(ctl->carry_len > 0 && ends->first_sym == ctl->carry_sym) { ctl->carry_len += ends->first_len; }
    else
    {
      if(ctl->carry_len > 0)
      {
        ctl->out_size += write_run(out, ctl->out_size, ctl->carry_sym, ctl->carry_len);
        ctl->runs_total++;
      }
      ctl->carry_sym = ends->first_sym; ctl->carry_len = ends->first_len;
    }
    if(m > 1)
    {
      ctl->out_size += write_run(out, ctl->out_size, ctl->carry_sym, ctl->carry_len);
      ctl->runs_total++;
      ctl->carry_sym = ends->last_sym; ctl->carry_len = ends->last_len;
      ctl->count = m - 2;
    }
  }
  ctl->slab_base = ctl->out_size;
}

// Bytes of a long run at output offset `state` (mod 64), given its natural size (head + full extension):
// the natural encoding is used whenever it fits into the rest of the block (support.h:267-280).
__device__ __forceinline__ uint32_t long_run_bytes_fast(uint32_t length, uint32_t natural, uint32_t state)
{
  // Lengths 42..169 (head + one extension byte) are nearly all of the long runs. With one byte left in the
  // block the head carries 41 symbols and the rest starts a new block: one more byte, or head + extension.
  if(natural == 2) { return 2u + ((state == 63u && length >= 2u * MAX_RUN - 1u) ? 1u : 0u); }
  return (RLE_BLOCK - state >= natural ? natural : long_run_bytes(length, state));
}

__device__ __forceinline__ uint32_t natural_bytes(uint32_t length)   // length >= MAX_RUN
{
  return 1u + bytecode_length((uint64_t)length - MAX_RUN);
}

// Transducer tile maps: bytes produced by the tile's long runs for each of the 64 residues q of
// (offset of the parallel part + bytes of its earlier long runs) mod 64, with a checkpoint every LONG_SUB
// runs. A long run preceded by `before` short runs starts at offset residue (before + q) mod 64, so the maps
// do not depend on where the slab lands in the output: they are computed before the writer state is known.
__global__ void __launch_bounds__(64)
enc_tile_maps(const uint32_t* __restrict__ long_len, const uint32_t* __restrict__ long_shorts, uint64_t n_long,
              uint32_t* __restrict__ tile_bytes, uint16_t* __restrict__ checkpoints)
{
  // The whole tile is staged first (coalesced, all loads in flight at once): the dependent chain below then
  // only touches shared memory. Per run one byte: bits 0-5 the residue of its `before`, bit 6 "two-byte run"
  // (length 42..169, nearly all long runs), bit 7 "length >= 83" (what the closed form needs to know).
  __shared__ uint32_t s_len[LONG_TILE];
  __shared__ uint8_t  s_entry[LONG_TILE];
  const uint64_t first = (uint64_t)blockIdx.x * LONG_TILE;
  const uint32_t count = (uint32_t)(first + LONG_TILE < n_long ? LONG_TILE : n_long - first);
  for(uint32_t k = threadIdx.x; k < count; k += 64)
  {
    uint32_t length = long_len[first + k];
    s_len[k] = length;
    s_entry[k] = (uint8_t)((long_shorts[first + k] & 63u) | (natural_bytes(length) == 2u ? 64u : 0u) | (length >= 2u * MAX_RUN - 1u ? 128u : 0u));
  }
  __syncthreads();
  uint32_t p = threadIdx.x;
  for(uint32_t chunk = 0; chunk < count; chunk += LONG_SUB)
  {
    checkpoints[((first + chunk) / LONG_SUB) * 64 + threadIdx.x] = (uint16_t)(p - threadIdx.x);
    const uint32_t end = (chunk + LONG_SUB < count ? chunk + LONG_SUB : count);
    for(uint32_t k = chunk; k < end; k++)
    {
      const uint32_t entry = s_entry[k];
      const uint32_t state = (entry + p) & 63u;
      if(entry & 64u) { p += 2u + ((state == 63u) ? (entry >> 7) : 0u); }     // same for all threads of the block
      else { uint32_t length = s_len[k]; p += long_run_bytes_fast(length, natural_bytes(length), state); }
    }
  }
  tile_bytes[(uint64_t)blockIdx.x * 64 + threadIdx.x] = p - threadIdx.x;
}

// Composition of the tile maps in order, starting from the residue of the slab's output offset; the maps
// are staged through shared memory so that every dependent step is a shared-memory lookup.
__global__ void __launch_bounds__(256)
enc_tile_scan(EncodeControl* ctl, const uint32_t* __restrict__ tile_bytes, uint64_t tiles,
              unsigned long long n_short, unsigned long long n_long, unsigned long long* __restrict__ tile_entry)
{
  __shared__ __align__(16) uint32_t staged[SCAN_CHUNK * 64];
  __shared__ unsigned long long entries[SCAN_CHUNK];
  __shared__ unsigned long long carried;
  const uint32_t base = (uint32_t)(ctl->slab_base & 63u);
  if(threadIdx.x == 0) { carried = 0; }
  __syncthreads();
  for(uint64_t first = 0; first < tiles; first += SCAN_CHUNK)
  {
    uint64_t count = (tiles - first < (uint64_t)SCAN_CHUNK ? tiles - first : (uint64_t)SCAN_CHUNK);
    {
      // 16-byte loads, all of a thread's loads issued before the first store
      const uint4* source = reinterpret_cast<const uint4*>(tile_bytes + first * 64);
      uint4* target = reinterpret_cast<uint4*>(staged);
      constexpr int ROUNDS = SCAN_CHUNK * 16 / 256;
      uint4 held[ROUNDS];
#pragma unroll
      for(int round = 0; round < ROUNDS; round++)
      {
        uint64_t k = (uint64_t)round * 256 + threadIdx.x;
        held[round] = (k < count * 16 ? source[k] : make_uint4(0, 0, 0, 0));
      }
#pragma unroll
      for(int round = 0; round < ROUNDS; round++) { target[round * 256 + threadIdx.x] = held[round]; }
    }
    __syncthreads();
    if(threadIdx.x == 0)
    {
      unsigned long long p = carried;
      for(uint64_t t = 0; t < count; t++) { entries[t] = p; p += staged[t * 64 + ((base + (uint32_t)p) & 63u)]; }
      carried = p;
    }
    __syncthreads();
    for(uint64_t t = threadIdx.x; t < count; t += blockDim.x) { tile_entry[first + t] = entries[t]; }
    __syncthreads();
  }
  if(threadIdx.x == 0)
  {
    ctl->n_short = n_short; ctl->n_long = n_long;
    ctl->long_bytes = carried;
    ctl->out_size = ctl->slab_base + n_short + carried;
    ctl->runs_total += ctl->count;
  }
}

// One thread per sub-tile of LONG_SUB long runs: starts from the tile's true entry and the checkpoint of
// that entry residue.
__global__ void enc_long_offsets(const EncodeControl* ctl, const uint32_t* __restrict__ long_len, const uint32_t* __restrict__ long_shorts,
                                 uint64_t n_long,
                                 const unsigned long long* __restrict__ tile_entry, const uint16_t* __restrict__ checkpoints,
                                 uint32_t* __restrict__ long_offset)
{
  uint64_t sub = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t first = sub * LONG_SUB;
  if(first >= n_long) { return; }
  uint64_t last = (first + LONG_SUB < n_long ? first + LONG_SUB : n_long);
  uint32_t base_state = (uint32_t)(ctl->slab_base & 63u);
  unsigned long long entry = tile_entry[first / LONG_TILE];
  unsigned long long p = entry + checkpoints[sub * 64 + ((base_state + (uint32_t)entry) & 63u)];
  for(uint64_t k = first; k < last; k++)
  {
    uint32_t length = long_len[k];
    long_offset[k] = (uint32_t)p;
    uint32_t state = (base_state + long_shorts[k] + (uint32_t)p) & 63u;
    p += long_run_bytes_fast(length, natural_bytes(length), state);
  }
}

//------------------------------------------------------------------------------
// Host orchestration

// `needed` and `valid_bytes` are global offsets (see OutputBuffer::origin).
int ensure_capacity(OutputBuffer* out, uint64_t needed, uint64_t valid_bytes, cudaStream_t stream)
{
  needed = (needed > out->origin ? needed - out->origin : 0);
  valid_bytes = (valid_bytes > out->origin ? valid_bytes - out->origin : 0);
  if(needed <= out->capacity && out->ptr != nullptr) { return BWTM_OK; }
  uint64_t capacity = std::max(needed + (needed >> 2), (uint64_t)(1 << 20));
  DeviceBuffer bigger; BWTM_TRY(bigger.allocate(capacity));
  if(out->ptr != nullptr && valid_bytes > 0)
  {
    BWTM_CUDA(cudaMemcpyAsync(bigger.ptr, out->ptr, valid_bytes, cudaMemcpyDeviceToDevice, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
  }
  if(out->sink != nullptr) { BWTM_CUDA(cudaStreamSynchronize(out->sink->stream)); }   // a copy may still read the old buffer
  device_free(out->ptr);
  out->ptr = static_cast<uint8_t*>(bigger.detach());
  out->capacity = capacity;
  return BWTM_OK;
}

// Device-driven copy into page-locked host memory (accessible through unified addressing). Unlike a DMA copy
// it does not occupy the copy engine, which the encoder's small read-backs of the following slab need.
__global__ void __launch_bounds__(256)
copy_to_host(uint8_t* __restrict__ host, const uint8_t* __restrict__ device, uint64_t bytes)
{
  uint64_t head = (16 - (reinterpret_cast<uintptr_t>(device) & 15)) & 15;
  if(head > bytes) { head = bytes; }
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, threads = (uint64_t)gridDim.x * blockDim.x;
  if(((reinterpret_cast<uintptr_t>(host) + head) & 15) != 0)   // differently aligned: byte copy
  {
    for(uint64_t i = tid; i < bytes; i += threads) { host[i] = device[i]; }
    return;
  }
  for(uint64_t i = tid; i < head; i += threads) { host[i] = device[i]; }
  uint64_t vectors = (bytes - head) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(device + head);
  uint4* dst = reinterpret_cast<uint4*>(host + head);
  // Few CTAs, four independent 16-byte copies per thread and round: enough bytes in flight for the PCIe link
  // without taking many SMs away from the encoder kernels running beside this one.
  uint64_t i = tid;
  for(; i + 3 * threads < vectors; i += 4 * threads)
  {
    uint4 v0 = src[i], v1 = src[i + threads], v2 = src[i + 2 * threads], v3 = src[i + 3 * threads];
    dst[i] = v0; dst[i + threads] = v1; dst[i + 2 * threads] = v2; dst[i + 3 * threads] = v3;
  }
  for(; i < vectors; i += threads) { dst[i] = src[i]; }
  for(uint64_t i = head + vectors * 16 + tid; i < bytes; i += threads) { host[i] = device[i]; }
}

static unsigned copy_ctas()
{
  const char* text = getenv("BWTM_COPY_CTAS");
  int n = (text == nullptr ? 8 : atoi(text));
  return (unsigned)(n < 1 ? 1 : n);
}

int flush_to_host(OutputBuffer* out, const EncodeControl* d_control, cudaStream_t stream)
{
  HostSink* sink = out->sink;
  if(sink == nullptr) { return BWTM_OK; }
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t final_bytes = ctl.out_size - out->origin;
  if(final_bytes <= sink->copied) { return BWTM_OK; }
  if(final_bytes > sink->capacity)
  {
    set_error("host output buffer too small: %llu bytes needed so far, capacity %llu",
              (unsigned long long)final_bytes, (unsigned long long)sink->capacity);
    return BWTM_ERR_CAPACITY;
  }
  BWTM_CUDA(cudaEventRecord(sink->ready, stream));
  BWTM_CUDA(cudaStreamWaitEvent(sink->stream, sink->ready, 0));
  if(sink->device_visible)
  {
    copy_to_host<<<copy_ctas(), 256, 0, sink->stream>>>(sink->ptr + sink->copied, out->ptr + sink->copied, final_bytes - sink->copied);
    BWTM_LAUNCH_CHECK();
  }
  else
  {
    BWTM_CUDA(cudaMemcpyAsync(sink->ptr + sink->copied, out->ptr + sink->copied, final_bytes - sink->copied,
                              cudaMemcpyDeviceToHost, sink->stream));
  }
  sink->copied = final_bytes;
  return BWTM_OK;
}

// K3 and the run-parallel half of K5, on plane chunks.
//
// A maximal run (what the reference's RunBuffer produces, utils.h:121-142) starts wherever a symbol differs
// from its predecessor: with the three bit planes of 32 positions in one uint4 that is three XORs. Runs are
// never stored. The slab is cut into tiles of TILE positions (one chunk per thread); a run belongs to the
// tile it starts in and ends at the next run start, which is in the same chunk, in a later chunk of the
// tile (block suffix-minimum) or at the first run start of a later tile (suffix-minimum over the tiles).
// Three passes read the chunks:
//   run_tile_survey : run starts, first and last start, (short, long) run counts per tile;
//   enc_collect_long: length and number of preceding short runs of every long run, in order;
//   enc_write_tiles : the bytes, once the transducer has placed the long runs.
// (short, long) counts travel packed in one 64-bit word: short low, long high. The first and the last run
// of the slab interact with the pending run of the writer and are left to enc_head.
constexpr int      RUN_THREADS = TILE / 32;     // 128
constexpr int      RUN_WARPS   = RUN_THREADS / 32;
constexpr uint32_t NO_START    = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long run_class(uint32_t length)
{
  return (length < (uint32_t)MAX_RUN ? 1ull : (1ull << 32));
}

__device__ __forceinline__ uint32_t chunk_symbol(const uint4& x, uint32_t i)
{
  return ((x.x >> i) & 1u) | (((x.y >> i) & 1u) << 1) | (((x.z >> i) & 1u) << 2);
}

// Run-start flags of the thread's chunk (bit i: slab position 32 * chunk + i starts a maximal run). Must be
// called by all threads of the block.
__device__ __forceinline__ uint32_t chunk_start_flags(const uint4* __restrict__ planes, uint64_t chunk, uint64_t n, uint4& x)
{
  const bool active = (chunk * 32 < n);
  x = make_uint4(0, 0, 0, 0);
  if(active) { x = planes[chunk]; }
  uint32_t tops = (x.x >> 31) | ((x.y >> 31) << 1) | ((x.z >> 31) << 2);
  uint32_t previous = __shfl_up_sync(0xFFFFFFFFu, tops, 1);
  if((threadIdx.x & 31) == 0 && active && chunk > 0)
  {
    uint4 q = planes[chunk - 1];
    previous = (q.x >> 31) | ((q.y >> 31) << 1) | ((q.z >> 31) << 2);
  }
  uint32_t flags = (x.x ^ ((x.x << 1) | (previous & 1u))) | (x.y ^ ((x.y << 1) | ((previous >> 1) & 1u))) |
                   (x.z ^ ((x.z << 1) | ((previous >> 2) & 1u)));
  if(chunk == 0) { flags |= 1u; }
  if(!active) { return 0; }
  uint64_t left = n - chunk * 32;
  if(left < 32) { flags &= low_mask((int)left); }
  return flags;
}

// First run start after the thread's chunk: the smallest `first` of the later threads of the block, or
// `after_tile`. `tile_first` gets the smallest `first` of the whole block.
__device__ __forceinline__ uint32_t tile_next_start(uint32_t first, uint32_t* warp_first, uint32_t after_tile, uint32_t& tile_first)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inclusive = first;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    uint32_t other = __shfl_down_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane + offset < 32 && other < inclusive) { inclusive = other; }
  }
  uint32_t next = __shfl_down_sync(0xFFFFFFFFu, inclusive, 1);
  if(lane == 31) { next = NO_START; }
  if(lane == 0) { warp_first[warp] = inclusive; }
  __syncthreads();
  tile_first = NO_START;
#pragma unroll
  for(int w = RUN_WARPS - 1; w >= 0; w--)
  {
    if(w > warp && warp_first[w] < next) { next = warp_first[w]; }
    if(warp_first[w] < tile_first) { tile_first = warp_first[w]; }
  }
  return (next < after_tile ? next : after_tile);
}

// Exclusive prefix of `value` over the threads of the block; `total` is the block's sum.
__device__ __forceinline__ unsigned long long tile_prefix(unsigned long long value, unsigned long long* warp_sums, unsigned long long& total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inclusive = value;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    unsigned long long other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= offset) { inclusive += other; }
  }
  if(lane == 31) { warp_sums[warp] = inclusive; }
  __syncthreads();
  unsigned long long before = 0; total = 0;
#pragma unroll
  for(int w = 0; w < RUN_WARPS; w++)
  {
    if(w < warp) { before += warp_sums[w]; }
    total += warp_sums[w];
  }
  return before + inclusive - value;
}

// Runs of one chunk. A run of MAX_RUN or more symbols that starts in a 32-position chunk extends beyond it,
// so only the last start of a chunk can be a long run: every other start is a short run whose length is the
// distance to the next start. Nothing loops over runs to classify them.
struct ChunkRuns
{
  uint4 x; uint32_t base, flags, live, top, next; bool long_here;
  // live: starts that belong to the parallel part; top: bit of the chunk's last start; next: first start after the chunk
  __device__ __forceinline__ unsigned long long classes() const
  {
    unsigned long long is_long = (long_here ? 1ull : 0ull);
    return (unsigned long long)__popc(live) - is_long + (is_long << 32);
  }
};

// `next` = NO_START means that the end of the chunk's last run is not known yet: it is left out.
__device__ __forceinline__ void classify_chunk(ChunkRuns& runs, uint32_t last_start)
{
  runs.live = runs.flags;
  runs.top = (runs.flags != 0 ? 31u - (uint32_t)__clz(runs.flags) : 0u);
  if(runs.base == 0) { runs.live &= ~1u; }                                            // first run of the slab
  if(last_start - runs.base < 32u) { runs.live &= ~(1u << (last_start - runs.base)); } // last run of the slab
  if(runs.next == NO_START) { runs.live &= ~(1u << runs.top); }
  runs.long_here = (((runs.live >> runs.top) & 1u) != 0 && runs.next - (runs.base + runs.top) >= (uint32_t)MAX_RUN);
}

__global__ void __launch_bounds__(RUN_THREADS)
run_tile_survey(const uint4* __restrict__ planes, uint64_t n, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_first,
                uint32_t* __restrict__ tile_last, unsigned long long* __restrict__ tile_class, SlabEnds* __restrict__ ends)
{
  __shared__ uint32_t warp_first[RUN_WARPS];
  __shared__ unsigned long long warp_sums[RUN_WARPS];
  __shared__ uint32_t warp_last[RUN_WARPS];
  const uint64_t chunk = (uint64_t)blockIdx.x * RUN_THREADS + threadIdx.x;
  const uint32_t base = (uint32_t)(chunk * 32);
  uint4 x;
  uint32_t flags = chunk_start_flags(planes, chunk, n, x);
  uint32_t first = (flags != 0 ? base + (uint32_t)(__ffs(flags) - 1) : NO_START);
  uint32_t last = (flags != 0 ? base + (uint32_t)(31 - __clz(flags)) : 0u);
  uint32_t block_first;
  uint32_t next = tile_next_start(first, warp_first, NO_START, block_first);

  // Runs whose end lies inside the tile; the tile's last run is classified by run_tile_resolve. The slab's
  // last run is the last run of its tile, so nothing has to be known about it here. The number of run
  // starts rides along in bits 48.. of the packed word (at most TILE per tile).
  ChunkRuns runs;
  runs.x = x; runs.base = base; runs.flags = flags; runs.next = next;
  classify_chunk(runs, NO_START);
  unsigned long long packed = ((unsigned long long)__popc(flags) << 48) + runs.classes();
#pragma unroll
  for(int offset = 16; offset > 0; offset >>= 1)
  {
    packed += __shfl_down_sync(0xFFFFFFFFu, packed, offset);
    uint32_t other = __shfl_down_sync(0xFFFFFFFFu, last, offset);
    if(other > last) { last = other; }
  }
  if((threadIdx.x & 31) == 0) { warp_sums[threadIdx.x >> 5] = packed; warp_last[threadIdx.x >> 5] = last; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    packed = 0; last = 0;
    for(int w = 0; w < RUN_WARPS; w++) { packed += warp_sums[w]; if(warp_last[w] > last) { last = warp_last[w]; } }
    uint32_t count = (uint32_t)(packed >> 48);
    tile_count[blockIdx.x] = count;
    tile_first[blockIdx.x] = block_first;
    tile_last[blockIdx.x] = last;
    // short counts fit 13 bits and long counts 7 bits per tile: unpack into the (low, high) words
    tile_class[blockIdx.x] = packed & 0x0000FFFFFFFFFFFFull;
    if(count > 0) { atomicMax(&(ends->last_start), last); }
  }
}

// next_first[t] = first run start in the tiles after t (n when there is none). Starts increase with the
// tile, so that is the first entry of the next tile that has one: nearly always tile t + 1. Tiles are grouped
// by GROUP_TILES to bound the search when very long runs leave many tiles without a start.
constexpr int GROUP_TILES = 1024;

__global__ void __launch_bounds__(256)
run_group_first(const uint32_t* __restrict__ tile_first, uint64_t tiles, uint32_t* __restrict__ group_first)
{
  __shared__ uint32_t warp_first[8];
  uint32_t smallest = NO_START;
  for(int k = 0; k < GROUP_TILES / 256; k++)
  {
    uint64_t t = (uint64_t)blockIdx.x * GROUP_TILES + k * 256 + threadIdx.x;
    if(t < tiles) { uint32_t v = tile_first[t]; if(v < smallest) { smallest = v; } }
  }
#pragma unroll
  for(int offset = 16; offset > 0; offset >>= 1)
  {
    uint32_t other = __shfl_down_sync(0xFFFFFFFFu, smallest, offset);
    if(other < smallest) { smallest = other; }
  }
  if((threadIdx.x & 31) == 0) { warp_first[threadIdx.x >> 5] = smallest; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    for(int w = 1; w < 8; w++) { if(warp_first[w] < smallest) { smallest = warp_first[w]; } }
    group_first[blockIdx.x] = smallest;
  }
}

__global__ void run_tile_next(const uint32_t* __restrict__ tile_first, const uint32_t* __restrict__ group_first, uint64_t tiles,
                              uint32_t n, uint32_t* __restrict__ next_first)
{
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= tiles) { return; }
  uint64_t group = t / GROUP_TILES, groups = div_up(tiles, (uint64_t)GROUP_TILES);
  uint64_t group_end = ((group + 1) * GROUP_TILES < tiles ? (group + 1) * GROUP_TILES : tiles);
  uint32_t found = NO_START;
  for(uint64_t u = t + 1; u < group_end && found == NO_START; u++) { found = tile_first[u]; }
  for(uint64_t g = group + 1; g < groups && found == NO_START; g++) { found = group_first[g]; }
  next_first[t] = (found == NO_START ? n : found);
}

// Classifies the last run of every tile (its end is now known), and describes the first and the last run of
// the slab for enc_head. tile_count must already be scanned (tile_count[tiles] = number of runs).
__global__ void run_tile_resolve(const uint4* __restrict__ planes, uint32_t n, uint64_t tiles, const uint32_t* __restrict__ run_base,
                                 const uint32_t* __restrict__ tile_last, const uint32_t* __restrict__ next_first,
                                 unsigned long long* __restrict__ tile_class, SlabEnds* __restrict__ ends)
{
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(t < tiles && run_base[t + 1] > run_base[t])
  {
    uint32_t position = tile_last[t];
    if(position != 0 && position != ends->last_start) { tile_class[t] += run_class(next_first[t] - position); }
  }
  if(t != 0) { return; }
  ends->runs = run_base[tiles];
  uint4 x = planes[0];
  ends->first_sym = chunk_symbol(x, 0);
  uint32_t first_len = next_first[0];
  if(run_base[1] > 1)   // the second run starts in tile 0: look for the first symbol that differs
  {
    for(uint32_t position = 1; position < (uint32_t)TILE; position++)
    {
      if((position & 31u) == 0) { x = planes[position >> 5]; }
      if(chunk_symbol(x, position & 31u) != ends->first_sym) { first_len = position; break; }
    }
  }
  ends->first_len = first_len;
  uint32_t start = ends->last_start;
  x = planes[start >> 5];
  ends->last_sym = chunk_symbol(x, start & 31u);
  ends->last_len = n - start;
}

// Common part of the two passes that need every run with its (short, long) prefix.
__device__ __forceinline__ unsigned long long tile_view(const uint4* __restrict__ planes, uint64_t n, const uint32_t* __restrict__ next_first,
                                                        const unsigned long long* __restrict__ class_base, const SlabEnds* __restrict__ ends,
                                                        uint32_t* warp_first, unsigned long long* warp_sums, ChunkRuns& runs)
{
  const uint64_t chunk = (uint64_t)blockIdx.x * RUN_THREADS + threadIdx.x;
  runs.base = (uint32_t)(chunk * 32);
  runs.flags = chunk_start_flags(planes, chunk, n, runs.x);
  uint32_t first = (runs.flags != 0 ? runs.base + (uint32_t)(__ffs(runs.flags) - 1) : NO_START);
  uint32_t block_first;
  runs.next = tile_next_start(first, warp_first, next_first[blockIdx.x], block_first);
  classify_chunk(runs, ends->last_start);
  unsigned long long total;
  return class_base[blockIdx.x] + tile_prefix(runs.classes(), warp_sums, total);
}

// Long runs in order: length and number of short runs before each (all the transducer needs).
__global__ void __launch_bounds__(RUN_THREADS)
enc_collect_long(const uint4* __restrict__ planes, uint64_t n, const uint32_t* __restrict__ next_first,
                 const unsigned long long* __restrict__ class_base, const SlabEnds* __restrict__ ends,
                 uint32_t* __restrict__ long_len, uint32_t* __restrict__ long_shorts)
{
  __shared__ uint32_t warp_first[RUN_WARPS];
  __shared__ unsigned long long warp_sums[RUN_WARPS];
  if((class_base[blockIdx.x + 1] >> 32) == (class_base[blockIdx.x] >> 32)) { return; }   // no long run starts in this tile
  ChunkRuns runs;
  unsigned long long before = tile_view(planes, n, next_first, class_base, ends, warp_first, warp_sums, runs);
  if(runs.long_here)
  {
    long_len[before >> 32] = runs.next - (runs.base + runs.top);
    long_shorts[before >> 32] = (uint32_t)before + (uint32_t)__popc(runs.live & low_mask((int)runs.top));
  }
}

// The bytes of a tile are one contiguous piece of the output. They are assembled in shared memory, laid out
// with the alignment of their destination, and stored 16 bytes at a time.
constexpr int STAGE_BYTES = 4864;   // a tile of 4096 one-symbol runs, the alignment shift and some long runs

__global__ void __launch_bounds__(RUN_THREADS)
enc_write_tiles(const EncodeControl* __restrict__ ctl, const uint4* __restrict__ planes, uint64_t n, const uint32_t* __restrict__ next_first,
                const unsigned long long* __restrict__ class_base, const SlabEnds* __restrict__ ends,
                const uint32_t* __restrict__ long_offset, uint8_t* __restrict__ out)
{
  __shared__ uint32_t warp_first[RUN_WARPS];
  __shared__ unsigned long long warp_sums[RUN_WARPS];
  __shared__ __align__(16) uint8_t stage[STAGE_BYTES];
  ChunkRuns runs;
  unsigned long long before = tile_view(planes, n, next_first, class_base, ends, warp_first, warp_sums, runs);
  const unsigned long long n_long = ctl->n_long, long_bytes = ctl->long_bytes, slab_base = ctl->slab_base;

  // Output range of the tile, from the prefixes of this tile and of the next one.
  unsigned long long tile_lo = class_base[blockIdx.x], tile_hi = class_base[blockIdx.x + 1];
  uint64_t tile_begin = slab_base + (tile_lo & 0xFFFFFFFFull) + ((tile_lo >> 32) < n_long ? (uint64_t)long_offset[tile_lo >> 32] : long_bytes);
  uint64_t tile_end = slab_base + (tile_hi & 0xFFFFFFFFull) + ((tile_hi >> 32) < n_long ? (uint64_t)long_offset[tile_hi >> 32] : long_bytes);
  if(tile_end <= tile_begin) { return; }
  const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(out + tile_begin) & 15u);
  const bool staged = (tile_end - tile_begin + shift <= (uint64_t)STAGE_BYTES);
  uint8_t* target = (staged ? stage : out);
  const uint64_t origin = (staged ? tile_begin - shift : 0);   // target[x - origin] holds output byte x (wraps harmlessly)

  uint64_t offset = slab_base + (before & 0xFFFFFFFFull) + ((before >> 32) < n_long ? (uint64_t)long_offset[before >> 32] : long_bytes);
  uint32_t remaining = runs.live;
  while(remaining != 0)
  {
    uint32_t i = (uint32_t)(__ffs(remaining) - 1);
    remaining &= remaining - 1;
    uint32_t later = runs.flags & ~low_mask((int)i + 1);
    uint32_t length = (later != 0 ? (uint32_t)(__ffs(later) - 1) - i : runs.next - (runs.base + i));
    uint32_t comp = chunk_symbol(runs.x, i);
    if(length < (uint32_t)MAX_RUN) { target[offset - origin] = (uint8_t)(comp + SIGMA * (length - 1)); offset++; }
    else { write_run(target, offset, comp, length, origin); }   // the chunk's last run: nothing follows it here
  }
  if(!staged) { return; }
  __syncthreads();

  uint8_t* destination = out + tile_begin;
  const uint32_t bytes = (uint32_t)(tile_end - tile_begin);
  uint32_t head = (shift == 0 ? 0u : 16u - shift);
  if(head > bytes) { head = bytes; }
  const uint32_t vectors = (bytes - head) / 16;
  if(threadIdx.x < head) { destination[threadIdx.x] = stage[shift + threadIdx.x]; }
  const uint4* staged_vectors = reinterpret_cast<const uint4*>(stage + shift + head);
  uint4* destination_vectors = reinterpret_cast<uint4*>(destination + head);
  for(uint32_t v = threadIdx.x; v < vectors; v += RUN_THREADS) { destination_vectors[v] = staged_vectors[v]; }
  const uint32_t done = head + vectors * 16;
  if(threadIdx.x < bytes - done) { destination[done + threadIdx.x] = stage[shift + done + threadIdx.x]; }
}

int SlabEncoder::init(uint64_t max_symbols_, cudaStream_t stream)
{
  max_symbols = max_symbols_;
  long_capacity = 0;
  slab_planes = nullptr; slab_symbols = 0;
  BWTM_TRY(placed.allocate(sizeof(EncodeControl)));
  BWTM_TRY(ends.allocate(sizeof(SlabEnds)));
  uint64_t tiles = div_up(max_symbols, TILE) + 1;
  BWTM_TRY(tile_count.allocate(tiles * sizeof(uint32_t)));
  BWTM_TRY(tile_first.allocate(tiles * sizeof(uint32_t)));
  BWTM_TRY(tile_last.allocate(tiles * sizeof(uint32_t)));
  BWTM_TRY(next_first.allocate(tiles * sizeof(uint32_t)));
  BWTM_TRY(group_first.allocate((tiles / GROUP_TILES + 1) * sizeof(uint32_t)));
  BWTM_TRY(class_base.allocate(tiles * sizeof(unsigned long long)));
  size_t count_temp = 0, class_temp = 0;
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, count_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)tiles, stream));
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, class_temp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t)tiles, stream));
  BWTM_TRY(cub_temp.allocate(std::max(count_temp, class_temp)));
  return BWTM_OK;
}

// Work arrays of the transducer, sized by the number of long runs of the slab; grow-only.
int SlabEncoder::reserve_long(uint64_t long_runs)
{
  if(long_runs <= long_capacity) { return BWTM_OK; }
  uint64_t capacity = long_runs + (long_runs >> 3) + 1024;
  uint64_t long_tiles = div_up(capacity, LONG_TILE);
  BWTM_TRY(long_len.allocate(capacity * sizeof(uint32_t)));
  BWTM_TRY(long_shorts.allocate(capacity * sizeof(uint32_t)));
  BWTM_TRY(long_offset.allocate(capacity * sizeof(uint32_t)));
  BWTM_TRY(tile_bytes.allocate(long_tiles * 64 * sizeof(uint32_t)));
  BWTM_TRY(tile_entry.allocate(long_tiles * sizeof(unsigned long long)));
  BWTM_TRY(checkpoints.allocate((capacity / LONG_SUB + 1) * 64 * sizeof(uint16_t)));
  long_capacity = capacity;
  return BWTM_OK;
}

// K3 and the state-free half of K5 for `symbols` consecutive symbols given as plane chunks (chunk 0 holds
// the first symbol in bit 0; bits beyond the last symbol are zero): the maximal runs and, for the runs
// [1, m - 1), the short/long prefixes, the long runs in order and their transducer tile maps. The chunks
// must stay valid until emit() has run.
int SlabEncoder::detect(const uint4* d_planes, uint64_t symbols, cudaStream_t stream)
{
  detected_runs = 0; part_count = 0; part_short = 0; part_long = 0;
  slab_planes = d_planes; slab_symbols = symbols;
  if(symbols == 0) { return BWTM_OK; }
  if(symbols > max_symbols) { set_error("slab of %llu symbols exceeds the encoder capacity", (unsigned long long)symbols); return BWTM_ERR_INTERNAL; }
  size_t temp_bytes = cub_temp.bytes;
  const uint64_t tiles = div_up(symbols, TILE);
  uint32_t* counts = tile_count.as<uint32_t>();
  unsigned long long* classes = class_base.as<unsigned long long>();
  BWTM_CUDA(cudaMemsetAsync(ends.ptr, 0, sizeof(SlabEnds), stream));
  run_tile_survey<<<(unsigned)tiles, RUN_THREADS, 0, stream>>>(d_planes, symbols, counts, tile_first.as<uint32_t>(), tile_last.as<uint32_t>(),
                                                               classes, ends.as<SlabEnds>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemsetAsync(counts + tiles, 0, sizeof(uint32_t), stream));
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp.ptr, temp_bytes, counts, counts, (int64_t)(tiles + 1), stream));
  count_launch(2);
  run_group_first<<<(unsigned)div_up(tiles, GROUP_TILES), 256, 0, stream>>>(tile_first.as<uint32_t>(), tiles, group_first.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  run_tile_next<<<(unsigned)div_up(tiles, 256), 256, 0, stream>>>(tile_first.as<uint32_t>(), group_first.as<uint32_t>(), tiles,
                                                                 (uint32_t)symbols, next_first.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  run_tile_resolve<<<(unsigned)div_up(tiles, 256), 256, 0, stream>>>(d_planes, (uint32_t)symbols, tiles, counts, tile_last.as<uint32_t>(),
                                                                    next_first.as<uint32_t>(), classes, ends.as<SlabEnds>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemsetAsync(classes + tiles, 0, sizeof(unsigned long long), stream));
  temp_bytes = cub_temp.bytes;
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp.ptr, temp_bytes, classes, classes, (int64_t)(tiles + 1), stream));
  count_launch(2);

  uint32_t total_runs = 0; unsigned long long totals = 0;
  BWTM_CUDA(cudaMemcpyAsync(&total_runs, counts + tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaMemcpyAsync(&totals, classes + tiles, sizeof(totals), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t m = total_runs;
  if(m == 0) { set_error("run detection found no runs in %llu symbols", (unsigned long long)symbols); return BWTM_ERR_INTERNAL; }
  detected_runs = m;
  if(m < 3) { return BWTM_OK; }
  part_count = m - 2; part_short = totals & 0xFFFFFFFFull; part_long = totals >> 32;
  if(part_short + part_long != part_count)
  {
    set_error("run classes (%llu short, %llu long) do not add up to %llu runs", (unsigned long long)part_short,
              (unsigned long long)part_long, (unsigned long long)part_count);
    return BWTM_ERR_INTERNAL;
  }
  if(part_long > 0)
  {
    BWTM_TRY(this->reserve_long(part_long));
    enc_collect_long<<<(unsigned)tiles, RUN_THREADS, 0, stream>>>(d_planes, symbols, next_first.as<uint32_t>(), classes, ends.as<SlabEnds>(),
                                                                 long_len.as<uint32_t>(), long_shorts.as<uint32_t>());
    BWTM_LAUNCH_CHECK();
    enc_tile_maps<<<(unsigned)div_up(part_long, LONG_TILE), 64, 0, stream>>>(long_len.as<uint32_t>(), long_shorts.as<uint32_t>(), part_long,
                                                                             tile_bytes.as<uint32_t>(), checkpoints.as<uint16_t>());
    BWTM_LAUNCH_CHECK();
  }
  return BWTM_OK;
}

// The state-dependent half of K5, first step: consumes the writer state in d_control (pending run, output
// offset) and leaves the state after this slab there. Cheap and sequential; the next slab or GPU slice
// can start from d_control as soon as this returns (stream order).
int SlabEncoder::advance(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  uint64_t m = detected_runs;
  if(m == 0) { return BWTM_OK; }
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  // Upper bound of what this slab can add: two sequential runs, one byte per short run, 16 per long run.
  BWTM_TRY(ensure_capacity(out, ctl.out_size + 512 + part_short + 16 * part_long, ctl.out_size, stream));
  enc_head<<<1, 1, 0, stream>>>(d_control, ends.as<SlabEnds>(), out->at_origin(), 0);
  BWTM_LAUNCH_CHECK();
  if(part_count == 0) { return BWTM_OK; }
  uint64_t long_tiles = div_up(part_long, LONG_TILE);
  enc_tile_scan<<<1, 256, 0, stream>>>(d_control, tile_bytes.as<uint32_t>(), long_tiles, part_short, part_long, tile_entry.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpyAsync(placed.ptr, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToDevice, stream));
  return BWTM_OK;
}

// Second step: writes the bytes of the parallel part at the offsets fixed by advance(). The plane chunks
// given to detect() are read again here.
int SlabEncoder::emit(OutputBuffer* out, cudaStream_t stream)
{
  if(detected_runs == 0 || part_count == 0) { return BWTM_OK; }
  const EncodeControl* where = placed.as<EncodeControl>();
  if(part_long > 0)
  {
    enc_long_offsets<<<(unsigned)div_up(div_up(part_long, LONG_SUB), 128), 128, 0, stream>>>(
      where, long_len.as<uint32_t>(), long_shorts.as<uint32_t>(), part_long,
      tile_entry.as<unsigned long long>(), checkpoints.as<uint16_t>(), long_offset.as<uint32_t>());
    BWTM_LAUNCH_CHECK();
  }
  enc_write_tiles<<<(unsigned)div_up(slab_symbols, TILE), RUN_THREADS, 0, stream>>>(
    where, slab_planes, slab_symbols, next_first.as<uint32_t>(), class_base.as<unsigned long long>(), ends.as<SlabEnds>(),
    long_offset.as<uint32_t>(), out->at_origin());
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

int SlabEncoder::write(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  BWTM_TRY(this->advance(out, d_control, stream));
  return this->emit(out, stream);
}

int SlabEncoder::encode(const uint4* d_planes, uint64_t symbols, OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  BWTM_TRY(this->detect(d_planes, symbols, stream));
  return this->write(out, d_control, stream);
}

// Flushes the pending run (bwt.cpp:279-281).
int SlabEncoder::finish(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  BWTM_TRY(ensure_capacity(out, ctl.out_size + 256, ctl.out_size, stream));
  enc_head<<<1, 1, 0, stream>>>(d_control, nullptr, out->at_origin(), 1);
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

uint64_t clamp_slab(uint64_t slab_symbols, uint64_t total, bool allow_large)
{
  if(slab_symbols == 0) { slab_symbols = 1ull << 30; }
  slab_symbols = std::min(slab_symbols, allow_large ? MAX_SLAB_SYMBOLS : (uint64_t)1 << 30);
  slab_symbols = div_up(slab_symbols, TILE) * TILE;
  return std::min(slab_symbols, div_up(std::max(total, (uint64_t)1), TILE) * TILE);
}

uint64_t interleave_tile_size() { return TILE; }

template<class KeyT>
int interleave_slab(const bwtm_index* a, const bwtm_index* b, const KeyT* d_keys, uint64_t key_base, uint64_t key_count,
                    uint64_t p0, uint64_t p1, uint4* d_merged, uint64_t* d_tile_j, cudaStream_t stream,
                    unsigned long long* d_distinct_keys)
{
  if(p1 <= p0) { return BWTM_OK; }
  uint64_t tiles = div_up(p1 - p0, TILE);
  k4_partition<KeyT><<<(unsigned)div_up(tiles + 1, 256), 256, 0, stream>>>(d_keys, key_base, key_count, p0, p1, tiles, d_tile_j);
  BWTM_LAUNCH_CHECK();
  k4_interleave<KeyT><<<(unsigned)tiles, IL_THREADS, 0, stream>>>(device_view(a), device_view(b), d_keys, key_base, d_tile_j, p0, p1, d_merged, d_distinct_keys);
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

template int interleave_slab<uint32_t>(const bwtm_index*, const bwtm_index*, const uint32_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint4*, uint64_t*, cudaStream_t, unsigned long long*);
template int interleave_slab<uint64_t>(const bwtm_index*, const bwtm_index*, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint4*, uint64_t*, cudaStream_t, unsigned long long*);

template<class KeyT>
int interleave_range(const bwtm_index* a, const bwtm_index* b, const KeyT* d_keys, uint64_t key_base, uint64_t key_count,
                     uint64_t begin, uint64_t end, uint64_t slab_symbols,
                     OutputBuffer* out, EncodeControl* d_control, bool finish,
                     float* interleave_ms, float* encode_ms, cudaStream_t stream, unsigned long long* d_distinct_keys,
                     uint4* d_result_records)
{
  slab_symbols = clamp_slab(slab_symbols, end - begin);
  uint64_t max_tiles = slab_symbols / TILE;
  DeviceBuffer merged, tile_j;
  // The merged symbols of a slab are plane chunks: written into the records of the result when the caller
  // provides them (then they already are the rank structure's bit planes), else into a slab-sized scratch.
  if(d_result_records == nullptr) { BWTM_TRY(merged.allocate(slab_symbols / 2)); }
  else if((begin & 31) != 0) { set_error("result records need a chunk-aligned range"); return BWTM_ERR_INTERNAL; }
  BWTM_TRY(tile_j.allocate((max_tiles + 2) * sizeof(uint64_t)));
  SlabEncoder encoder;
  BWTM_TRY(encoder.init(slab_symbols, stream));
  EventTimer timer(stream);
  for(uint64_t p0 = begin; p0 < end; p0 += slab_symbols)
  {
    uint64_t p1 = std::min(p0 + slab_symbols, end);
    uint4* planes = (d_result_records != nullptr ? d_result_records + (p0 >> 5) : merged.as<uint4>());
    timer.start();
    BWTM_TRY(interleave_slab<KeyT>(a, b, d_keys, key_base, key_count, p0, p1, planes, tile_j.as<uint64_t>(), stream, d_distinct_keys));
    *interleave_ms += timer.stop();
    timer.start();
    BWTM_TRY(encoder.encode(planes, p1 - p0, out, d_control, stream));
    *encode_ms += timer.stop();
    BWTM_TRY(flush_to_host(out, d_control, stream));
  }
  if(finish)
  {
    timer.start();
    BWTM_TRY(encoder.finish(out, d_control, stream));
    *encode_ms += timer.stop();
    BWTM_TRY(flush_to_host(out, d_control, stream));
  }
  return BWTM_OK;
}

template int interleave_range<uint32_t>(const bwtm_index*, const bwtm_index*, const uint32_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                                        OutputBuffer*, EncodeControl*, bool, float*, float*, cudaStream_t, unsigned long long*, uint4*);
template int interleave_range<uint64_t>(const bwtm_index*, const bwtm_index*, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t,
                                        OutputBuffer*, EncodeControl*, bool, float*, float*, cudaStream_t, unsigned long long*, uint4*);

//------------------------------------------------------------------------------

int bit_length_host(uint64_t v) { int n = 0; while(v > 0) { n++; v >>= 1; } return (n == 0 ? 1 : n); }

// Wraps freshly encoded RLE bytes into an index (K0 unless skipped). `counts` (6 values) are the
// expected per-comp counts, or NULL.
int finish_index(OutputBuffer* out, uint64_t rle_bytes, const uint64_t* counts, uint64_t sequences, bool skip_index,
                 cudaStream_t stream, bwtm_index** result, DeviceBuffer* filled_records, uint64_t size)
{
  DeviceBuffer exact; BWTM_TRY(exact.allocate(rle_bytes + RLE_PADDING, true));
  BWTM_CUDA(cudaMemcpyAsync(exact.ptr, out->ptr, rle_bytes, cudaMemcpyDeviceToDevice, stream));
  BWTM_CUDA(cudaMemsetAsync(exact.as<uint8_t>() + rle_bytes, 0, RLE_PADDING, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  if(out->sink == nullptr && !out->borrowed) { device_free(out->ptr); out->ptr = nullptr; out->capacity = 0; }   // else a host copy may still read it

  if(!skip_index)
  {
    if(filled_records != nullptr && filled_records->ptr != nullptr)
    {
      BWTM_TRY(index_from_planes(exact.as<uint8_t>(), rle_bytes, filled_records->as<uint4>(), size, stream, result));
      filled_records->detach();
    }
    else { BWTM_TRY(index_from_device_rle(exact.as<uint8_t>(), rle_bytes, stream, result)); }
    exact.detach();
    bwtm_index* m = *result;
    for(int c = 0; counts != nullptr && c < SIGMA; c++)
    {
      if(m->counts[c] != counts[c])
      {
        set_error("encoded count of comp %d is %llu, expected %llu", c, (unsigned long long)m->counts[c],
                  (unsigned long long)counts[c]);
        index_free(m); *result = nullptr;
        return BWTM_ERR_INTERNAL;
      }
    }
    return BWTM_OK;
  }

  bwtm_index* m = new bwtm_index();
  std::memset(m, 0, sizeof(bwtm_index));
  cudaGetDevice(&(m->device));
  m->d_rle = static_cast<uint8_t*>(exact.detach()); m->rle_bytes = rle_bytes;
  m->sequences = sequences;                                                   // bwt.cpp:305-306
  for(int c = 0; c < SIGMA; c++) { m->counts[c] = counts[c]; m->size += counts[c]; }
  m->C[0] = 0;
  for(int c = 0; c < SIGMA; c++) { m->C[c + 1] = m->C[c] + m->counts[c]; }  // fmi.cpp:367-368
  m->device_bytes = rle_bytes + RLE_PADDING;
  *result = m;
  return BWTM_OK;
}

// Records whose plane chunks hold a complete sequence of n symbols -> run-length bytes (K3 + K5, slab by slab)
// -> index. Takes the records over on success.
static int index_from_filled_records(DeviceBuffer& records, uint64_t n, uint64_t slab_symbols, cudaStream_t stream, bwtm_index** out)
{
  uint64_t slab = clamp_slab(slab_symbols, n);
  SlabEncoder encoder; BWTM_TRY(encoder.init(slab, stream));
  DeviceBuffer control; BWTM_TRY(control.allocate(sizeof(EncodeControl)));
  BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream));
  OutputBuffer buffer = { nullptr, 0, 0, nullptr };
  int rc = ensure_capacity(&buffer, n / 4 + (1 << 20), 0, stream);
  for(uint64_t p0 = 0; rc == BWTM_OK && p0 < n; p0 += slab)
  {
    rc = encoder.encode(records.as<uint4>() + (p0 >> 5), std::min(slab, n - p0), &buffer, control.as<EncodeControl>(), stream);
  }
  if(rc == BWTM_OK) { rc = encoder.finish(&buffer, control.as<EncodeControl>(), stream); }
  EncodeControl ctl;
  if(rc == BWTM_OK && cudaMemcpy(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost) != cudaSuccess)
  {
    set_error("cannot read the encoder state"); rc = BWTM_ERR_CUDA;
  }
  if(rc == BWTM_OK) { rc = finish_index(&buffer, ctl.out_size, nullptr, 0, false, stream, out, &records, n); }
  device_free(buffer.ptr);
  return rc;
}

int index_from_symbols(const uint8_t* d_symbols, uint64_t n, uint64_t slab_symbols, cudaStream_t stream, bwtm_index** out)
{
  if(n == 0) { set_error("empty sequence"); return BWTM_ERR_ARGUMENT; }
  // Symbols -> plane chunks of the records (the layout the encoder reads).
  if((reinterpret_cast<uintptr_t>(d_symbols) & 15) != 0) { set_error("symbol array is not 16-byte aligned"); return BWTM_ERR_INTERNAL; }
  DeviceBuffer records;
  uint64_t record_bytes = ((n >> RECORD_SHIFT) + 1) * 64;
  BWTM_TRY(records.allocate(record_bytes, true));
  BWTM_CUDA(cudaMemsetAsync(records.ptr, 0, record_bytes, stream));
  BWTM_TRY(planes_from_symbols(d_symbols, 0, n, records.as<uint4>(), stream));
  return index_from_filled_records(records, n, slab_symbols, stream, out);
}

// One run per byte (RopeBWT: length << 3 | comp; SGA: comp << 5 | length; lengths 1..31): the device counterpart
// of RopeData::read (formats.cpp:286-310) and SGAData::read (formats.cpp:403-429). Positions of the runs come
// from a two-level prefix sum of the lengths, every run sets its bits in the plane chunks, and K3 finds the
// MAXIMAL runs there, which is what the reference's RunBuffer makes of consecutive runs of one symbol.
constexpr int RUN_BYTES_PER_BLOCK = 1024;

__device__ __forceinline__ void decode_run_byte(uint32_t byte, int layout, uint32_t& comp, uint32_t& length)
{
  if(layout == BWTM_RUNS_SGA) { comp = byte >> 5; length = byte & 0x1Fu; }
  else { comp = byte & 7u; length = byte >> 3; }
}

__global__ void __launch_bounds__(256)
run_bytes_survey(const uint8_t* __restrict__ runs, uint64_t n_runs, int layout, unsigned long long* __restrict__ block_sums,
                 unsigned long long* __restrict__ invalid)
{
  __shared__ uint32_t warp_sums[8];
  uint32_t sum = 0, bad = 0;
  for(int k = 0; k < RUN_BYTES_PER_BLOCK / 256; k++)
  {
    uint64_t i = (uint64_t)blockIdx.x * RUN_BYTES_PER_BLOCK + k * 256 + threadIdx.x;
    if(i < n_runs)
    {
      uint32_t comp, length; decode_run_byte(runs[i], layout, comp, length);
      sum += length; bad += (length == 0 || comp >= (uint32_t)SIGMA ? 1u : 0u);
    }
  }
  if(bad != 0) { atomicAdd(invalid, (unsigned long long)bad); }
#pragma unroll
  for(int offset = 16; offset > 0; offset >>= 1) { sum += __shfl_down_sync(0xFFFFFFFFu, sum, offset); }
  if((threadIdx.x & 31) == 0) { warp_sums[threadIdx.x >> 5] = sum; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    uint32_t total = 0;
    for(int w = 0; w < 8; w++) { total += warp_sums[w]; }
    block_sums[blockIdx.x] = total;
  }
}

// Thread t of a block takes the bytes 4t .. 4t+3 of the block's 1024: positions by a block scan of the sums.
__global__ void __launch_bounds__(256)
run_bytes_fill(const uint8_t* __restrict__ runs, uint64_t n_runs, int layout, const unsigned long long* __restrict__ block_start,
               uint32_t* __restrict__ record_words)
{
  __shared__ uint32_t warp_sums[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t comp[4], length[4], sum = 0;
#pragma unroll
  for(int k = 0; k < 4; k++)
  {
    uint64_t i = (uint64_t)blockIdx.x * RUN_BYTES_PER_BLOCK + 4 * threadIdx.x + k;
    comp[k] = 0; length[k] = 0;
    if(i < n_runs) { decode_run_byte(runs[i], layout, comp[k], length[k]); }
    sum += length[k];
  }
  uint32_t inclusive = sum;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    uint32_t other = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= offset) { inclusive += other; }
  }
  if(lane == 31) { warp_sums[warp] = inclusive; }
  __syncthreads();
  uint64_t position = block_start[blockIdx.x] + inclusive - sum;
  for(int w = 0; w < warp; w++) { position += warp_sums[w]; }
#pragma unroll
  for(int k = 0; k < 4; k++)
  {
    uint64_t p = position; uint32_t remaining = length[k];
    position += length[k];
    if(comp[k] == 0) { continue; }
    while(remaining > 0)   // at most two words: a run has at most 31 symbols
    {
      uint32_t t = (uint32_t)(p & 31u);
      uint32_t take = (remaining < 32u - t ? remaining : 32u - t);
      uint32_t mask = low_mask((int)take) << t;
      uint32_t* chunk = record_words + (p >> 5) * 4;
      if(comp[k] & 1u) { atomicOr(chunk + 0, mask); }
      if(comp[k] & 2u) { atomicOr(chunk + 1, mask); }
      if(comp[k] & 4u) { atomicOr(chunk + 2, mask); }
      p += take; remaining -= take;
    }
  }
}

int index_from_run_bytes(const uint8_t* d_runs, uint64_t n_runs, int layout, uint64_t slab_symbols, cudaStream_t stream, bwtm_index** out)
{
  if(n_runs == 0) { set_error("empty sequence"); return BWTM_ERR_ARGUMENT; }
  if(layout != BWTM_RUNS_ROPEBWT && layout != BWTM_RUNS_SGA) { set_error("unknown run byte layout %d", layout); return BWTM_ERR_ARGUMENT; }
  const uint64_t blocks = div_up(n_runs, RUN_BYTES_PER_BLOCK);
  DeviceBuffer sums; BWTM_TRY(sums.allocate((blocks + 2) * sizeof(unsigned long long)));
  unsigned long long* block_start = sums.as<unsigned long long>();
  unsigned long long* invalid = block_start + blocks + 1;
  BWTM_CUDA(cudaMemsetAsync(block_start + blocks, 0, 2 * sizeof(unsigned long long), stream));
  run_bytes_survey<<<(unsigned)blocks, 256, 0, stream>>>(d_runs, n_runs, layout, block_start, invalid);
  BWTM_LAUNCH_CHECK();
  size_t temp_bytes = 0;
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, block_start, block_start, (int64_t)(blocks + 1), stream));
  DeviceBuffer temp; BWTM_TRY(temp.allocate(temp_bytes));
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(temp.ptr, temp_bytes, block_start, block_start, (int64_t)(blocks + 1), stream));
  count_launch(2);
  unsigned long long tail[2] = { 0, 0 };   // total symbols, invalid bytes
  BWTM_CUDA(cudaMemcpyAsync(tail, block_start + blocks, sizeof(tail), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  if(tail[1] != 0)
  {
    set_error("%llu run bytes have length 0 or a comp value above %d", tail[1], SIGMA - 1);
    return BWTM_ERR_ALPHABET;
  }
  const uint64_t n = tail[0];
  DeviceBuffer records;
  uint64_t record_bytes = ((n >> RECORD_SHIFT) + 1) * 64;
  BWTM_TRY(records.allocate(record_bytes, true));
  BWTM_CUDA(cudaMemsetAsync(records.ptr, 0, record_bytes, stream));
  run_bytes_fill<<<(unsigned)blocks, 256, 0, stream>>>(d_runs, n_runs, layout, block_start, records.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  return index_from_filled_records(records, n, slab_symbols, stream, out);
}

//------------------------------------------------------------------------------
// Search in batches (options.sequence_blocks): the counterpart of the reference's bounded buffers.
//
// The reference never holds the whole rank array as plain values: run buffers are sorted and compressed, merged
// into thread buffers and merge buffers and spilled to disk (fmi.cpp:164-257, fmi.h:45-80), and mergeBWT streams
// them back (support.h:576-638). On the device the array stays in HBM, but sorting it in one piece needs a second
// buffer of the same size. With S > 1 the sequences of b are searched in S batches; every batch is sorted on its own
// (the scratch buffer is batch-sized) and kept as a sorted run. The interleave then walks over ranges of A
// positions: the pieces of the S runs that fall into a range are gathered, merged pairwise and consumed at once.
// Peak memory: |b| keys + 2 batch-sized buffers instead of 2 |b| keys.

// Memory the pool could hand out right now: free device memory plus what the pool holds but does not use.
// cudaMemGetInfo costs milliseconds: only called when the answer can matter (see choose_batches).
static uint64_t available_device_bytes()
{
  size_t free_bytes = 0, total_bytes = 0;
  if(cudaMemGetInfo(&free_bytes, &total_bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
  uint64_t available = free_bytes;
  int device = 0; cudaMemPool_t pool;
  if(cudaGetDevice(&device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
  {
    uint64_t reserved = 0, used = 0;
    if(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
       cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used) { available += reserved - used; }
  }   // (what the pool of the indexes holds in reserve is not counted: it serves the result of the merge)
  cudaGetLastError();
  return available;
}

// Number of search batches: options.sequence_blocks when given (> 0); else 1 whenever the one-shot merge fits in the
// memory that is available now -- two key buffers, the records of the result, the output bytes and the slab work
// buffers -- and otherwise as many batches as make the two batch-sized buffers take an eighth of it. The free memory is
// only asked when the key buffers exceed an eighth of the device (small merges never pay for the query).
static uint64_t choose_batches(const bwtm_merge_options* options, const bwtm_index* a, const bwtm_index* b, uint64_t key_bytes)
{
  const uint64_t n_b = b->size;
  uint64_t batches = options->sequence_blocks;
  if(const char* env = getenv("BWTM_SEQUENCE_BLOCKS")) { batches = strtoull(env, nullptr, 10); }
  if(batches == 0)
  {
    batches = 1;
    if(2 * n_b * key_bytes > device_total_bytes() / 8)
    {
      const uint64_t available = available_device_bytes();
      const uint64_t rle_bytes = a->rle_bytes + b->rle_bytes;
      const uint64_t one_shot = 2 * n_b * key_bytes + (a->size + b->size) / 2 + rle_bytes + (rle_bytes >> 2) + (6ull << 30);
      if(available > 0 && one_shot > available - (available >> 4))
      {
        uint64_t batch_budget = std::max<uint64_t>(available / 16, 1ull << 28);
        batches = div_up(n_b * key_bytes, batch_budget);
      }
    }
  }
  return std::max<uint64_t>(1, std::min(batches, b->sequences));
}

template<class KeyT>
static int merge_in_batches(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options, uint64_t batches,
                            bwtm_index** result, bwtm_timings* timings)
{
  cudaStream_t stream = 0;
  const uint64_t n_a = a->size, n_b = b->size, m = b->sequences;
  const int S = (int)batches;
  const int bits = bit_length_host(n_a);
  EventTimer timer(stream);

  // 1. search and sort, batch by batch: S sorted runs in `runs`
  DeviceBuffer runs, scratch, counters;
  BWTM_TRY(runs.allocate(n_b * sizeof(KeyT)));
  BWTM_TRY(counters.allocate(walk_counters_bytes()));
  std::vector<unsigned long long> run_offsets(S + 1, 0);
  float search_ms = 0.0f, sort_ms = 0.0f;
  for(int k = 0; k < S; k++)
  {
    uint64_t first = (uint64_t)(((__uint128_t)m * k) / S), last = (uint64_t)(((__uint128_t)m * (k + 1)) / S);
    run_offsets[k + 1] = run_offsets[k];
    if(last == first) { continue; }
    uint64_t emitted = 0;
    timer.start();
    BWTM_TRY(walk_sequences<KeyT>(a, b, first, last - 1, runs.as<KeyT>() + run_offsets[k], n_b - run_offsets[k], &emitted, stream));
    search_ms += timer.stop();
    run_offsets[k + 1] = run_offsets[k] + emitted;
    if(emitted == 0) { continue; }
    timer.start();
    if(emitted * sizeof(KeyT) > scratch.bytes) { BWTM_TRY(scratch.allocate(emitted * sizeof(KeyT) + (emitted * sizeof(KeyT) >> 3))); }
    KeyT* sorted = nullptr;
    BWTM_TRY(sort_keys<KeyT>(runs.as<KeyT>() + run_offsets[k], scratch.as<KeyT>(), emitted, bits, &sorted, stream, n_a + 1));
    if(sorted != runs.as<KeyT>() + run_offsets[k])
    {
      BWTM_CUDA(cudaMemcpyAsync(runs.as<KeyT>() + run_offsets[k], sorted, emitted * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream));
    }
    sort_ms += timer.stop();
  }
  scratch.release();
  timings->search_seconds = search_ms * 1e-3; timings->sort_seconds = sort_ms * 1e-3;
  timings->walk_kernel_launches = S; timings->search_batches = S;
  timings->ra_values = run_offsets[S];
  if(run_offsets[S] != n_b)
  {
    set_error("the rank array has %llu values but the inserted BWT has %llu symbols: not a valid multi-string BWT",
              (unsigned long long)run_offsets[S], (unsigned long long)n_b);
    return BWTM_ERR_INTERNAL;
  }

  // 2. + 3. ranges of A positions, the pieces of the runs in each gathered, merged, interleaved and encoded
  DeviceBuffer distinct; BWTM_TRY(distinct.allocate(sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemsetAsync(distinct.ptr, 0, sizeof(unsigned long long), stream));
  DeviceBuffer control; BWTM_TRY(control.allocate(sizeof(EncodeControl)));
  BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream));
  OutputBuffer out = { nullptr, 0, 0, nullptr };
  int rc = ensure_capacity(&out, a->rle_bytes + b->rle_bytes + ((a->rle_bytes + b->rle_bytes) >> 2) + (1 << 20), 0, stream);
  float interleave_ms = 0.0f, encode_ms = 0.0f, merge_ms = 0.0f;
  if(rc == BWTM_OK)
  {
    rc = merge_ranges<KeyT>(a, b, runs.as<KeyT>(), run_offsets, 0, n_a + 1, 0, 0, n_a + n_b, options, &out, control.as<EncodeControl>(), true,
                            &merge_ms, &interleave_ms, &encode_ms, distinct.as<unsigned long long>(), stream);
  }
  if(rc != BWTM_OK) { device_free(out.ptr); return rc; }
  timings->sort_seconds += merge_ms * 1e-3;
  timings->interleave_seconds = interleave_ms * 1e-3; timings->encode_seconds = encode_ms * 1e-3;
  runs.release();

  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpy(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost));
  unsigned long long ra_runs = 0;
  BWTM_CUDA(cudaMemcpy(&ra_runs, distinct.ptr, sizeof(ra_runs), cudaMemcpyDeviceToHost));
  timings->ra_runs = ra_runs; timings->merged_runs = ctl.runs_total; timings->merged_bytes = ctl.out_size;
  if(options->host_output != nullptr)
  {
    if(options->host_output_capacity < ctl.out_size) { set_error("host_output is too small for %llu bytes", (unsigned long long)ctl.out_size); device_free(out.ptr); return BWTM_ERR_CAPACITY; }
    BWTM_CUDA(cudaMemcpyAsync(options->host_output, out.ptr, ctl.out_size, cudaMemcpyDeviceToHost, stream));
  }
  timer.start();
  uint64_t counts[SIGMA];
  for(int c = 0; c < SIGMA; c++) { counts[c] = a->counts[c] + b->counts[c]; }
  rc = finish_index(&out, ctl.out_size, counts, a->sequences + b->sequences, options->skip_index != 0, stream, result);
  timings->index_seconds = timer.stop() * 1e-3;
  device_free(out.ptr);
  return rc;
}

template<class KeyT>
static int merge_impl(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
                      bwtm_index** result, bwtm_timings* timings)
{
  cudaStream_t stream = 0;
  uint64_t n_b = b->size;
  BWTM_TRY(prepare_walk(a, b, b->size, stream, timings));
  const uint64_t batches = choose_batches(options, a, b, sizeof(KeyT));
  if(batches > 1) { return merge_in_batches<KeyT>(a, b, options, batches, result, timings); }
  timings->search_batches = 1;
  DeviceBuffer keys, alt;
  BWTM_TRY(keys.allocate(n_b * sizeof(KeyT)));
  BWTM_TRY(alt.allocate(n_b * sizeof(KeyT)));

  EventTimer timer(stream);
  KeyT* sorted = nullptr;
  {
    // The first partition level of the sort needs the histogram of its digit: the walk counts it while it writes.
    WalkHistogram histogram = { nullptr, 0, 1, nullptr, 0 };
    DeviceBuffer digit_counts;
    uint64_t fine_bins = 0;
    if(sort_plan_histogram(n_b, bit_length_host(a->size), a->size + 1, &histogram, &fine_bins))
    {
      const uint64_t counters = (fine_bins > 0 ? fine_bins : histogram.bins);
      BWTM_TRY(digit_counts.allocate(counters * sizeof(unsigned long long)));
      BWTM_CUDA(cudaMemsetAsync(digit_counts.ptr, 0, counters * sizeof(unsigned long long), stream));
      if(fine_bins > 0) { histogram.fine_counts = digit_counts.as<unsigned long long>(); }
      else { histogram.counts = digit_counts.as<unsigned long long>(); }
    }
    timer.start();
    uint64_t emitted = 0;
    BWTM_TRY(walk_sequences<KeyT>(a, b, 0, b->sequences - 1, keys.as<KeyT>(), n_b, &emitted, stream, &histogram));
    timings->search_seconds = timer.stop() * 1e-3;
    timings->walk_kernel_launches = 1;
    if(emitted != n_b)
    {
      set_error("the rank array has %llu values but the inserted BWT has %llu symbols: not a valid multi-string BWT",
                (unsigned long long)emitted, (unsigned long long)n_b);
      return BWTM_ERR_INTERNAL;
    }
    timings->ra_values = emitted;
    timer.start();
    BWTM_TRY(sort_keys<KeyT>(keys.as<KeyT>(), alt.as<KeyT>(), n_b, bit_length_host(a->size), &sorted, stream, a->size + 1, &histogram));
    timings->sort_seconds = timer.stop() * 1e-3;
  }
  if(sorted == keys.as<KeyT>()) { alt.release(); } else { keys.release(); }

  DeviceBuffer distinct; BWTM_TRY(distinct.allocate(sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemsetAsync(distinct.ptr, 0, sizeof(unsigned long long), stream));

  DeviceBuffer control; BWTM_TRY(control.allocate(sizeof(EncodeControl)));
  BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream));
  OutputBuffer out = { nullptr, 0, 0, nullptr };
  HostSink sink = { options->host_output, options->host_output_capacity, 0, nullptr, nullptr, false };
  if(options->host_output != nullptr)
  {
    cudaPointerAttributes attributes;
    if(cudaPointerGetAttributes(&attributes, options->host_output) == cudaSuccess && attributes.type == cudaMemoryTypeHost)
    {
      sink.device_visible = true;   // page-locked: kernels can write it directly
    }
    cudaGetLastError();
    BWTM_CUDA(cudaStreamCreateWithFlags(&sink.stream, cudaStreamNonBlocking));
    BWTM_CUDA(cudaEventCreateWithFlags(&sink.ready, cudaEventDisableTiming));
    out.sink = &sink;
  }
  auto close_sink = [&]()
  {
    if(sink.stream != nullptr) { cudaStreamSynchronize(sink.stream); cudaStreamDestroy(sink.stream); sink.stream = nullptr; }
    if(sink.ready != nullptr) { cudaEventDestroy(sink.ready); sink.ready = nullptr; }
    out.sink = nullptr;
  };
  int rc = ensure_capacity(&out, a->rle_bytes + b->rle_bytes + ((a->rle_bytes + b->rle_bytes) >> 2) + (1 << 20), 0, stream);
  float interleave_ms = 0.0f, encode_ms = 0.0f;
  // Records of the result: their plane words are filled slab by slab from the merged symbols, so K0 does
  // not have to decode the bytes K5 has just written. BWTM_RLE_INDEX=1 keeps the decoding route (tests).
  DeviceBuffer records;
  if(rc == BWTM_OK && options->skip_index == 0 && getenv("BWTM_RLE_INDEX") == nullptr)
  {
    uint64_t record_bytes = (((a->size + b->size) >> RECORD_SHIFT) + 1) * 64;
    rc = records.allocate(record_bytes, true);
    if(rc == BWTM_OK && cudaMemsetAsync(records.ptr, 0, record_bytes, stream) != cudaSuccess) { set_error("cannot clear the records"); rc = BWTM_ERR_CUDA; }
  }
  if(rc == BWTM_OK)
  {
    rc = interleave_range<KeyT>(a, b, sorted, 0, n_b, 0, a->size + b->size, options->slab_symbols,
                                &out, control.as<EncodeControl>(), true, &interleave_ms, &encode_ms, stream,
                                distinct.as<unsigned long long>(), records.as<uint4>());
  }
  if(rc != BWTM_OK) { close_sink(); device_free(out.ptr); return rc; }
  timings->interleave_seconds = interleave_ms * 1e-3;
  timings->encode_seconds = encode_ms * 1e-3;
  keys.release(); alt.release();

  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpy(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost));
  {
    unsigned long long ra_runs = 0;
    BWTM_CUDA(cudaMemcpy(&ra_runs, distinct.ptr, sizeof(ra_runs), cudaMemcpyDeviceToHost));
    timings->ra_runs = ra_runs;
  }
  timings->merged_runs = ctl.runs_total;
  timings->merged_bytes = ctl.out_size;

  timer.start();
  uint64_t counts[SIGMA];
  for(int c = 0; c < SIGMA; c++) { counts[c] = a->counts[c] + b->counts[c]; }
  rc = finish_index(&out, ctl.out_size, counts, a->sequences + b->sequences, options->skip_index != 0, stream, result,
                    &records, a->size + b->size);
  timings->index_seconds = timer.stop() * 1e-3;
  close_sink();              // the last part of the streaming download overlaps the index build
  device_free(out.ptr);
  return rc;
}

int merge_local(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
                bwtm_index** result, bwtm_timings* timings)
{
  // BWTM_FORCE_WIDE=1 runs the 64-bit key and position paths on small inputs (tests).
  if(a->size < 0xFFFFFFFFull && getenv("BWTM_FORCE_WIDE") == nullptr) { return merge_impl<uint32_t>(a, b, options, result, timings); }
  return merge_impl<uint64_t>(a, b, options, result, timings);
}

} // namespace bwtm

//------------------------------------------------------------------------------
// C ABI

using namespace bwtm;

extern "C"
{

int bwtm_merge(bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options, bwtm_index** out, bwtm_timings* timings)
{
  bwtm_merge_options defaults; std::memset(&defaults, 0, sizeof(defaults));
  if(options == nullptr) { options = &defaults; }
  bwtm_timings local; std::memset(&local, 0, sizeof(local));
  bool keep = (options->keep_inputs != 0);

  int rc = BWTM_OK;
  if(a == nullptr || b == nullptr || out == nullptr) { set_error("null argument"); rc = BWTM_ERR_ARGUMENT; }
  else if(a->d_records == nullptr || b->d_records == nullptr) { set_error("an input has no rank structure (it was built with skip_index)"); rc = BWTM_ERR_ARGUMENT; }
  else if(b->sequences == 0) { set_error("the inserted BWT has no sequences"); rc = BWTM_ERR_ARGUMENT; }
  if(rc == BWTM_OK)
  {
    *out = nullptr;
    uint64_t launches_before = bwtm_kernel_launches();
    auto start = std::chrono::steady_clock::now();
    rc = merge_local(a, b, options, out, &local);
    cudaDeviceSynchronize();
    local.total_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    local.kernel_launches = bwtm_kernel_launches() - launches_before;
  }
  if(!keep) { index_free(a); index_free(b); }
  if(timings != nullptr) { *timings = local; }
  return rc;
}

int bwtm_index_create_runs(const uint8_t* runs, uint64_t n_runs, int layout, uint64_t slab_symbols, bwtm_index** out)
{
  if(runs == nullptr || out == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    cudaGetLastError(); set_error("no CUDA device available; this library has no CPU fallback"); return BWTM_ERR_CUDA;
  }
  if(n_runs == 0) { set_error("empty sequence"); return BWTM_ERR_ARGUMENT; }
  DeviceBuffer d_runs; BWTM_TRY(d_runs.allocate(n_runs));
  BWTM_CUDA(cudaMemcpy(d_runs.ptr, runs, n_runs, cudaMemcpyHostToDevice));
  return index_from_run_bytes(d_runs.as<uint8_t>(), n_runs, layout, slab_symbols, 0, out);
}

int bwtm_index_create_plain(const uint8_t* comps, uint64_t n, uint64_t slab_symbols, bwtm_index** out)
{
  if(comps == nullptr || out == nullptr || n == 0) { set_error("null or empty argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    cudaGetLastError(); set_error("no CUDA device available; this library has no CPU fallback"); return BWTM_ERR_CUDA;
  }
  DeviceBuffer symbols; BWTM_TRY(symbols.allocate(n));
  BWTM_CUDA(cudaMemcpy(symbols.ptr, comps, n, cudaMemcpyHostToDevice));
  return index_from_symbols(symbols.as<uint8_t>(), n, slab_symbols, 0, out);
}

int bwtm_rank_array(const bwtm_index* a, const bwtm_index* b, uint64_t seq_first, uint64_t seq_last,
                    uint64_t* out_sorted, uint64_t capacity, uint64_t* n_values)
{
  if(a == nullptr || b == nullptr || out_sorted == nullptr || n_values == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  if(seq_first > seq_last || seq_last >= b->sequences) { set_error("invalid sequence range"); return BWTM_ERR_ARGUMENT; }
  if(a->d_records == nullptr || b->d_records == nullptr) { set_error("an input has no rank structure (it was built with skip_index)"); return BWTM_ERR_ARGUMENT; }
  DeviceBuffer keys, alt;
  BWTM_TRY(keys.allocate(capacity * sizeof(uint64_t)));
  BWTM_TRY(alt.allocate(capacity * sizeof(uint64_t)));
  uint64_t emitted = 0;
  {
    bwtm_timings unused; std::memset(&unused, 0, sizeof(unused));   // same choice of walk as a merge of these inputs
    BWTM_TRY(prepare_walk(const_cast<bwtm_index*>(a), const_cast<bwtm_index*>(b), b->size, 0, &unused));
  }
  BWTM_TRY(walk_sequences<uint64_t>(a, b, seq_first, seq_last, keys.as<uint64_t>(), capacity, &emitted, 0));
  uint64_t* sorted = nullptr;
  BWTM_TRY(sort_keys<uint64_t>(keys.as<uint64_t>(), alt.as<uint64_t>(), emitted, bit_length_host(a->size), &sorted, 0, a->size + 1));
  BWTM_CUDA(cudaMemcpy(out_sorted, sorted, emitted * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  *n_values = emitted;
  return BWTM_OK;
}

} // extern "C"
