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

__device__ __forceinline__ uint32_t fetch_symbol(const DeviceIndex& idx, uint64_t pos, uint64_t& cached_group, uint4& cached)
{
  uint64_t group = pos >> 5;
  if(group != cached_group) { cached = __ldg(idx.records + group); cached_group = group; }
  uint32_t t = (uint32_t)(pos & 31u);
  return ((cached.x >> t) & 1u) | (((cached.y >> t) & 1u) << 1) | (((cached.z >> t) & 1u) << 2);
}

template<class KeyT>
__global__ void __launch_bounds__(IL_THREADS)
k4_interleave(DeviceIndex a, DeviceIndex b, const KeyT* __restrict__ keys, uint64_t key_base,
              const uint64_t* __restrict__ tile_j, uint64_t begin, uint64_t end, uint8_t* __restrict__ merged,
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
  for(uint64_t k = tid; k < j1 - j0; k += IL_THREADS)
  {
    uint64_t j = j0 + k;
    KeyT key = keys[j - key_base];
    uint32_t q = (uint32_t)(j + (uint64_t)key - d0);
    atomicOr(&bitmap[q >> 5], 1u << (q & 31u));
    new_values += (j == key_base || keys[j - key_base - 1] != key) ? 1u : 0u;
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

  uint32_t w[4] = { 0, 0, 0, 0 };
  uint64_t group_a = ~0ull, group_b = ~0ull;
  uint4 chunk_a = make_uint4(0, 0, 0, 0), chunk_b = make_uint4(0, 0, 0, 0);
#pragma unroll
  for(int s = 0; s < PER_THREAD; s++)
  {
    if(s < valid)
    {
      uint32_t sym;
      if((flags >> s) & 1u) { sym = fetch_symbol(b, jb, group_b, chunk_b); jb++; }
      else                  { sym = fetch_symbol(a, ia, group_a, chunk_a); ia++; }
      w[s >> 2] |= sym << (8 * (s & 3));
    }
  }
  *reinterpret_cast<uint4*>(merged + (d0 - begin) + p0) = make_uint4(w[0], w[1], w[2], w[3]);
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
__device__ __forceinline__ uint32_t write_run(uint8_t* __restrict__ out, uint64_t offset, uint32_t comp, uint64_t length)
{
  uint64_t pos = offset;
  while(length > 0)
  {
    if(length < (uint64_t)MAX_RUN) { out[pos++] = (uint8_t)(comp + SIGMA * (length - 1)); break; }
    uint32_t remaining = RLE_BLOCK - (uint32_t)(pos & 63u);
    uint32_t basic = (remaining > 1 ? MAX_RUN : MAX_RUN - 1);
    out[pos++] = (uint8_t)(comp + SIGMA * (basic - 1)); length -= basic; remaining--;
    if(remaining > 0)
    {
      uint64_t extension = length;
      if(bytecode_length(extension) > remaining) { extension = (1ull << (7 * remaining)) - 1; }
      length -= extension;
      while(extension > 0x7Fu) { out[pos++] = (uint8_t)((extension & 0x7Fu) | 0x80u); extension >>= 7; }
      out[pos++] = (uint8_t)extension;
    }
  }
  return (uint32_t)(pos - offset);
}

// Sequential glue between slabs (and GPU slices): the RunBuffer state (utils.h:121-142). The pending run
// absorbs the slab's first run when the symbols agree; it is written as soon as the slab shows that it
// has ended; the slab's last run becomes the new pending run. The runs in between, [1, m - 1), never
// depend on the pending run: they are the parallel part.
__global__ void enc_head(EncodeControl* ctl, const uint8_t* __restrict__ sym, const uint32_t* __restrict__ start,
                         uint64_t m, uint8_t* __restrict__ out, int finish)
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
  if(m > 0)
  {
    if(ctl->carry_len > 0 && sym[0] == ctl->carry_sym) { ctl->carry_len += start[1] - start[0]; }
    else
    {
      if(ctl->carry_len > 0)
      {
        ctl->out_size += write_run(out, ctl->out_size, ctl->carry_sym, ctl->carry_len);
        ctl->runs_total++;
      }
      ctl->carry_sym = sym[0]; ctl->carry_len = start[1] - start[0];
    }
    if(m > 1)
    {
      ctl->out_size += write_run(out, ctl->out_size, ctl->carry_sym, ctl->carry_len);
      ctl->runs_total++;
      ctl->carry_sym = sym[m - 1]; ctl->carry_len = start[m] - start[m - 1];
      ctl->count = m - 2;
    }
  }
  ctl->slab_base = ctl->out_size;
}

// Runs are stored as (symbol, start position in the slab); start[m] = slab length closes the last run.
struct RunClass   // 1 in the low word for a short run, 1 in the high word for a long run
{
  const uint32_t* start;
  __host__ __device__ __forceinline__ unsigned long long operator()(const uint64_t& k) const
  {
    return (start[k + 1] - start[k] < (uint32_t)MAX_RUN ? 1ull : (1ull << 32));
  }
};
using RunClassIterator = cub::TransformInputIterator<unsigned long long, RunClass, cub::CountingInputIterator<uint64_t>>;

__global__ void enc_collect_long(unsigned long long* __restrict__ stats, const uint32_t* __restrict__ start, const unsigned long long* __restrict__ scan,
                                 uint64_t count, uint32_t* __restrict__ long_list)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= count) { return; }
  uint32_t length = start[k + 1] - start[k];
  unsigned long long s = scan[k];
  bool is_long = (length >= (uint32_t)MAX_RUN);
  if(is_long) { long_list[s >> 32] = (uint32_t)k; }
  if(k == count - 1)
  {
    stats[0] = (s & 0xFFFFFFFFull) + (is_long ? 0 : 1);   // short runs of the parallel part
    stats[1] = (s >> 32) + (is_long ? 1 : 0);             // long runs
  }
}

// Bytes of a long run at output offset `state` (mod 64), given its natural size (head + full extension):
// the natural encoding is used whenever it fits into the rest of the block (support.h:267-280).
__device__ __forceinline__ uint32_t long_run_bytes_fast(uint32_t length, uint32_t natural, uint32_t state)
{
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
enc_tile_maps(const uint32_t* __restrict__ start, const unsigned long long* __restrict__ scan,
              const uint32_t* __restrict__ long_list, uint64_t n_long, uint32_t* __restrict__ tile_bytes,
              uint16_t* __restrict__ checkpoints)
{
  __shared__ uint32_t s_len[LONG_SUB], s_nat[LONG_SUB], s_before[LONG_SUB];
  uint64_t first = (uint64_t)blockIdx.x * LONG_TILE;
  uint64_t last = (first + LONG_TILE < n_long ? first + LONG_TILE : n_long);
  uint32_t p = threadIdx.x;
  for(uint64_t chunk = first; chunk < last; chunk += LONG_SUB)
  {
    uint64_t sub = chunk / LONG_SUB;
    checkpoints[sub * 64 + threadIdx.x] = (uint16_t)(p - threadIdx.x);
    __syncthreads();
    if(threadIdx.x < LONG_SUB && chunk + threadIdx.x < last)
    {
      uint32_t idx = long_list[chunk + threadIdx.x];
      uint32_t length = start[idx + 1] - start[idx];
      s_len[threadIdx.x] = length; s_nat[threadIdx.x] = natural_bytes(length);
      s_before[threadIdx.x] = (uint32_t)scan[idx];
    }
    __syncthreads();
    int count = (int)(last - chunk < (uint64_t)LONG_SUB ? last - chunk : (uint64_t)LONG_SUB);
    for(int k = 0; k < count; k++)
    {
      uint32_t state = (s_before[k] + p) & 63u;
      p += long_run_bytes_fast(s_len[k], s_nat[k], state);
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
  __shared__ uint32_t staged[SCAN_CHUNK * 64];
  __shared__ unsigned long long entries[SCAN_CHUNK];
  __shared__ unsigned long long carried;
  const uint32_t base = (uint32_t)(ctl->slab_base & 63u);
  if(threadIdx.x == 0) { carried = 0; }
  __syncthreads();
  for(uint64_t first = 0; first < tiles; first += SCAN_CHUNK)
  {
    uint64_t count = (tiles - first < (uint64_t)SCAN_CHUNK ? tiles - first : (uint64_t)SCAN_CHUNK);
    for(uint64_t k = threadIdx.x; k < count * 64; k += blockDim.x) { staged[k] = tile_bytes[first * 64 + k]; }
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
__global__ void enc_long_offsets(const EncodeControl* ctl, const uint32_t* __restrict__ start, const unsigned long long* __restrict__ scan,
                                 const uint32_t* __restrict__ long_list, uint64_t n_long,
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
    uint32_t idx = long_list[k];
    uint32_t length = start[idx + 1] - start[idx];
    long_offset[k] = (uint32_t)p;
    uint32_t state = (base_state + (uint32_t)scan[idx] + (uint32_t)p) & 63u;
    p += long_run_bytes_fast(length, natural_bytes(length), state);
  }
}

__global__ void enc_write(const EncodeControl* ctl, const uint8_t* __restrict__ sym, const uint32_t* __restrict__ start,
                          const unsigned long long* __restrict__ scan, const uint32_t* __restrict__ long_offset,
                          uint64_t count, uint8_t* __restrict__ out)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= count) { return; }
  unsigned long long s = scan[k];
  uint64_t longs_before = s >> 32;
  uint64_t bytes_before = (s & 0xFFFFFFFFull) + (longs_before < ctl->n_long ? (uint64_t)long_offset[longs_before] : ctl->long_bytes);
  uint64_t offset = ctl->slab_base + bytes_before;
  uint32_t length = start[k + 1] - start[k], comp = sym[k];
  if(length < (uint32_t)MAX_RUN) { out[offset] = (uint8_t)(comp + SIGMA * (length - 1)); }
  else { write_run(out, offset, comp, length); }
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
  for(uint64_t i = tid; i < vectors; i += threads) { dst[i] = src[i]; }
  for(uint64_t i = head + vectors * 16 + tid; i < bytes; i += threads) { host[i] = device[i]; }
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
    copy_to_host<<<64, 256, 0, sink->stream>>>(sink->ptr + sink->copied, out->ptr + sink->copied, final_bytes - sink->copied);
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

// K3: maximal runs of a symbol array (what the reference's RunBuffer produces, utils.h:121-142).
// A run starts wherever a symbol differs from its predecessor. Pass 1 counts the run starts of every
// RUN_TILE-symbol tile, a scan turns the counts into offsets, pass 2 writes (symbol, start) for every run.
constexpr int RUN_THREADS = 256;
constexpr int RUN_TILE    = RUN_THREADS * 16;

// Bit i of the result is set when symbol i of the 16 differs from the one before it (`previous` for i = 0).
__device__ __forceinline__ uint32_t run_start_flags(const uint4& q, uint32_t previous)
{
  uint32_t w[4] = { q.x, q.y, q.z, q.w };
  uint32_t flags = 0;
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    uint32_t shifted = (w[j] << 8) | (j == 0 ? (previous & 0xFFu) : (w[j - 1] >> 24));
    uint32_t diff = w[j] ^ shifted;
    uint32_t nonzero = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;   // high bit of every differing byte
    // gather the four high bits into a nibble
    uint32_t nibble = ((nonzero >> 7) & 1u) | ((nonzero >> 14) & 2u) | ((nonzero >> 21) & 4u) | ((nonzero >> 28) & 8u);
    flags |= nibble << (4 * j);
  }
  return flags;
}

// Loads the 16 symbols of slot `slot` (zero-padded beyond n) and the symbol before them (0xFF at the start).
__device__ __forceinline__ uint4 load_symbols16(const uint8_t* __restrict__ symbols, uint64_t n, uint64_t slot, uint32_t& previous, int& valid)
{
  uint64_t first = slot * 16;
  valid = (first >= n ? 0 : (n - first < 16 ? (int)(n - first) : 16));
  previous = (first == 0 || first > n ? 0xFFu : (uint32_t)symbols[first - 1]);
  uint4 q = make_uint4(0, 0, 0, 0);
  if(valid == 16) { q = *reinterpret_cast<const uint4*>(symbols + first); }
  else if(valid > 0)
  {
    uint32_t w[4] = { 0, 0, 0, 0 };
    for(int i = 0; i < valid; i++) { w[i >> 2] |= (uint32_t)symbols[first + i] << (8 * (i & 3)); }
    q = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return q;
}

__global__ void __launch_bounds__(RUN_THREADS)
run_tile_counts(const uint8_t* __restrict__ symbols, uint64_t n, uint32_t* __restrict__ tile_counts)
{
  __shared__ uint32_t warp_sums[RUN_THREADS / 32];
  uint32_t previous; int valid;
  uint4 q = load_symbols16(symbols, n, (uint64_t)blockIdx.x * RUN_THREADS + threadIdx.x, previous, valid);
  uint32_t flags = run_start_flags(q, previous) & ((1u << valid) - 1u);
  uint32_t count = __popc(flags);
#pragma unroll
  for(int offset = 16; offset > 0; offset >>= 1) { count += __shfl_down_sync(0xFFFFFFFFu, count, offset); }
  if((threadIdx.x & 31) == 0) { warp_sums[threadIdx.x >> 5] = count; }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    uint32_t total = 0;
    for(int w = 0; w < RUN_THREADS / 32; w++) { total += warp_sums[w]; }
    tile_counts[blockIdx.x] = total;
  }
}

__global__ void __launch_bounds__(RUN_THREADS)
run_tile_write(const uint8_t* __restrict__ symbols, uint64_t n, const uint32_t* __restrict__ tile_offsets,
               uint8_t* __restrict__ run_sym, uint32_t* __restrict__ run_start)
{
  __shared__ uint32_t warp_sums[RUN_THREADS / 32];
  uint32_t previous; int valid;
  uint64_t slot = (uint64_t)blockIdx.x * RUN_THREADS + threadIdx.x;
  uint4 q = load_symbols16(symbols, n, slot, previous, valid);
  uint32_t flags = run_start_flags(q, previous) & ((1u << valid) - 1u);
  uint32_t count = __popc(flags);
  // exclusive prefix of the counts within the block
  uint32_t inclusive = count;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for(int offset = 1; offset < 32; offset <<= 1)
  {
    uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inclusive, offset);
    if(lane >= offset) { inclusive += v; }
  }
  if(lane == 31) { warp_sums[threadIdx.x >> 5] = inclusive; }
  __syncthreads();
  uint32_t before = tile_offsets[blockIdx.x] + inclusive - count;
  for(int w = 0; w < (int)(threadIdx.x >> 5); w++) { before += warp_sums[w]; }
  uint32_t words[4] = { q.x, q.y, q.z, q.w };
  while(flags != 0)
  {
    int i = __ffs(flags) - 1; flags &= flags - 1;
    run_sym[before] = (uint8_t)(words[i >> 2] >> (8 * (i & 3)));
    run_start[before] = (uint32_t)(slot * 16 + i);
    before++;
  }
}

int SlabEncoder::init(uint64_t max_symbols_, cudaStream_t stream)
{
  max_symbols = max_symbols_;
  run_capacity = 0;
  BWTM_TRY(num_runs.allocate(4 * sizeof(uint64_t)));
  BWTM_TRY(placed.allocate(sizeof(EncodeControl)));
  uint64_t tiles = div_up(max_symbols, RUN_TILE) + 1;
  BWTM_TRY(run_tiles.allocate(tiles * sizeof(uint32_t)));
  size_t tile_temp = 0, scan_temp = 0;
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tile_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)tiles, stream));
  {
    RunClassIterator classes(cub::CountingInputIterator<uint64_t>(0), RunClass{ nullptr });
    BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, classes, (unsigned long long*)nullptr, (int64_t)max_symbols, stream));
  }
  BWTM_TRY(cub_temp.allocate(std::max(tile_temp, scan_temp)));
  return BWTM_OK;
}

// Work arrays sized by the number of runs of the slab (a fifth of the symbols for read collections), grow-only.
int SlabEncoder::reserve_runs(uint64_t runs)
{
  if(runs <= run_capacity) { return BWTM_OK; }
  uint64_t capacity = runs + (runs >> 3) + 1024;
  uint64_t max_long = std::min(capacity, max_symbols / MAX_RUN + 1);
  uint64_t max_long_tiles = div_up(max_long, LONG_TILE);
  BWTM_TRY(run_sym.allocate(capacity));
  BWTM_TRY(run_start.allocate((capacity + 1) * sizeof(uint32_t)));
  BWTM_TRY(scan.allocate(capacity * sizeof(unsigned long long)));
  BWTM_TRY(long_list.allocate(max_long * sizeof(uint32_t)));
  BWTM_TRY(long_offset.allocate(max_long * sizeof(uint32_t)));
  BWTM_TRY(tile_bytes.allocate(max_long_tiles * 64 * sizeof(uint32_t)));
  BWTM_TRY(tile_entry.allocate(max_long_tiles * sizeof(unsigned long long)));
  BWTM_TRY(checkpoints.allocate((max_long / LONG_SUB + 1) * 64 * sizeof(uint16_t)));
  run_capacity = capacity;
  return BWTM_OK;
}

// K3 and the state-free half of K5: maximal runs of `symbols` consecutive symbols, and for the runs
// [1, m - 1) the short/long scan, the compacted long runs and their transducer tile maps.
int SlabEncoder::detect(const uint8_t* d_symbols, uint64_t symbols, cudaStream_t stream)
{
  detected_runs = 0; part_count = 0; part_short = 0; part_long = 0;
  if(symbols == 0) { return BWTM_OK; }
  if(symbols > max_symbols) { set_error("slab of %llu symbols exceeds the encoder capacity", (unsigned long long)symbols); return BWTM_ERR_INTERNAL; }
  size_t temp_bytes = cub_temp.bytes;
  BWTM_CUDA(cudaMemsetAsync(num_runs.ptr, 0, 4 * sizeof(uint64_t), stream));
  uint64_t tiles = div_up(symbols, RUN_TILE);
  run_tile_counts<<<(unsigned)tiles, RUN_THREADS, 0, stream>>>(d_symbols, symbols, run_tiles.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemsetAsync(run_tiles.as<uint32_t>() + tiles, 0, sizeof(uint32_t), stream));
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp.ptr, temp_bytes, run_tiles.as<uint32_t>(), run_tiles.as<uint32_t>(), (int64_t)(tiles + 1), stream));
  count_launch(2);
  uint32_t total_runs = 0;
  BWTM_CUDA(cudaMemcpyAsync(&total_runs, run_tiles.as<uint32_t>() + tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t m = total_runs;
  if(m == 0) { set_error("run detection found no runs in %llu symbols", (unsigned long long)symbols); return BWTM_ERR_INTERNAL; }
  BWTM_TRY(this->reserve_runs(m));
  run_tile_write<<<(unsigned)tiles, RUN_THREADS, 0, stream>>>(d_symbols, symbols, run_tiles.as<uint32_t>(), run_sym.as<uint8_t>(), run_start.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  uint32_t slab_length = (uint32_t)symbols;
  BWTM_CUDA(cudaMemcpyAsync(run_start.as<uint32_t>() + m, &slab_length, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  detected_runs = m;
  if(m < 3) { return BWTM_OK; }

  uint64_t count = m - 2;
  const uint32_t* len = run_start.as<uint32_t>() + 1;   // starts of the parallel part, runs [1, m - 1)
  unsigned long long* stats = num_runs.as<unsigned long long>() + 1;
  RunClassIterator classes(cub::CountingInputIterator<uint64_t>(0), RunClass{ len });
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp.ptr, temp_bytes, classes, scan.as<unsigned long long>(), (int64_t)count, stream));
  count_launch(2);
  enc_collect_long<<<(unsigned)div_up(count, 256), 256, 0, stream>>>(stats, len, scan.as<unsigned long long>(), count, long_list.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  unsigned long long host_stats[2] = { 0, 0 };
  BWTM_CUDA(cudaMemcpyAsync(host_stats, stats, sizeof(host_stats), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  part_count = count; part_short = host_stats[0]; part_long = host_stats[1];
  uint64_t long_tiles = div_up(part_long, LONG_TILE);
  if(long_tiles > 0)
  {
    enc_tile_maps<<<(unsigned)long_tiles, 64, 0, stream>>>(len, scan.as<unsigned long long>(), long_list.as<uint32_t>(), part_long,
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
  enc_head<<<1, 1, 0, stream>>>(d_control, run_sym.as<uint8_t>(), run_start.as<uint32_t>(), m, out->at_origin(), 0);
  BWTM_LAUNCH_CHECK();
  if(part_count == 0) { return BWTM_OK; }
  uint64_t long_tiles = div_up(part_long, LONG_TILE);
  enc_tile_scan<<<1, 256, 0, stream>>>(d_control, tile_bytes.as<uint32_t>(), long_tiles, part_short, part_long, tile_entry.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpyAsync(placed.ptr, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToDevice, stream));
  return BWTM_OK;
}

// Second step: writes the bytes of the parallel part at the offsets fixed by advance().
int SlabEncoder::emit(OutputBuffer* out, cudaStream_t stream)
{
  if(detected_runs == 0 || part_count == 0) { return BWTM_OK; }
  const EncodeControl* where = placed.as<EncodeControl>();
  const uint8_t* sym = run_sym.as<uint8_t>() + 1;
  const uint32_t* len = run_start.as<uint32_t>() + 1;
  if(part_long > 0)
  {
    enc_long_offsets<<<(unsigned)div_up(div_up(part_long, LONG_SUB), 128), 128, 0, stream>>>(
      where, len, scan.as<unsigned long long>(), long_list.as<uint32_t>(), part_long,
      tile_entry.as<unsigned long long>(), checkpoints.as<uint16_t>(), long_offset.as<uint32_t>());
    BWTM_LAUNCH_CHECK();
  }
  enc_write<<<(unsigned)div_up(part_count, 256), 256, 0, stream>>>(where, sym, len, scan.as<unsigned long long>(), long_offset.as<uint32_t>(), part_count, out->at_origin());
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

int SlabEncoder::write(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  BWTM_TRY(this->advance(out, d_control, stream));
  return this->emit(out, stream);
}

int SlabEncoder::encode(const uint8_t* d_symbols, uint64_t symbols, OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  BWTM_TRY(this->detect(d_symbols, symbols, stream));
  return this->write(out, d_control, stream);
}

// Flushes the pending run (bwt.cpp:279-281).
int SlabEncoder::finish(OutputBuffer* out, EncodeControl* d_control, cudaStream_t stream)
{
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, d_control, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  BWTM_TRY(ensure_capacity(out, ctl.out_size + 256, ctl.out_size, stream));
  enc_head<<<1, 1, 0, stream>>>(d_control, nullptr, nullptr, 0, out->at_origin(), 1);
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
                    uint64_t p0, uint64_t p1, uint8_t* d_merged, uint64_t* d_tile_j, cudaStream_t stream,
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

template int interleave_slab<uint32_t>(const bwtm_index*, const bwtm_index*, const uint32_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint8_t*, uint64_t*, cudaStream_t, unsigned long long*);
template int interleave_slab<uint64_t>(const bwtm_index*, const bwtm_index*, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, uint8_t*, uint64_t*, cudaStream_t, unsigned long long*);

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
  BWTM_TRY(merged.allocate(slab_symbols));
  BWTM_TRY(tile_j.allocate((max_tiles + 2) * sizeof(uint64_t)));
  SlabEncoder encoder;
  BWTM_TRY(encoder.init(slab_symbols, stream));
  EventTimer timer(stream);

  for(uint64_t p0 = begin; p0 < end; p0 += slab_symbols)
  {
    uint64_t p1 = std::min(p0 + slab_symbols, end);
    timer.start();
    BWTM_TRY(interleave_slab<KeyT>(a, b, d_keys, key_base, key_count, p0, p1, merged.as<uint8_t>(), tile_j.as<uint64_t>(), stream, d_distinct_keys));
    // The rank structure of the result takes its plane words straight from the merged symbols.
    if(d_result_records != nullptr) { BWTM_TRY(planes_from_symbols(merged.as<uint8_t>(), p0, p1 - p0, d_result_records, stream)); }
    *interleave_ms += timer.stop();

    timer.start();
    BWTM_TRY(encoder.encode(merged.as<uint8_t>(), p1 - p0, out, d_control, stream));
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
  DeviceBuffer exact; BWTM_TRY(exact.allocate(rle_bytes + RLE_PADDING));
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

int index_from_symbols(const uint8_t* d_symbols, uint64_t n, uint64_t slab_symbols, cudaStream_t stream, bwtm_index** out)
{
  if(n == 0) { set_error("empty sequence"); return BWTM_ERR_ARGUMENT; }
  uint64_t slab = clamp_slab(slab_symbols, n);
  SlabEncoder encoder; BWTM_TRY(encoder.init(slab, stream));
  DeviceBuffer control; BWTM_TRY(control.allocate(sizeof(EncodeControl)));
  BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream));
  OutputBuffer buffer = { nullptr, 0, 0, nullptr };
  int rc = ensure_capacity(&buffer, n / 4 + (1 << 20), 0, stream);
  // The records take their plane words from the symbols when those can be read 32 at a time.
  DeviceBuffer records;
  if((reinterpret_cast<uintptr_t>(d_symbols) & 15) == 0 && (slab & 31) == 0)
  {
    uint64_t record_bytes = ((n >> RECORD_SHIFT) + 1) * 64;
    BWTM_TRY(records.allocate(record_bytes));
    BWTM_CUDA(cudaMemsetAsync(records.ptr, 0, record_bytes, stream));
  }
  for(uint64_t p0 = 0; rc == BWTM_OK && p0 < n; p0 += slab)
  {
    uint64_t count = std::min(slab, n - p0);
    rc = encoder.encode(d_symbols + p0, count, &buffer, control.as<EncodeControl>(), stream);
    if(rc == BWTM_OK && records.ptr != nullptr) { rc = planes_from_symbols(d_symbols + p0, p0, count, records.as<uint4>(), stream); }
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

// K1 and K2 overlapped. The walk is bound by random line requests and leaves most of the HBM bandwidth
// idle; the radix sort is the opposite. B's sequences are walked in chunks on one stream (with a reduced
// number of resident blocks per SM) while finished chunks are sorted and merged pairwise on a second
// stream. The RA multiset does not depend on how B's sequences are split (fmi.cpp:355).
template<class KeyT>
static int walk_sort_pipelined(const bwtm_index* a, const bwtm_index* b, KeyT* keys, KeyT* alt, uint64_t n_b, int bits, int chunks,
                               KeyT** sorted, bwtm_timings* timings)
{
  const uint64_t m = b->sequences;
  const uint64_t counter_stride = div_up(walk_counters_bytes(), 64) * 64;
  int walk_blocks = 6;
  if(const char* env = getenv("BWTM_WALK_CTAS")) { walk_blocks = atoi(env); }

  DeviceBuffer counters, cursor, sort_temp, merge_temp;
  BWTM_TRY(counters.allocate(counter_stride * chunks));
  BWTM_TRY(cursor.allocate(sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemsetAsync(counters.ptr, 0, counter_stride * chunks, 0));
  BWTM_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(unsigned long long), 0));
  uint64_t max_chunk = std::min(n_b, (n_b / chunks) * 2 + (1 << 20));
  size_t sort_bytes = 0, merge_bytes = 0;
  {
    cub::DoubleBuffer<KeyT> probe(keys, alt);
    BWTM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, probe, (int64_t)n_b, 0, bits, 0));
    BWTM_CUDA(cub::DeviceMerge::MergeKeys(nullptr, merge_bytes, keys, (int)std::min<uint64_t>(n_b, 0x3FFFFFFF), keys, (int)std::min<uint64_t>(n_b, 0x3FFFFFFF), alt, ::cuda::std::less<>{}, 0));
  }
  (void)max_chunk;
  BWTM_TRY(sort_temp.allocate(sort_bytes)); BWTM_TRY(merge_temp.allocate(merge_bytes));

  std::vector<unsigned char> host_counters(counter_stride * chunks, 0);
  std::vector<unsigned long long> host_cursor(chunks, 0);
  unsigned long long* pinned = nullptr;
  BWTM_CUDA(cudaMallocHost(&pinned, sizeof(unsigned long long) * chunks + counter_stride * chunks));
  unsigned char* pinned_counters = reinterpret_cast<unsigned char*>(pinned + chunks);

  cudaStream_t s_walk = nullptr, s_sort = nullptr;
  cudaEvent_t ready = nullptr, begin = nullptr, walked_all = nullptr, done = nullptr;
  std::vector<cudaEvent_t> walked(chunks, nullptr);
  int rc = BWTM_OK;
  auto cleanup = [&]()
  {
    if(s_walk) { cudaStreamSynchronize(s_walk); cudaStreamDestroy(s_walk); }
    if(s_sort) { cudaStreamSynchronize(s_sort); cudaStreamDestroy(s_sort); }
    for(cudaEvent_t e : walked) { if(e) { cudaEventDestroy(e); } }
    for(cudaEvent_t e : { ready, begin, walked_all, done }) { if(e) { cudaEventDestroy(e); } }
    if(pinned) { cudaFreeHost(pinned); }
  };
#define BWTM_PIPE(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) { rc = cuda_failed(e__, #call, __FILE__, __LINE__); cleanup(); return rc; } } while(0)
#define BWTM_PIPE_TRY(call) do { rc = (call); if(rc != BWTM_OK) { cleanup(); return rc; } } while(0)

  BWTM_PIPE(cudaStreamCreateWithFlags(&s_walk, cudaStreamNonBlocking));
  BWTM_PIPE(cudaStreamCreateWithFlags(&s_sort, cudaStreamNonBlocking));
  BWTM_PIPE(cudaEventCreate(&ready)); BWTM_PIPE(cudaEventCreate(&begin));
  BWTM_PIPE(cudaEventCreate(&walked_all)); BWTM_PIPE(cudaEventCreate(&done));
  for(int c = 0; c < chunks; c++) { BWTM_PIPE(cudaEventCreateWithFlags(&walked[c], cudaEventDisableTiming)); }
  BWTM_PIPE(cudaEventRecord(ready, 0));
  BWTM_PIPE(cudaStreamWaitEvent(s_walk, ready, 0)); BWTM_PIPE(cudaStreamWaitEvent(s_sort, ready, 0));
  BWTM_PIPE(cudaEventRecord(begin, s_walk));

  // All walks are enqueued up front; each leaves its end offset and counters in pinned memory.
  for(int c = 0; c < chunks; c++)
  {
    uint64_t first = (uint64_t)(((__uint128_t)m * c) / chunks), last = (uint64_t)(((__uint128_t)m * (c + 1)) / chunks);
    if(last > first)
    {
      BWTM_PIPE_TRY(walk_sequences_async<KeyT>(a, b, first, last - 1, keys, n_b, counters.as<unsigned char>() + counter_stride * c,
                                               cursor.as<unsigned long long>(), walk_blocks, s_walk));
    }
    BWTM_PIPE(cudaMemcpyAsync(pinned + c, cursor.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_walk));
    BWTM_PIPE(cudaMemcpyAsync(pinned_counters + counter_stride * c, counters.as<unsigned char>() + counter_stride * c, counter_stride,
                              cudaMemcpyDeviceToHost, s_walk));
    BWTM_PIPE(cudaEventRecord(walked[c], s_walk));
  }
  BWTM_PIPE(cudaEventRecord(walked_all, s_walk));
  timings->walk_kernel_launches = chunks;

  // Sort every chunk as it arrives; merge equal-level neighbours (a binary counter of sorted runs).
  struct SortedRun { uint64_t begin, end; int level; KeyT* where; };
  std::vector<SortedRun> runs;
  uint64_t previous_end = 0;
  for(int c = 0; c < chunks; c++)
  {
    BWTM_PIPE(cudaEventSynchronize(walked[c]));
    if(walk_counters_check(pinned_counters + counter_stride * c) || pinned[c] > n_b)
    {
      set_error("the rank array has more than %llu values: the inserted BWT is not a valid multi-string BWT", (unsigned long long)n_b);
      rc = BWTM_ERR_INTERNAL; cleanup(); return rc;
    }
    uint64_t chunk_begin = previous_end, chunk_end = pinned[c];
    previous_end = chunk_end;
    if(chunk_end == chunk_begin) { continue; }
    cub::DoubleBuffer<KeyT> buffers(keys + chunk_begin, alt + chunk_begin);
    size_t bytes = sort_temp.bytes;
    BWTM_PIPE(cub::DeviceRadixSort::SortKeys(sort_temp.ptr, bytes, buffers, (int64_t)(chunk_end - chunk_begin), 0, bits, s_sort));
    count_launch((uint64_t)(2 + (bits + 7) / 8));
    KeyT* base = (buffers.Current() == keys + chunk_begin ? keys : alt);
    runs.push_back({ chunk_begin, chunk_end, 0, base });
    while(runs.size() >= 2 && runs[runs.size() - 1].level == runs[runs.size() - 2].level)
    {
      SortedRun right = runs.back(); runs.pop_back();
      SortedRun left = runs.back(); runs.pop_back();
      KeyT* target = (left.where == keys ? alt : keys);
      if(right.where != left.where)   // different pass parity cannot happen (same bit count), but stay safe
      {
        BWTM_PIPE(cudaMemcpyAsync(left.where + right.begin, right.where + right.begin, (right.end - right.begin) * sizeof(KeyT), cudaMemcpyDeviceToDevice, s_sort));
      }
      size_t mb = merge_temp.bytes;
      BWTM_PIPE(cub::DeviceMerge::MergeKeys(merge_temp.ptr, mb, left.where + left.begin, (int)(left.end - left.begin),
                                            left.where + right.begin, (int)(right.end - right.begin), target + left.begin,
                                            ::cuda::std::less<>{}, s_sort));
      count_launch(2);
      runs.push_back({ left.begin, right.end, left.level + 1, target });
    }
  }
  // Leftover runs of different levels (chunk count not a power of two, or empty chunks): merge right to left.
  while(runs.size() >= 2)
  {
    SortedRun right = runs.back(); runs.pop_back();
    SortedRun left = runs.back(); runs.pop_back();
    KeyT* target = (left.where == keys ? alt : keys);
    if(right.where != left.where)
    {
      BWTM_PIPE(cudaMemcpyAsync(left.where + right.begin, right.where + right.begin, (right.end - right.begin) * sizeof(KeyT), cudaMemcpyDeviceToDevice, s_sort));
    }
    size_t mb = merge_temp.bytes;
    BWTM_PIPE(cub::DeviceMerge::MergeKeys(merge_temp.ptr, mb, left.where + left.begin, (int)(left.end - left.begin),
                                          left.where + right.begin, (int)(right.end - right.begin), target + left.begin,
                                          ::cuda::std::less<>{}, s_sort));
    count_launch(2);
    runs.push_back({ left.begin, right.end, std::max(left.level, right.level) + 1, target });
  }
  BWTM_PIPE(cudaEventRecord(done, s_sort));
  BWTM_PIPE(cudaEventSynchronize(done));
  float walk_ms = 0.0f, total_ms = 0.0f;
  cudaEventElapsedTime(&walk_ms, begin, walked_all);
  cudaEventElapsedTime(&total_ms, begin, done);
  timings->search_seconds = walk_ms * 1e-3;
  timings->sort_seconds = std::max(0.0f, total_ms - walk_ms) * 1e-3;   // what the overlap did not hide
  timings->ra_values = previous_end;
  *sorted = (runs.empty() ? keys : runs.back().where);
  uint64_t emitted = previous_end;
  cleanup();
#undef BWTM_PIPE
#undef BWTM_PIPE_TRY
  if(emitted != n_b)
  {
    set_error("the rank array has %llu values but the inserted BWT has %llu symbols: not a valid multi-string BWT",
              (unsigned long long)emitted, (unsigned long long)n_b);
    return BWTM_ERR_INTERNAL;
  }
  return BWTM_OK;
}

template<class KeyT>
static int merge_impl(const bwtm_index* a, const bwtm_index* b, const bwtm_merge_options* options,
                      bwtm_index** result, bwtm_timings* timings)
{
  cudaStream_t stream = 0;
  uint64_t n_b = b->size;
  DeviceBuffer keys, alt;
  BWTM_TRY(keys.allocate(n_b * sizeof(KeyT)));
  BWTM_TRY(alt.allocate(n_b * sizeof(KeyT)));

  EventTimer timer(stream);
  KeyT* sorted = nullptr;
  // Overlapping the walk with the sort of finished chunks is implemented (walk_sort_pipelined) but off by
  // default: on B200 the two kernels slow each other down by as much as the overlap hides (config 2: 93 ms
  // either way, profiles/r01_pipeline_sweep.txt). BWTM_PIPELINE_CHUNKS=<n> turns it on.
  int chunks = 1;
  if(const char* env = getenv("BWTM_PIPELINE_CHUNKS")) { chunks = atoi(env); }
  uint64_t pipeline_min = 1ull << 26;   // below this the overlap does not pay for the extra launches
  if(const char* env = getenv("BWTM_PIPELINE_MIN")) { pipeline_min = strtoull(env, nullptr, 10); }
  if(chunks > 1 && n_b >= pipeline_min && n_b < 0x7FFFFFFFull && b->sequences >= (uint64_t)chunks)
  {
    BWTM_TRY(walk_sort_pipelined<KeyT>(a, b, keys.as<KeyT>(), alt.as<KeyT>(), n_b, bit_length_host(a->size), chunks, &sorted, timings));
  }
  else
  {
    timer.start();
    uint64_t emitted = 0;
    BWTM_TRY(walk_sequences<KeyT>(a, b, 0, b->sequences - 1, keys.as<KeyT>(), n_b, &emitted, stream));
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
    BWTM_TRY(sort_keys<KeyT>(keys.as<KeyT>(), alt.as<KeyT>(), n_b, bit_length_host(a->size), &sorted, stream));
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
    rc = records.allocate(record_bytes);
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

int merge_local(const bwtm_index* a, const bwtm_index* b, const bwtm_merge_options* options,
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
  DeviceBuffer keys, alt;
  BWTM_TRY(keys.allocate(capacity * sizeof(uint64_t)));
  BWTM_TRY(alt.allocate(capacity * sizeof(uint64_t)));
  uint64_t emitted = 0;
  BWTM_TRY(walk_sequences<uint64_t>(a, b, seq_first, seq_last, keys.as<uint64_t>(), capacity, &emitted, 0));
  uint64_t* sorted = nullptr;
  BWTM_TRY(sort_keys<uint64_t>(keys.as<uint64_t>(), alt.as<uint64_t>(), emitted, 64, &sorted, 0));
  BWTM_CUDA(cudaMemcpy(out_sorted, sorted, emitted * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  *n_values = emitted;
  return BWTM_OK;
}

} // extern "C"
