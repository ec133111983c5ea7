// K0: device rank structure from run-length bytes, and the batched query entry points.
//
// Replaces BWT::build (bwt.cpp:476-512): the reference scans the run bytes once and stores, per
// 64-byte block, the last sequence position and six cumulative counts in seven sparse bitvectors.
// Here the run bytes are decoded once into position-addressed 64-byte records (bwtm_common.cuh).
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "bwtm_internal.cuh"

namespace bwtm
{

//------------------------------------------------------------------------------
// Errors, launch counter

static thread_local std::string last_error_message;
static std::atomic<uint64_t>* launch_counter()
{
  static std::atomic<uint64_t> counter(0);
  return &counter;
}

void set_error(const char* fmt, ...)
{
  char buffer[1024];
  va_list args; va_start(args, fmt);
  vsnprintf(buffer, sizeof(buffer), fmt, args);
  va_end(args);
  last_error_message = buffer;
}

int cuda_failed(cudaError_t err, const char* what, const char* file, int line)
{
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)err, cudaGetErrorString(err), what, file, line);
  return (err == cudaErrorMemoryAllocation ? BWTM_ERR_MEMORY : BWTM_ERR_CUDA);
}

void count_launch(uint64_t n) { launch_counter()->fetch_add(n); }

// Total memory of the current device, asked once per device (cudaMemGetInfo takes milliseconds).
uint64_t device_total_bytes()
{
  static uint64_t totals[64] = { 0 };
  int device = 0;
  if(cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) { cudaGetLastError(); return 0; }
  if(totals[device] == 0)
  {
    size_t free_bytes = 0, total_bytes = 0;
    if(cudaMemGetInfo(&free_bytes, &total_bytes) == cudaSuccess) { totals[device] = total_bytes; }
    cudaGetLastError();
  }
  return totals[device];
}

// Device memory comes from two stream-ordered pools with an unlimited release threshold: buffers freed by one merge
// are reused by the next one instead of going back to the driver (cudaMalloc/cudaFree of the multi-GB work buffers cost
// more than the kernels they serve).
//   * the device's default pool holds the WORK buffers of a merge (keys, scratch, slabs);
//   * a second pool holds what outlives the call that made it: the buffers of an index (run-length bytes, records,
//     superblock tables, pair records).
// With one pool the long-lived result of every merge was carved out of the block that had served the largest work
// buffer, the next merge found no block of that size any more, and the pool kept remapping memory: steady-state
// merges of config 5 on two GPUs took 0.5 - 4.8 s instead of 0.3 s (profiles/r02_bench_c5_n2_one_pool.json).
static int configure_pool()
{
  static thread_local int configured_device = -1;
  int device = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  if(device == configured_device) { return BWTM_OK; }
  cudaMemPool_t pool;
  BWTM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = UINT64_MAX;
  BWTM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  configured_device = device;
  return BWTM_OK;
}

static std::mutex resident_pool_mutex;
static cudaMemPool_t resident_pools[64] = { nullptr };

static int resident_pool(cudaMemPool_t* pool)
{
  int device = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  if(device < 0 || device >= 64) { set_error("device ordinal out of range"); return BWTM_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(resident_pool_mutex);
  if(resident_pools[device] == nullptr)
  {
    cudaMemPoolProps props; std::memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t created;
    BWTM_CUDA(cudaMemPoolCreate(&created, &props));
    uint64_t threshold = UINT64_MAX;
    BWTM_CUDA(cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold, &threshold));
    resident_pools[device] = created;
  }
  *pool = resident_pools[device];
  return BWTM_OK;
}

// Experiment switch BWTM_RESIDENT: "pool" (default) = the second pool, "default" = everything from the default pool,
// "malloc" = cudaMalloc for the buffers of an index.
static int resident_mode()
{
  static int mode = -1;
  if(mode < 0)
  {
    const char* env = getenv("BWTM_RESIDENT");
    mode = (env == nullptr ? 0 : (env[0] == 'd' ? 1 : (env[0] == 'm' ? 2 : 0)));
  }
  return mode;
}

static std::mutex plain_allocations_mutex;
static std::vector<void*> plain_allocations;   // pointers that came from cudaMalloc
static std::vector<uint64_t> plain_sizes;
static uint64_t plain_bytes = 0;

// cudaMalloc instead of a pool: for the few large structures that are read at random for a long time (pair records
// built ahead of time), where the plain allocation's mapping is measurably faster and its cost is paid once.
int device_alloc_plain(void** ptr, uint64_t n)
{
  *ptr = nullptr;
  if(n == 0) { n = 16; }
  cudaError_t err = cudaMalloc(ptr, n);
  if(err != cudaSuccess)
  {
    *ptr = nullptr; cudaGetLastError();
    set_error("device allocation of %llu bytes failed: %s", (unsigned long long)n, cudaGetErrorString(err));
    return BWTM_ERR_MEMORY;
  }
  std::lock_guard<std::mutex> lock(plain_allocations_mutex);
  plain_allocations.push_back(*ptr); plain_sizes.push_back(n); plain_bytes += n;
  return BWTM_OK;
}

int device_alloc(void** ptr, uint64_t n, bool resident)
{
  *ptr = nullptr;
  if(n == 0) { n = 16; }
  BWTM_TRY(configure_pool());
  cudaError_t err;
  if(resident && resident_mode() == 2) { return device_alloc_plain(ptr, n); }
  if(resident && resident_mode() == 0)
  {
    cudaMemPool_t pool;
    BWTM_TRY(resident_pool(&pool));
    err = cudaMallocFromPoolAsync(ptr, n, pool, 0);
  }
  else { err = cudaMallocAsync(ptr, n, 0); }
  if(err != cudaSuccess)
  {
    *ptr = nullptr;
    set_error("device allocation of %llu bytes failed: %s", (unsigned long long)n, cudaGetErrorString(err));
    cudaGetLastError();
    return BWTM_ERR_MEMORY;
  }
  return BWTM_OK;
}

void device_free(void* ptr)
{
  if(ptr == nullptr) { return; }
  {
    std::lock_guard<std::mutex> lock(plain_allocations_mutex);
    for(size_t k = 0; k < plain_allocations.size(); k++)
    {
      if(plain_allocations[k] == ptr)
      {
        plain_bytes -= plain_sizes[k];
        plain_allocations[k] = plain_allocations.back(); plain_allocations.pop_back();
        plain_sizes[k] = plain_sizes.back(); plain_sizes.pop_back();
        cudaFree(ptr);
        return;
      }
    }
  }
  cudaFreeAsync(ptr, 0);
}

int DeviceBuffer::allocate(uint64_t n, bool resident)
{
  this->release();
  BWTM_TRY(device_alloc(&(this->ptr), n, resident));
  this->bytes = (n == 0 ? 16 : n);
  return BWTM_OK;
}

void DeviceBuffer::release()
{
  if(this->ptr != nullptr) { device_free(this->ptr); this->ptr = nullptr; this->bytes = 0; }
}

//------------------------------------------------------------------------------
// Run decoding of one 64-byte block held in shared memory (Run::read, support.h:244-250;
// ByteCode::read, support.h:172-184). `limit` is the number of valid bytes in the block.

constexpr int K0_THREADS = 128;
constexpr int K0_STRIDE  = 68;   // bytes per staged block: 17 words, so that lanes hit distinct banks

template<class Visitor>
__device__ __forceinline__ void decode_staged_block(const uint8_t* block, int limit, Visitor&& visit)
{
  int pos = 0;
  while(pos < limit)
  {
    uint32_t code = block[pos++];
    uint32_t comp = code % SIGMA;
    uint64_t length = code / SIGMA + 1;
    if(length >= MAX_RUN)
    {
      int shift = 0;
      while(pos < limit)
      {
        uint32_t b = block[pos++];
        if(shift < 64) { length += (uint64_t)(b & 0x7Fu) << shift; }
        shift += 7;
        if(!(b & 0x80u)) { break; }
      }
    }
    visit(comp, length);
  }
}

// Coalesced copy of K0_THREADS consecutive 64-byte blocks into padded shared memory.
__device__ __forceinline__ void stage_blocks(const uint8_t* rle, uint64_t first_block, uint64_t rle_bytes, uint8_t* smem)
{
  // The rle buffer is padded with zeros to a multiple of 16 beyond rle_bytes, so 16-byte loads are safe.
  const uint4* src = reinterpret_cast<const uint4*>(rle + first_block * RLE_BLOCK);
  uint64_t total_vec = div_up(rle_bytes, 16);
  uint64_t base_vec = first_block * (RLE_BLOCK / 16);
  for(int v = threadIdx.x; v < K0_THREADS * (RLE_BLOCK / 16); v += K0_THREADS)
  {
    uint4 value = make_uint4(0, 0, 0, 0);
    if(base_vec + v < total_vec) { value = src[v]; }
    int block = v >> 2, part = v & 3;
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + block * K0_STRIDE + part * 16);
    dst[0] = value.x; dst[1] = value.y; dst[2] = value.z; dst[3] = value.w;
  }
}

__global__ void __launch_bounds__(K0_THREADS)
k0_block_lengths(const uint8_t* __restrict__ rle, uint64_t rle_bytes, uint64_t blocks, uint64_t* __restrict__ lengths)
{
  __shared__ __align__(16) uint8_t smem[K0_THREADS * K0_STRIDE];
  uint64_t first_block = (uint64_t)blockIdx.x * K0_THREADS;
  stage_blocks(rle, first_block, rle_bytes, smem);
  __syncthreads();

  uint64_t block = first_block + threadIdx.x;
  if(block >= blocks) { return; }
  uint64_t begin = block * RLE_BLOCK;
  int limit = (int)(rle_bytes - begin < (uint64_t)RLE_BLOCK ? rle_bytes - begin : (uint64_t)RLE_BLOCK);
  uint64_t total = 0;
  decode_staged_block(smem + threadIdx.x * K0_STRIDE, limit, [&](uint32_t, uint64_t length) { total += length; });
  lengths[block] = total;
}

// Sets the plane bits of every run. Records must be zero-initialised.
//
// A CTA decodes K0_THREADS consecutive 64-byte blocks, i.e. one contiguous range of positions. When the
// range fits, the three planes of the range are assembled in shared memory (shared atomics), the 16-byte
// chunks that lie entirely inside the range are written with plain coalesced stores and only the first and
// last chunk, which neighbouring CTAs may share, go through global atomics. Ranges that do not fit (very long
// runs) fall back to global atomics for every word.
constexpr int K0_PLANE_WORDS = 3072;   // 32-position chunks per CTA range held in shared memory

__device__ __forceinline__ void set_run_bits(uint32_t* p0, uint32_t* p1, uint32_t* p2, uint64_t stride,
                                             uint64_t first_chunk, uint32_t comp, uint64_t pos, uint64_t length)
{
  uint64_t p = pos, remaining = length;
  while(remaining > 0)
  {
    uint64_t index = ((p >> 5) - first_chunk) * stride;
    uint32_t t = (uint32_t)(p & 31u);
    uint32_t take = (uint32_t)(remaining < (uint64_t)(32 - t) ? remaining : (uint64_t)(32 - t));
    uint32_t mask = low_mask((int)take) << t;
    if(comp & 1u) { atomicOr(p0 + index, mask); }
    if(comp & 2u) { atomicOr(p1 + index, mask); }
    if(comp & 4u) { atomicOr(p2 + index, mask); }
    p += take; remaining -= take;
  }
}

__global__ void __launch_bounds__(K0_THREADS)
k0_fill_planes(const uint8_t* __restrict__ rle, uint64_t rle_bytes, uint64_t blocks,
               const uint64_t* __restrict__ starts, uint32_t* __restrict__ record_words)
{
  __shared__ __align__(16) uint8_t smem[K0_THREADS * K0_STRIDE];
  __shared__ uint32_t planes[3][K0_PLANE_WORDS];
  uint64_t first_block = (uint64_t)blockIdx.x * K0_THREADS;
  stage_blocks(rle, first_block, rle_bytes, smem);

  uint64_t last_block = (first_block + K0_THREADS < blocks ? first_block + K0_THREADS : blocks);
  uint64_t range_begin = starts[first_block], range_end = starts[last_block];
  uint64_t first_chunk = range_begin >> 5;
  uint64_t chunks = (range_end > range_begin ? ((range_end - 1) >> 5) - first_chunk + 1 : 0);
  bool staged = (chunks <= (uint64_t)K0_PLANE_WORDS);
  if(staged)
  {
    for(uint64_t w = threadIdx.x; w < chunks; w += K0_THREADS) { planes[0][w] = 0; planes[1][w] = 0; planes[2][w] = 0; }
  }
  __syncthreads();

  uint64_t block = first_block + threadIdx.x;
  if(block < blocks)
  {
    uint64_t begin = block * RLE_BLOCK;
    int limit = (int)(rle_bytes - begin < (uint64_t)RLE_BLOCK ? rle_bytes - begin : (uint64_t)RLE_BLOCK);
    uint64_t pos = starts[block];
    if(staged)
    {
      // The runs of a block cover consecutive positions: the bits of the current 32-position word are
      // collected in registers. A word the thread fills from bit 0 to bit 31 belongs to it alone and is
      // stored; the words at the two ends of its range are shared with the neighbouring threads (atomicOr).
      uint32_t word = (uint32_t)((pos >> 5) - first_chunk), t = (uint32_t)(pos & 31u);
      uint32_t acc0 = 0, acc1 = 0, acc2 = 0;
      bool owned = (t == 0);
      decode_staged_block(smem + threadIdx.x * K0_STRIDE, limit, [&](uint32_t comp, uint64_t length)
      {
        const uint32_t b0 = 0u - (comp & 1u), b1 = 0u - ((comp >> 1) & 1u), b2 = 0u - ((comp >> 2) & 1u);
        uint64_t remaining = length;
        while(remaining > 0)
        {
          uint32_t take = (uint32_t)(remaining < (uint64_t)(32u - t) ? remaining : (uint64_t)(32u - t));
          uint32_t mask = low_mask((int)take) << t;
          acc0 |= mask & b0; acc1 |= mask & b1; acc2 |= mask & b2;
          t += take; remaining -= take;
          if(t == 32u)
          {
            if(owned) { planes[0][word] = acc0; planes[1][word] = acc1; planes[2][word] = acc2; }
            else
            {
              if(acc0 != 0) { atomicOr(&planes[0][word], acc0); }
              if(acc1 != 0) { atomicOr(&planes[1][word], acc1); }
              if(acc2 != 0) { atomicOr(&planes[2][word], acc2); }
            }
            word++; t = 0; acc0 = 0; acc1 = 0; acc2 = 0; owned = true;
          }
        }
      });
      if(t != 0)
      {
        if(acc0 != 0) { atomicOr(&planes[0][word], acc0); }
        if(acc1 != 0) { atomicOr(&planes[1][word], acc1); }
        if(acc2 != 0) { atomicOr(&planes[2][word], acc2); }
      }
    }
    else
    {
      decode_staged_block(smem + threadIdx.x * K0_STRIDE, limit, [&](uint32_t comp, uint64_t length)
      {
        if(comp != 0) { set_run_bits(record_words, record_words + 1, record_words + 2, 4, 0, comp, pos, length); }
        pos += length;
      });
    }
  }
  if(!staged) { return; }
  __syncthreads();

  for(uint64_t w = threadIdx.x; w < chunks; w += K0_THREADS)
  {
    uint32_t a = planes[0][w], b = planes[1][w], c = planes[2][w];
    uint32_t* chunk = record_words + (first_chunk + w) * 4;
    if(w > 0 && w + 1 < chunks) { *reinterpret_cast<uint4*>(chunk) = make_uint4(a, b, c, 0); }
    else
    {
      if(a != 0) { atomicOr(chunk + 0, a); }
      if(b != 0) { atomicOr(chunk + 1, b); }
      if(c != 0) { atomicOr(chunk + 2, c); }
    }
  }
}

// Per-record counts of comps 1..5 (positions beyond `size` hold comp 0 and are never counted).
__global__ void k0_record_counts(const uint4* __restrict__ records, uint64_t n_records, uint32_t* __restrict__ counts)
{
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= n_records) { return; }
  uint32_t local[SIGMA] = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    uint4 q = records[4 * r + j];
#pragma unroll
    for(uint32_t c = 1; c < SIGMA; c++) { local[c] += __popc(match_mask(q, c)); }
  }
#pragma unroll
  for(uint32_t c = 1; c < SIGMA; c++) { counts[(c - 1) * n_records + r] = local[c]; }
}

// Packs the five 25-bit superblock-relative counters into the header words and writes the
// superblock table.
__global__ void k0_pack_headers(uint32_t* __restrict__ record_words, uint64_t n_records,
                                const uint64_t* __restrict__ cumulative /* 5 x n_records, exclusive */,
                                uint64_t* __restrict__ super)
{
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= n_records) { return; }
  uint64_t sb = r >> SUPER_RECORD_SHIFT;
  uint64_t sb_record = sb << SUPER_RECORD_SHIFT;
  uint64_t rel[SIGMA];
  uint64_t others = 0;
#pragma unroll
  for(uint32_t c = 1; c < SIGMA; c++)
  {
    uint64_t base = cumulative[(c - 1) * n_records + sb_record];
    rel[c] = cumulative[(c - 1) * n_records + r] - base;
    others += base;
  }
  uint64_t lo = rel[1] | (rel[2] << 25) | (rel[3] << 50);
  uint64_t hi = (rel[3] >> 14) | (rel[4] << 11) | (rel[5] << 36);
  uint32_t* words = record_words + r * 16;
  words[3]  = (uint32_t)lo;
  words[7]  = (uint32_t)(lo >> 32);
  words[11] = (uint32_t)hi;
  words[15] = (uint32_t)(hi >> 32);
  if(r == sb_record)
  {
    uint64_t* row = super + sb * SUPER_STRIDE;
    row[0] = (sb_record << RECORD_SHIFT) - others;
#pragma unroll
    for(uint32_t c = 1; c < SIGMA; c++) { row[c] = cumulative[(c - 1) * n_records + sb_record]; }
    row[6] = 0; row[7] = 0;
  }
}

struct CastU64
{
  __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t& x) const { return (uint64_t)x; }
};

int rle_block_starts(const uint8_t* d_rle, uint64_t rle_bytes, uint64_t* d_starts, cudaStream_t stream)
{
  uint64_t blocks = div_up(rle_bytes, RLE_BLOCK);
  BWTM_CUDA(cudaMemsetAsync(d_starts, 0, (blocks + 1) * sizeof(uint64_t), stream));
  if(blocks == 0) { return BWTM_OK; }
  k0_block_lengths<<<(unsigned)div_up(blocks, K0_THREADS), K0_THREADS, 0, stream>>>(d_rle, rle_bytes, blocks, d_starts);
  BWTM_LAUNCH_CHECK();
  // Exclusive scan over blocks + 1 entries (the last input is 0): d_starts[blocks] = total length.
  size_t temp_bytes = 0;
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, d_starts, d_starts, blocks + 1, stream));
  DeviceBuffer temp; BWTM_TRY(temp.allocate(temp_bytes));
  BWTM_CUDA(cub::DeviceScan::ExclusiveSum(temp.ptr, temp_bytes, d_starts, d_starts, blocks + 1, stream));
  count_launch(2);
  BWTM_CUDA(cudaStreamSynchronize(stream));
  return BWTM_OK;
}

void index_free(bwtm_index* index)
{
  if(index == nullptr) { return; }
  device_free(index->d_rle);
  device_free(index->d_records);
  device_free(index->d_super);
  device_free(index->d_pairs);
  device_free(index->d_pair_super);
  delete index;
}

// Plane words of 32 consecutive symbols held one per byte (the interleave's output): bit i of plane k is
// bit k of symbol i. Four symbols of a 32-bit word are gathered with one multiply.
__global__ void k0_planes_from_symbols(const uint8_t* __restrict__ symbols, uint64_t first_position, uint64_t count,
                                       uint4* __restrict__ records)
{
  uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(chunk * 32 >= count) { return; }
  uint32_t planes[3] = { 0, 0, 0 };
  uint64_t remaining = count - chunk * 32;
  if(remaining < 32)   // the ragged end of the sequence: nothing is read or set past it
  {
    for(uint32_t i = 0; i < (uint32_t)remaining; i++)
    {
      uint32_t value = symbols[chunk * 32 + i];
      planes[0] |= (value & 1u) << i; planes[1] |= ((value >> 1) & 1u) << i; planes[2] |= ((value >> 2) & 1u) << i;
    }
    records[(first_position >> 5) + chunk] = make_uint4(planes[0], planes[1], planes[2], 0);
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(symbols + chunk * 32);
  uint4 lo = src[0], hi = src[1];
  uint32_t w[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
#pragma unroll
  for(int j = 0; j < 8; j++)
  {
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      uint32_t nibble = ((((w[j] >> k) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
      planes[k] |= nibble << (4 * j);
    }
  }
  records[(first_position >> 5) + chunk] = make_uint4(planes[0], planes[1], planes[2], 0);
}

int planes_from_symbols(const uint8_t* d_symbols, uint64_t first_position, uint64_t count, uint4* d_records, cudaStream_t stream)
{
  if(count == 0) { return BWTM_OK; }
  if((first_position & 31) != 0) { set_error("slab not aligned to 32 positions"); return BWTM_ERR_INTERNAL; }
  uint64_t chunks = div_up(count, 32);
  k0_planes_from_symbols<<<(unsigned)div_up(chunks, 256), 256, 0, stream>>>(d_symbols, first_position, count, d_records);
  BWTM_LAUNCH_CHECK();
  return BWTM_OK;
}

// Second half of K0: per-record counts, their scans, headers and the superblock table, for records whose
// planes are filled. Takes ownership of d_rle and d_records on success.
int index_from_planes(uint8_t* d_rle, uint64_t rle_bytes, uint4* d_records, uint64_t size, cudaStream_t stream, bwtm_index** out)
{
  uint64_t n_records = (size >> RECORD_SHIFT) + 1;
  uint64_t n_super = ((n_records - 1) >> SUPER_RECORD_SHIFT) + 1;
  DeviceBuffer super; BWTM_TRY(super.allocate(n_super * SUPER_STRIDE * sizeof(uint64_t), true));
  BWTM_CUDA(cudaMemsetAsync(super.ptr, 0, n_super * SUPER_STRIDE * sizeof(uint64_t), stream));

  DeviceBuffer counts; BWTM_TRY(counts.allocate(5 * n_records * sizeof(uint32_t)));
  DeviceBuffer cumulative; BWTM_TRY(cumulative.allocate(5 * n_records * sizeof(uint64_t)));
  k0_record_counts<<<(unsigned)div_up(n_records, 256), 256, 0, stream>>>(d_records, n_records, counts.as<uint32_t>());
  BWTM_LAUNCH_CHECK();

  size_t temp_bytes = 0;
  {
    cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> in(counts.as<uint32_t>(), CastU64());
    BWTM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, in, cumulative.as<uint64_t>(), n_records, stream));
  }
  DeviceBuffer temp; BWTM_TRY(temp.allocate(temp_bytes));
  for(int c = 0; c < 5; c++)
  {
    cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> in(counts.as<uint32_t>() + c * n_records, CastU64());
    BWTM_CUDA(cub::DeviceScan::ExclusiveSum(temp.ptr, temp_bytes, in, cumulative.as<uint64_t>() + c * n_records, n_records, stream));
    count_launch(2);
  }

  k0_pack_headers<<<(unsigned)div_up(n_records, 256), 256, 0, stream>>>(
    reinterpret_cast<uint32_t*>(d_records), n_records, cumulative.as<uint64_t>(), super.as<uint64_t>());
  BWTM_LAUNCH_CHECK();

  // Totals: exclusive prefix of the last record + its own counts.
  uint64_t last_cum[5]; uint32_t last_cnt[5];
  BWTM_CUDA(cudaStreamSynchronize(stream));
  for(int c = 0; c < 5; c++)
  {
    BWTM_CUDA(cudaMemcpy(&last_cum[c], cumulative.as<uint64_t>() + c * n_records + (n_records - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
    BWTM_CUDA(cudaMemcpy(&last_cnt[c], counts.as<uint32_t>() + c * n_records + (n_records - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }

  bwtm_index* index = new bwtm_index();
  std::memset(index, 0, sizeof(bwtm_index));
  BWTM_CUDA(cudaGetDevice(&(index->device)));
  index->rle_bytes = rle_bytes;
  index->n_records = n_records; index->n_super = n_super;
  index->size = size;
  uint64_t others = 0;
  for(int c = 1; c < SIGMA; c++) { index->counts[c] = last_cum[c - 1] + last_cnt[c - 1]; others += index->counts[c]; }
  index->counts[0] = size - others;
  index->sequences = index->counts[0];
  index->C[0] = 0;
  for(int c = 0; c < SIGMA; c++) { index->C[c + 1] = index->C[c] + index->counts[c]; }
  index->device_bytes = rle_bytes + RLE_PADDING + n_records * 64 + super.bytes;
  index->d_rle = d_rle;
  index->d_records = d_records;
  index->d_super = static_cast<uint64_t*>(super.detach());
  *out = index;
  return BWTM_OK;
}

int index_from_device_rle(uint8_t* d_rle, uint64_t rle_bytes, cudaStream_t stream, bwtm_index** out)
{
  if(rle_bytes == 0) { set_error("empty BWT"); return BWTM_ERR_ARGUMENT; }
  uint64_t blocks = div_up(rle_bytes, RLE_BLOCK);

  DeviceBuffer starts; BWTM_TRY(starts.allocate((blocks + 1) * sizeof(uint64_t)));
  BWTM_TRY(rle_block_starts(d_rle, rle_bytes, starts.as<uint64_t>(), stream));
  uint64_t size = 0;
  BWTM_CUDA(cudaMemcpy(&size, starts.as<uint64_t>() + blocks, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if(size == 0) { set_error("BWT decodes to an empty sequence"); return BWTM_ERR_ARGUMENT; }

  uint64_t n_records = (size >> RECORD_SHIFT) + 1;
  DeviceBuffer records; BWTM_TRY(records.allocate(n_records * 64, true));
  BWTM_CUDA(cudaMemsetAsync(records.ptr, 0, n_records * 64, stream));
  k0_fill_planes<<<(unsigned)div_up(blocks, K0_THREADS), K0_THREADS, 0, stream>>>(
    d_rle, rle_bytes, blocks, starts.as<uint64_t>(), records.as<uint32_t>());
  BWTM_LAUNCH_CHECK();
  starts.release();
  BWTM_TRY(index_from_planes(d_rle, rle_bytes, records.as<uint4>(), size, stream, out));
  records.detach();
  return BWTM_OK;
}

//------------------------------------------------------------------------------
// Query kernels

__global__ void query_rank(DeviceIndex idx, const uint64_t* __restrict__ positions, const uint8_t* __restrict__ comps,
                           uint64_t n, uint64_t* __restrict__ out)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) { return; }
  out[k] = rank_any(idx, positions[k], comps[k]);
}

// FMI::LF(i): (C[c] + rank(i, c), c) with c = BWT[i]; (0, 0) for i >= size (bwt.cpp:448-449 returns
// rank 0 and comp 0, to which LF adds C[0] = 0).
__global__ void query_lf(DeviceIndex idx, const uint64_t* __restrict__ positions, uint64_t n,
                         uint64_t* __restrict__ out_positions, uint8_t* __restrict__ out_comps)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) { return; }
  uint64_t i = positions[k];
  if(i >= idx.size) { out_positions[k] = 0; out_comps[k] = 0; return; }
  uint32_t comp;
  uint64_t next = lf_step(idx, i, comp);
  if(comp == 0) { next = rank_any(idx, i, 0); }
  out_positions[k] = next; out_comps[k] = (uint8_t)comp;
}

// K6: FMI::find (fmi.h:195-209), one thread per pattern; writes Range::length of the result.
__global__ void query_count(DeviceIndex idx, const uint8_t* __restrict__ patterns, const uint64_t* __restrict__ offsets,
                            uint64_t n, const uint8_t* __restrict__ char2comp, uint64_t* __restrict__ out)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) { return; }
  uint64_t begin = offsets[k], end = offsets[k + 1];
  if(begin == end) { out[k] = idx.size; return; }   // range (0, size - 1)
  end--;
  uint32_t c = patterns[end]; if(char2comp != nullptr) { c = char2comp[c]; }
  if(c >= SIGMA) { out[k] = 0; return; }
  uint64_t first = idx.C[c], second = idx.C[c + 1] - 1;
  while(first + 1 <= second + 1 && end != begin)    // !Range::empty(range), utils.h:80-83
  {
    end--;
    c = patterns[end]; if(char2comp != nullptr) { c = char2comp[c]; }
    if(c >= SIGMA) { first = 1; second = 0; break; }
    uint64_t f = idx.C[c] + rank_any(idx, first, c);
    uint64_t s = idx.C[c] + rank_any(idx, second + 1, c) - 1;
    first = f; second = s;
  }
  out[k] = second + 1 - first;
}

__global__ void extract_symbols(DeviceIndex idx, uint64_t first, uint64_t count, uint8_t* __restrict__ out)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= count) { return; }
  uint64_t i = first + k;
  uint4 q = idx.records[4 * (i >> RECORD_SHIFT) + ((i >> 5) & 3)];
  uint32_t t = (uint32_t)(i & 31u);
  out[k] = (uint8_t)(((q.x >> t) & 1u) | (((q.y >> t) & 1u) << 1) | (((q.z >> t) & 1u) << 2));
}

// Block samples of BWT::build: cumulative counts through RLE block k = rank(start of block k + 1, c).
__global__ void sample_blocks(DeviceIndex idx, const uint64_t* __restrict__ starts, uint64_t blocks,
                              uint64_t* __restrict__ block_ends, uint64_t* __restrict__ cumulative)
{
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= blocks) { return; }
  uint64_t next = starts[k + 1];
  block_ends[k] = next - 1;
  for(uint32_t c = 0; c < SIGMA; c++) { cumulative[c * blocks + k] = rank_any(idx, next, c); }
}

} // namespace bwtm

//------------------------------------------------------------------------------
// C ABI: library, index, queries

using namespace bwtm;

static int check_device()
{
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if(err != cudaSuccess || count == 0)
  {
    cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback",
              err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    return BWTM_ERR_CUDA;
  }
  return BWTM_OK;
}

extern "C"
{

const char* bwtm_last_error(void) { return last_error_message.c_str(); }
const char* bwtm_version(void) { return "bwtm_b200 0.1 (sm_100a)"; }

int bwtm_device_count(int* count)
{
  if(count == nullptr) { set_error("count is NULL"); return BWTM_ERR_ARGUMENT; }
  cudaError_t err = cudaGetDeviceCount(count);
  if(err != cudaSuccess) { *count = 0; return cuda_failed(err, "cudaGetDeviceCount", __FILE__, __LINE__); }
  return BWTM_OK;
}

int bwtm_set_device(int device)
{
  BWTM_TRY(check_device());
  BWTM_CUDA(cudaSetDevice(device));
  return BWTM_OK;
}

uint64_t bwtm_kernel_launches(void) { return launch_counter()->load(); }

int bwtm_index_create_device(const void* rle_device, uint64_t rle_bytes, bwtm_index** out)
{
  if(rle_device == nullptr || out == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  BWTM_TRY(check_device());
  DeviceBuffer rle; BWTM_TRY(rle.allocate(rle_bytes + RLE_PADDING, true));
  BWTM_CUDA(cudaMemcpy(rle.ptr, rle_device, rle_bytes, cudaMemcpyDeviceToDevice));
  BWTM_CUDA(cudaMemset(rle.as<uint8_t>() + rle_bytes, 0, RLE_PADDING));
  BWTM_TRY(index_from_device_rle(rle.as<uint8_t>(), rle_bytes, 0, out));
  rle.detach();
  return BWTM_OK;
}

int bwtm_index_create(const uint8_t* rle, uint64_t rle_bytes, const uint64_t* expected_counts, bwtm_index** out)
{
  if(rle == nullptr || out == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  BWTM_TRY(check_device());
  DeviceBuffer d_rle; BWTM_TRY(d_rle.allocate(rle_bytes + RLE_PADDING, true));
  BWTM_CUDA(cudaMemcpy(d_rle.ptr, rle, rle_bytes, cudaMemcpyHostToDevice));
  BWTM_CUDA(cudaMemset(d_rle.as<uint8_t>() + rle_bytes, 0, RLE_PADDING));
  bwtm_index* index = nullptr;
  BWTM_TRY(index_from_device_rle(d_rle.as<uint8_t>(), rle_bytes, 0, &index));
  d_rle.detach();
  if(expected_counts != nullptr)
  {
    for(int c = 0; c < SIGMA; c++)
    {
      if(expected_counts[c] != index->counts[c])
      {
        set_error("count of comp %d is %llu, expected %llu", c,
                  (unsigned long long)index->counts[c], (unsigned long long)expected_counts[c]);
        index_free(index);
        return BWTM_ERR_ARGUMENT;
      }
    }
  }
  *out = index;
  return BWTM_OK;
}

static int check_counts(bwtm_index* index, const uint64_t* expected_counts)
{
  for(int c = 0; expected_counts != nullptr && c < SIGMA; c++)
  {
    if(expected_counts[c] != index->counts[c])
    {
      set_error("count of comp %d is %llu, expected %llu", c,
                (unsigned long long)index->counts[c], (unsigned long long)expected_counts[c]);
      return BWTM_ERR_ARGUMENT;
    }
  }
  return BWTM_OK;
}

int bwtm_index_create_pair(const uint8_t* rle_a, uint64_t rle_bytes_a, const uint64_t* expected_counts_a,
                           const uint8_t* rle_b, uint64_t rle_bytes_b, const uint64_t* expected_counts_b,
                           bwtm_index** out_a, bwtm_index** out_b)
{
  if(rle_a == nullptr || rle_b == nullptr || out_a == nullptr || out_b == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  *out_a = nullptr; *out_b = nullptr;
  BWTM_TRY(check_device());
  DeviceBuffer d_a, d_b;
  BWTM_TRY(d_a.allocate(rle_bytes_a + RLE_PADDING, true)); BWTM_TRY(d_b.allocate(rle_bytes_b + RLE_PADDING, true));
  BWTM_CUDA(cudaMemsetAsync(d_a.as<uint8_t>() + rle_bytes_a, 0, RLE_PADDING, 0));
  BWTM_CUDA(cudaMemsetAsync(d_b.as<uint8_t>() + rle_bytes_b, 0, RLE_PADDING, 0));
  BWTM_CUDA(cudaStreamSynchronize(0));   // the buffers come from stream 0's pool: they are usable on the copy stream from here on

  // Both uploads go to a copy stream back to back; K0 of the first input starts as soon as its bytes are
  // there and overlaps the upload of the second one.
  cudaStream_t copy = nullptr; cudaEvent_t arrived_a = nullptr, arrived_b = nullptr;
  int rc = BWTM_OK;
  bwtm_index *index_a = nullptr, *index_b = nullptr;
  auto cleanup = [&]()
  {
    if(copy != nullptr) { cudaStreamSynchronize(copy); cudaStreamDestroy(copy); }
    if(arrived_a != nullptr) { cudaEventDestroy(arrived_a); }
    if(arrived_b != nullptr) { cudaEventDestroy(arrived_b); }
  };
  auto failed = [&](const char* what) { set_error("%s: %s", what, cudaGetErrorString(cudaGetLastError())); cleanup(); return BWTM_ERR_CUDA; };
  if(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking) != cudaSuccess) { return failed("cannot create the copy stream"); }
  if(cudaEventCreateWithFlags(&arrived_a, cudaEventDisableTiming) != cudaSuccess ||
     cudaEventCreateWithFlags(&arrived_b, cudaEventDisableTiming) != cudaSuccess) { return failed("cannot create events"); }
  if(cudaMemcpyAsync(d_a.ptr, rle_a, rle_bytes_a, cudaMemcpyHostToDevice, copy) != cudaSuccess || cudaEventRecord(arrived_a, copy) != cudaSuccess ||
     cudaMemcpyAsync(d_b.ptr, rle_b, rle_bytes_b, cudaMemcpyHostToDevice, copy) != cudaSuccess || cudaEventRecord(arrived_b, copy) != cudaSuccess)
  {
    return failed("cannot upload the run-length bytes");
  }
  if(cudaStreamWaitEvent(0, arrived_a, 0) != cudaSuccess) { return failed("cannot order the streams"); }
  rc = index_from_device_rle(d_a.as<uint8_t>(), rle_bytes_a, 0, &index_a);
  if(rc == BWTM_OK) { d_a.detach(); rc = check_counts(index_a, expected_counts_a); }
  if(rc == BWTM_OK)
  {
    // The second input is still on its way: the pair records of the first one (which a merge of the two would build
    // first thing) are made meanwhile. The second input's length is not known yet: its bytes scale it.
    uint64_t expected_b = 0;
    if(expected_counts_b != nullptr) { for(int c = 0; c < SIGMA; c++) { expected_b += expected_counts_b[c]; } }
    else { expected_b = (uint64_t)((double)index_a->size * ((double)rle_bytes_b / (double)rle_bytes_a)); }
    build_pairs_ahead(index_a, expected_b, 0);
  }
  if(rc == BWTM_OK && cudaStreamWaitEvent(0, arrived_b, 0) != cudaSuccess) { set_error("cannot order the streams"); rc = BWTM_ERR_CUDA; }
  if(rc == BWTM_OK) { rc = index_from_device_rle(d_b.as<uint8_t>(), rle_bytes_b, 0, &index_b); }
  if(rc == BWTM_OK) { d_b.detach(); rc = check_counts(index_b, expected_counts_b); }
  cleanup();
  if(rc != BWTM_OK) { index_free(index_a); index_free(index_b); return rc; }
  *out_a = index_a; *out_b = index_b;
  return BWTM_OK;
}

// Bytes of the library's memory pool in use now and at their highest since the last reset.
int bwtm_memory_stats(uint64_t* used_bytes, uint64_t* peak_bytes, int reset_peak)
{
  int device = 0; cudaMemPool_t pool;
  BWTM_CUDA(cudaGetDevice(&device));
  BWTM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t used = 0, peak = 0;
  BWTM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
  BWTM_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &peak));
  if(reset_peak) { uint64_t zero = 0; BWTM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &zero)); }
  cudaMemPool_t resident;
  BWTM_TRY(resident_pool(&resident));   // the indexes live in a pool of their own: the sum of the two peaks bounds the real one
  uint64_t resident_used = 0, resident_peak = 0;
  BWTM_CUDA(cudaMemPoolGetAttribute(resident, cudaMemPoolAttrUsedMemCurrent, &resident_used));
  BWTM_CUDA(cudaMemPoolGetAttribute(resident, cudaMemPoolAttrUsedMemHigh, &resident_peak));
  if(reset_peak) { uint64_t zero = 0; BWTM_CUDA(cudaMemPoolSetAttribute(resident, cudaMemPoolAttrUsedMemHigh, &zero)); }
  uint64_t plain = 0;
  { std::lock_guard<std::mutex> lock(plain_allocations_mutex); plain = plain_bytes; }   // pair records built ahead of time
  if(used_bytes != nullptr) { *used_bytes = used + resident_used + plain; }
  if(peak_bytes != nullptr) { *peak_bytes = peak + resident_peak + plain; }
  return BWTM_OK;
}

int bwtm_index_destroy(bwtm_index* index)
{
  index_free(index);
  return BWTM_OK;
}

int bwtm_index_get_info(const bwtm_index* index, bwtm_index_info* info)
{
  if(index == nullptr || info == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  info->sequences = index->sequences; info->bases = index->size; info->rle_bytes = index->rle_bytes;
  for(int c = 0; c < SIGMA; c++) { info->counts[c] = index->counts[c]; }
  for(int c = 0; c <= SIGMA; c++) { info->C[c] = index->C[c]; }
  info->device_bytes = index->device_bytes;
  return BWTM_OK;
}

int bwtm_index_download(const bwtm_index* index, uint8_t* out_rle, uint64_t capacity, uint64_t* rle_bytes)
{
  if(index == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  if(rle_bytes != nullptr) { *rle_bytes = index->rle_bytes; }
  if(out_rle == nullptr) { return BWTM_OK; }
  if(capacity < index->rle_bytes) { set_error("output buffer too small"); return BWTM_ERR_CAPACITY; }
  BWTM_CUDA(cudaMemcpy(out_rle, index->d_rle, index->rle_bytes, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

// Every query reads the rank records: an index built with skip_index has none (RLE bytes only).
static int require_records(const bwtm_index* index)
{
  if(index->d_records != nullptr) { return BWTM_OK; }
  set_error("the index has no rank structure (it was built with skip_index): only info and download are available");
  return BWTM_ERR_ARGUMENT;
}

int bwtm_index_samples(const bwtm_index* index, uint64_t* block_ends, uint64_t* cumulative, uint64_t blocks)
{
  if(index == nullptr || block_ends == nullptr || cumulative == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  uint64_t expected = div_up(index->rle_bytes, RLE_BLOCK);
  if(blocks != expected) { set_error("the index has %llu blocks", (unsigned long long)expected); return BWTM_ERR_ARGUMENT; }
  DeviceBuffer starts; BWTM_TRY(starts.allocate((blocks + 1) * sizeof(uint64_t)));
  BWTM_TRY(rle_block_starts(index->d_rle, index->rle_bytes, starts.as<uint64_t>(), 0));
  DeviceBuffer ends; BWTM_TRY(ends.allocate(blocks * sizeof(uint64_t)));
  DeviceBuffer cum; BWTM_TRY(cum.allocate(SIGMA * blocks * sizeof(uint64_t)));
  sample_blocks<<<(unsigned)div_up(blocks, 256), 256>>>(device_view(index), starts.as<uint64_t>(), blocks,
                                                         ends.as<uint64_t>(), cum.as<uint64_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(block_ends, ends.ptr, blocks * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  BWTM_CUDA(cudaMemcpy(cumulative, cum.ptr, SIGMA * blocks * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

int bwtm_index_extract(const bwtm_index* index, uint64_t first, uint64_t count, uint8_t* out_comps)
{
  if(index == nullptr || out_comps == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  if(first > index->size || count > index->size - first) { set_error("range out of bounds"); return BWTM_ERR_ARGUMENT; }
  if(count == 0) { return BWTM_OK; }
  DeviceBuffer out; BWTM_TRY(out.allocate(count));
  extract_symbols<<<(unsigned)div_up(count, 256), 256>>>(device_view(index), first, count, out.as<uint8_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(out_comps, out.ptr, count, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

int bwtm_index_hash(const bwtm_index* index, uint64_t* hash)
{
  if(index == nullptr || hash == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  // FNV-1a is inherently sequential; the symbols are extracted on the device and folded on the host.
  const uint64_t CHUNK = 64ull << 20;
  std::vector<uint8_t> buffer(index->size < CHUNK ? index->size : CHUNK);
  uint64_t h = 0xcbf29ce484222325ULL;
  for(uint64_t first = 0; first < index->size; first += CHUNK)
  {
    uint64_t count = (index->size - first < CHUNK ? index->size - first : CHUNK);
    BWTM_TRY(bwtm_index_extract(index, first, count, buffer.data()));
    for(uint64_t i = 0; i < count; i++) { h = (h ^ buffer[i]) * 0x100000001b3ULL; }
  }
  *hash = h;
  return BWTM_OK;
}

int bwtm_rank(const bwtm_index* index, const uint64_t* positions, const uint8_t* comps, uint64_t n, uint64_t* out_ranks)
{
  if(index == nullptr || positions == nullptr || comps == nullptr || out_ranks == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  if(n == 0) { return BWTM_OK; }
  DeviceBuffer pos, cmp, res;
  BWTM_TRY(pos.allocate(n * 8)); BWTM_TRY(cmp.allocate(n)); BWTM_TRY(res.allocate(n * 8));
  BWTM_CUDA(cudaMemcpy(pos.ptr, positions, n * 8, cudaMemcpyHostToDevice));
  BWTM_CUDA(cudaMemcpy(cmp.ptr, comps, n, cudaMemcpyHostToDevice));
  query_rank<<<(unsigned)div_up(n, 256), 256>>>(device_view(index), pos.as<uint64_t>(), cmp.as<uint8_t>(), n, res.as<uint64_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(out_ranks, res.ptr, n * 8, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

int bwtm_lf(const bwtm_index* index, const uint64_t* positions, uint64_t n, uint64_t* out_positions, uint8_t* out_comps)
{
  if(index == nullptr || positions == nullptr || out_positions == nullptr || out_comps == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  if(n == 0) { return BWTM_OK; }
  DeviceBuffer pos, res, cmp;
  BWTM_TRY(pos.allocate(n * 8)); BWTM_TRY(res.allocate(n * 8)); BWTM_TRY(cmp.allocate(n));
  BWTM_CUDA(cudaMemcpy(pos.ptr, positions, n * 8, cudaMemcpyHostToDevice));
  query_lf<<<(unsigned)div_up(n, 256), 256>>>(device_view(index), pos.as<uint64_t>(), n, res.as<uint64_t>(), cmp.as<uint8_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(out_positions, res.ptr, n * 8, cudaMemcpyDeviceToHost));
  BWTM_CUDA(cudaMemcpy(out_comps, cmp.ptr, n, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

int bwtm_count(const bwtm_index* index, const uint8_t* patterns, const uint64_t* offsets, uint64_t n,
               const uint8_t* char2comp, uint64_t* out_counts)
{
  if(index == nullptr || offsets == nullptr || out_counts == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  BWTM_TRY(require_records(index));
  if(n == 0) { return BWTM_OK; }
  uint64_t total = offsets[n];
  if(total > 0 && patterns == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  DeviceBuffer pat, off, map, res;
  BWTM_TRY(pat.allocate(total)); BWTM_TRY(off.allocate((n + 1) * 8)); BWTM_TRY(res.allocate(n * 8));
  if(total > 0) { BWTM_CUDA(cudaMemcpy(pat.ptr, patterns, total, cudaMemcpyHostToDevice)); }
  BWTM_CUDA(cudaMemcpy(off.ptr, offsets, (n + 1) * 8, cudaMemcpyHostToDevice));
  if(char2comp != nullptr)
  {
    BWTM_TRY(map.allocate(256));
    BWTM_CUDA(cudaMemcpy(map.ptr, char2comp, 256, cudaMemcpyHostToDevice));
  }
  query_count<<<(unsigned)div_up(n, 128), 128>>>(device_view(index), pat.as<uint8_t>(), off.as<uint64_t>(), n,
                                                  char2comp != nullptr ? map.as<uint8_t>() : nullptr, res.as<uint64_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaMemcpy(out_counts, res.ptr, n * 8, cudaMemcpyDeviceToHost));
  return BWTM_OK;
}

} // extern "C"
