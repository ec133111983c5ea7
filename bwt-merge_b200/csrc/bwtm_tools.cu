// Fixture tools: synthetic reads, a sort-based builder of the multi-string BWT, and the HBM
// random-access microbenchmark.  None of this exists in the reference (it only merges; its inputs
// came from other tools, README.md:21).  The builder makes inputs of benchmark size on the device and
// gives an independent answer for BWT(A ++ B) to compare merges with at sizes no CPU oracle reaches.
//
// Multi-string BWT (paper.tex:141-145): read i ends with its own endmarker $_i, $_i < $_j for i < j,
// endmarkers sort before bases; row i < reads is the suffix "$_i".  All suffixes of all reads are
// sorted by LSD radix passes over 63-bit words of 21 symbols (3 bits each, endmarker and padding
// = 0); the sort is stable and starts in (read, offset) order, which breaks ties between equal
// suffixes of different reads by read index.
#include <algorithm>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "bwtm_merge.cuh"

namespace bwtm
{

//------------------------------------------------------------------------------
// Counter-based generator: twin of bwtm_b200/synth.py

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t stream_base(uint64_t seed, uint64_t stream)
{
  return mix64(seed + stream * 0xD1342543DE82EF95ull);
}

__global__ void gen_genome(uint8_t* __restrict__ genome, uint64_t length, uint64_t base)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < length) { genome[i] = (uint8_t)((mix64(base + i) & 3u) + 1u); }
}

// One thread per base of the segment's reads.
__global__ void gen_reads(const uint8_t* __restrict__ genome, uint64_t genome_len, uint64_t reads, uint64_t read_len,
                          uint64_t first_read, uint64_t threshold, uint64_t base_start, uint64_t base_sub, uint64_t base_shift,
                          uint8_t* __restrict__ out)
{
  uint64_t local = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(local >= reads * read_len) { return; }
  uint64_t read = first_read + local / read_len, k = local % read_len;
  uint64_t cell = read * read_len + k;
  uint64_t start = mix64(base_start + read) % (genome_len - read_len + 1);
  uint32_t b = genome[start + k] - 1u;
  if((mix64(base_sub + cell) >> 11) < threshold) { b = (b + 1u + (uint32_t)(mix64(base_shift + cell) % 3u)) & 3u; }
  out[local] = (uint8_t)(b + 1u);
}

//------------------------------------------------------------------------------
// Suffix sorting

constexpr int WORD_SYMBOLS = 21;

// 63-bit key of symbols [21 w, 21 w + 21) of suffix id (read = id / (L + 1), offset = id % (L + 1)).
template<class IdT>
__global__ void suffix_keys(const uint8_t* __restrict__ reads, uint64_t read_len, const IdT* __restrict__ ids,
                            uint64_t n, int word, uint64_t* __restrict__ keys)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) { return; }
  uint64_t id = ids[i];
  uint64_t read = id / (read_len + 1), offset = id - read * (read_len + 1);
  uint64_t first = offset + (uint64_t)word * WORD_SYMBOLS;
  const uint8_t* row = reads + read * read_len;
  uint64_t key = 0;
#pragma unroll
  for(int k = 0; k < WORD_SYMBOLS; k++)
  {
    uint64_t p = first + k;
    uint64_t comp = (p < read_len ? row[p] : 0);
    key = (key << 3) | comp;
  }
  keys[i] = key;
}

template<class IdT>
__global__ void init_ids(IdT* ids, uint64_t n)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) { ids[i] = (IdT)i; }
}

template<class IdT>
__global__ void bwt_symbols(const uint8_t* __restrict__ reads, uint64_t read_len, const IdT* __restrict__ ids,
                            uint64_t n, uint8_t* __restrict__ out)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) { return; }
  uint64_t id = ids[i];
  uint64_t read = id / (read_len + 1), offset = id - read * (read_len + 1);
  out[i] = (offset > 0 ? reads[read * read_len + offset - 1] : 0);
}

template<class IdT>
static int build_from_device_reads(const uint8_t* d_reads, uint64_t reads, uint64_t read_len, bwtm_index** out)
{
  cudaStream_t stream = 0;
  uint64_t n = reads * (read_len + 1);
  int words = (int)div_up(read_len + 1, WORD_SYMBOLS);

  DeviceBuffer keys, keys_alt, ids, ids_alt, temp;
  BWTM_TRY(keys.allocate(n * sizeof(uint64_t))); BWTM_TRY(keys_alt.allocate(n * sizeof(uint64_t)));
  BWTM_TRY(ids.allocate(n * sizeof(IdT))); BWTM_TRY(ids_alt.allocate(n * sizeof(IdT)));
  cub::DoubleBuffer<uint64_t> key_buffers(keys.as<uint64_t>(), keys_alt.as<uint64_t>());
  cub::DoubleBuffer<IdT> id_buffers(ids.as<IdT>(), ids_alt.as<IdT>());
  size_t temp_bytes = 0;
  BWTM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key_buffers, id_buffers, (int64_t)n, 0, 63, stream));
  BWTM_TRY(temp.allocate(temp_bytes));

  unsigned grid = (unsigned)div_up(n, 256);
  init_ids<IdT><<<grid, 256, 0, stream>>>(id_buffers.Current(), n);
  BWTM_LAUNCH_CHECK();
  for(int w = words - 1; w >= 0; w--)
  {
    suffix_keys<IdT><<<grid, 256, 0, stream>>>(d_reads, read_len, id_buffers.Current(), n, w, key_buffers.Current());
    BWTM_LAUNCH_CHECK();
    BWTM_CUDA(cub::DeviceRadixSort::SortPairs(temp.ptr, temp_bytes, key_buffers, id_buffers, (int64_t)n, 0, 63, stream));
    count_launch(10);
  }
  keys.release(); keys_alt.release(); temp.release();

  DeviceBuffer symbols; BWTM_TRY(symbols.allocate(n));
  bwt_symbols<IdT><<<grid, 256, 0, stream>>>(d_reads, read_len, id_buffers.Current(), n, symbols.as<uint8_t>());
  BWTM_LAUNCH_CHECK();
  BWTM_CUDA(cudaStreamSynchronize(stream));
  ids.release(); ids_alt.release();
  return index_from_symbols(symbols.as<uint8_t>(), n, 0, stream, out);
}

static int build_dispatch(const uint8_t* d_reads, uint64_t reads, uint64_t read_len, bwtm_index** out)
{
  if(reads == 0 || read_len == 0) { set_error("empty read collection"); return BWTM_ERR_ARGUMENT; }
  uint64_t n = reads * (read_len + 1);
  if(n < 0xFFFFFFFFull) { return build_from_device_reads<uint32_t>(d_reads, reads, read_len, out); }
  return build_from_device_reads<uint64_t>(d_reads, reads, read_len, out);
}

//------------------------------------------------------------------------------
// Random-access microbenchmark

template<int GRANULE>
__global__ void __launch_bounds__(256)
gather_kernel(const uint4* __restrict__ table, uint64_t granules, uint64_t loads_per_thread, uint64_t seed,
              uint32_t* __restrict__ sink)
{
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t state = mix64(seed + tid);
  uint32_t acc = 0;
  constexpr int VEC = GRANULE / 16;
  constexpr int BATCH = (GRANULE == 128 ? 2 : 4);   // independent granules in flight per thread
  for(uint64_t k = 0; k < loads_per_thread; k += BATCH)
  {
    uint4 q[BATCH][VEC];
#pragma unroll
    for(int b = 0; b < BATCH; b++)
    {
      state = state * 6364136223846793005ull + 1442695040888963407ull;
      uint64_t g = __umul64hi(state, granules);     // uniform in [0, granules)
      const uint4* p = table + g * VEC;
#pragma unroll
      for(int v = 0; v < VEC; v++) { q[b][v] = __ldg(p + v); }
    }
#pragma unroll
    for(int b = 0; b < BATCH; b++)
    {
#pragma unroll
      for(int v = 0; v < VEC; v++) { acc ^= q[b][v].x ^ q[b][v].y ^ q[b][v].z ^ q[b][v].w; }
    }
  }
  if(acc == 0x12345678u) { sink[0] = acc; }
}

// Dependent random record reads in the shape of the rank/LF kernel: LANES lanes share one walker and read
// one GRANULE-byte record with a single instruction (lane j reads 16-byte piece j, or pieces 2j and 2j+1).
template<int GRANULE, int LANES>
__global__ void __launch_bounds__(256)
chase_kernel(const uint4* __restrict__ table, uint64_t granules, uint64_t steps, uint64_t seed, uint32_t* __restrict__ sink)
{
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t walker = tid / LANES; int sub = (int)(tid % LANES);
  constexpr int VEC = GRANULE / 16, PER_LANE = VEC / LANES;
  uint64_t state = mix64(seed + walker);
  uint32_t acc = 0;
  for(uint64_t k = 0; k < steps; k++)
  {
    uint64_t g = __umul64hi(state, granules);
    const uint4* p = table + g * VEC + sub * PER_LANE;
    uint32_t x = 0;
#pragma unroll
    for(int v = 0; v < PER_LANE; v++) { uint4 q = __ldg(p + v); x ^= q.x ^ q.y ^ q.z ^ q.w; }
#pragma unroll
    for(int o = 1; o < LANES; o <<= 1) { x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o); }
    acc ^= x;
    state = state * 6364136223846793005ull + 1442695040888963407ull + x;   // the next address depends on the data
  }
  if(acc == 0x12345678u) { sink[0] = acc; }
}

} // namespace bwtm

using namespace bwtm;

extern "C"
{

int bwtm_tools_build_synthetic(uint64_t genome_len, uint64_t genome_seed, uint64_t read_len, uint64_t error_threshold,
                               const bwtm_read_segment* segments, uint64_t n_segments, bwtm_index** out)
{
  if(segments == nullptr || out == nullptr || n_segments == 0) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  if(read_len == 0 || genome_len < read_len) { set_error("invalid genome or read length"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available"); return BWTM_ERR_CUDA; }
  uint64_t reads = 0;
  for(uint64_t s = 0; s < n_segments; s++) { reads += segments[s].reads; }

  DeviceBuffer genome, matrix;
  BWTM_TRY(genome.allocate(genome_len)); BWTM_TRY(matrix.allocate(reads * read_len));
  gen_genome<<<(unsigned)div_up(genome_len, 256), 256>>>(genome.as<uint8_t>(), genome_len, stream_base(genome_seed, 0));
  BWTM_LAUNCH_CHECK();
  uint64_t first = 0;
  for(uint64_t s = 0; s < n_segments; s++)
  {
    uint64_t cells = segments[s].reads * read_len;
    if(cells == 0) { continue; }
    gen_reads<<<(unsigned)div_up(cells, 256), 256>>>(genome.as<uint8_t>(), genome_len, segments[s].reads, read_len, segments[s].first_read, error_threshold,
                                                     stream_base(segments[s].seed, 1), stream_base(segments[s].seed, 2),
                                                     stream_base(segments[s].seed, 3), matrix.as<uint8_t>() + first * read_len);
    BWTM_LAUNCH_CHECK();
    first += segments[s].reads;
  }
  genome.release();
  return build_dispatch(matrix.as<uint8_t>(), reads, read_len, out);
}

int bwtm_tools_build_from_reads(const uint8_t* read_comps, uint64_t reads, uint64_t read_len, bwtm_index** out)
{
  if(read_comps == nullptr || out == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available"); return BWTM_ERR_CUDA; }
  DeviceBuffer matrix; BWTM_TRY(matrix.allocate(reads * read_len));
  BWTM_CUDA(cudaMemcpy(matrix.ptr, read_comps, reads * read_len, cudaMemcpyHostToDevice));
  return build_dispatch(matrix.as<uint8_t>(), reads, read_len, out);
}

int bwtm_tools_gather_bench(uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iterations, double* gbytes_per_second)
{
  if(gbytes_per_second == nullptr || (granule != 32 && granule != 64 && granule != 128) || table_bytes < 4096)
  {
    set_error("invalid argument"); return BWTM_ERR_ARGUMENT;
  }
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available"); return BWTM_ERR_CUDA; }
  DeviceBuffer table, sink;
  BWTM_TRY(table.allocate(table_bytes)); BWTM_TRY(sink.allocate(16));
  BWTM_CUDA(cudaMemset(table.ptr, 1, table_bytes));
  int device = 0, sms = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  BWTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  uint64_t threads = (uint64_t)sms * 2048;
  uint64_t per_thread = std::max<uint64_t>(4, (n_loads / threads) & ~3ull);
  uint64_t granules = table_bytes / granule;
  cudaEvent_t begin, end;
  BWTM_CUDA(cudaEventCreate(&begin)); BWTM_CUDA(cudaEventCreate(&end));
  double best = 0.0;
  for(int it = 0; it < iterations + 1; it++)
  {
    BWTM_CUDA(cudaEventRecord(begin));
    unsigned grid = (unsigned)(threads / 256);
    if(granule == 32)       { gather_kernel<32><<<grid, 256>>>(table.as<uint4>(), granules, per_thread, 1234 + it, sink.as<uint32_t>()); }
    else if(granule == 64)  { gather_kernel<64><<<grid, 256>>>(table.as<uint4>(), granules, per_thread, 1234 + it, sink.as<uint32_t>()); }
    else                    { gather_kernel<128><<<grid, 256>>>(table.as<uint4>(), granules, per_thread, 1234 + it, sink.as<uint32_t>()); }
    BWTM_LAUNCH_CHECK();
    BWTM_CUDA(cudaEventRecord(end));
    BWTM_CUDA(cudaEventSynchronize(end));
    float ms = 0.0f; BWTM_CUDA(cudaEventElapsedTime(&ms, begin, end));
    double gbs = (double)(threads * per_thread) * granule / (ms * 1e-3) / 1e9;
    if(it > 0 && gbs > best) { best = gbs; }   // iteration 0 is the warm-up
  }
  cudaEventDestroy(begin); cudaEventDestroy(end);
  *gbytes_per_second = best;
  return BWTM_OK;
}

int bwtm_tools_chase_bench(uint64_t table_bytes, uint32_t granule, uint64_t n_loads, uint32_t threads_per_sm,
                           uint32_t l2_fetch_granularity, double* gbytes_per_second)
{
  if(gbytes_per_second == nullptr || (granule != 32 && granule != 64 && granule != 128) || table_bytes < 4096 ||
     threads_per_sm == 0 || threads_per_sm % 256 != 0 || threads_per_sm > 2048)
  {
    set_error("invalid argument"); return BWTM_ERR_ARGUMENT;
  }
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); set_error("no CUDA device available"); return BWTM_ERR_CUDA; }
  if(l2_fetch_granularity != 0) { BWTM_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, l2_fetch_granularity)); }
  DeviceBuffer table, sink;
  BWTM_TRY(table.allocate(table_bytes)); BWTM_TRY(sink.allocate(16));
  BWTM_CUDA(cudaMemset(table.ptr, 1, table_bytes));
  int device = 0, sms = 0;
  BWTM_CUDA(cudaGetDevice(&device));
  BWTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  uint64_t threads = (uint64_t)sms * threads_per_sm;
  const uint64_t lanes = (granule == 32 ? 2 : 4);
  uint64_t walkers = threads / lanes;
  uint64_t steps = std::max<uint64_t>(1, n_loads / walkers);
  uint64_t granules = table_bytes / granule;
  unsigned grid = (unsigned)(threads / 256);
  cudaEvent_t begin, end;
  BWTM_CUDA(cudaEventCreate(&begin)); BWTM_CUDA(cudaEventCreate(&end));
  double best = 0.0;
  for(int it = 0; it < 3; it++)
  {
    BWTM_CUDA(cudaEventRecord(begin));
    if(granule == 32)       { chase_kernel<32, 2><<<grid, 256>>>(table.as<uint4>(), granules, steps, 99 + it, sink.as<uint32_t>()); }
    else if(granule == 64)  { chase_kernel<64, 4><<<grid, 256>>>(table.as<uint4>(), granules, steps, 99 + it, sink.as<uint32_t>()); }
    else                    { chase_kernel<128, 4><<<grid, 256>>>(table.as<uint4>(), granules, steps, 99 + it, sink.as<uint32_t>()); }
    BWTM_LAUNCH_CHECK();
    BWTM_CUDA(cudaEventRecord(end));
    BWTM_CUDA(cudaEventSynchronize(end));
    float ms = 0.0f; BWTM_CUDA(cudaEventElapsedTime(&ms, begin, end));
    double gbs = (double)(walkers * steps) * granule / (ms * 1e-3) / 1e9;
    if(it > 0 && gbs > best) { best = gbs; }
  }
  cudaEventDestroy(begin); cudaEventDestroy(end);
  *gbytes_per_second = best;
  return BWTM_OK;
}

} // extern "C"
