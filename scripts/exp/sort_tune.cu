// Experiment (not product): cub::DeviceRadixSort onesweep tuning for 1.5 G uniform 31-bit keys on B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>

using namespace cub;

template<int THREADS, int ITEMS, bool USE_ONESWEEP, int RANK = RADIX_RANK_MATCH_EARLY_COUNTS_ANY, int PRIMARY = 7, int DTHREADS = 512, int DITEMS = 23>
struct hub
{
  using KeyT = uint32_t; using ValueT = NullType; using OffsetT = unsigned long long;
  static constexpr bool KEYS_ONLY = true;
  using DominantT = KeyT;
  struct Policy : ChainedPolicy<1000, Policy, Policy>
  {
    static constexpr bool ONESWEEP = USE_ONESWEEP;
    static constexpr int ONESWEEP_RADIX_BITS = 8;
    using HistogramPolicy    = AgentRadixSortHistogramPolicy<128, 16, 1, KeyT, ONESWEEP_RADIX_BITS>;
    using ExclusiveSumPolicy = AgentRadixSortExclusiveSumPolicy<256, ONESWEEP_RADIX_BITS>;
    static constexpr int PRIMARY_RADIX_BITS = PRIMARY, SINGLE_TILE_RADIX_BITS = 6, SEGMENTED_RADIX_BITS = 6;
    using OnesweepPolicy = AgentRadixSortOnesweepPolicy<THREADS, ITEMS, DominantT, 1, (RadixRankAlgorithm)RANK, BLOCK_SCAN_RAKING_MEMOIZE, RADIX_SORT_STORE_DIRECT, ONESWEEP_RADIX_BITS>;
    using ScanPolicy = AgentScanPolicy<512, 23, OffsetT, BLOCK_LOAD_WARP_TRANSPOSE, LOAD_DEFAULT, BLOCK_STORE_WARP_TRANSPOSE, BLOCK_SCAN_RAKING_MEMOIZE>;
    using DownsweepPolicy = AgentRadixSortDownsweepPolicy<DTHREADS, DITEMS, DominantT, BLOCK_LOAD_TRANSPOSE, LOAD_DEFAULT, RADIX_RANK_MATCH, BLOCK_SCAN_WARP_SCANS, PRIMARY_RADIX_BITS>;
    using AltDownsweepPolicy = AgentRadixSortDownsweepPolicy<256, 47, DominantT, BLOCK_LOAD_TRANSPOSE, LOAD_DEFAULT, RADIX_RANK_MEMOIZE, BLOCK_SCAN_WARP_SCANS, PRIMARY_RADIX_BITS - 1>;
    using UpsweepPolicy    = AgentRadixSortUpsweepPolicy<256, 23, DominantT, LOAD_DEFAULT, PRIMARY_RADIX_BITS>;
    using AltUpsweepPolicy = AgentRadixSortUpsweepPolicy<256, 47, DominantT, LOAD_DEFAULT, PRIMARY_RADIX_BITS - 1>;
    using SingleTilePolicy = AgentRadixSortDownsweepPolicy<256, 19, DominantT, BLOCK_LOAD_DIRECT, LOAD_LDG, RADIX_RANK_MEMOIZE, BLOCK_SCAN_WARP_SCANS, SINGLE_TILE_RADIX_BITS>;
    using SegmentedPolicy = AgentRadixSortDownsweepPolicy<192, 39, DominantT, BLOCK_LOAD_TRANSPOSE, LOAD_DEFAULT, RADIX_RANK_MEMOIZE, BLOCK_SCAN_WARP_SCANS, SEGMENTED_RADIX_BITS>;
    using AltSegmentedPolicy = AgentRadixSortDownsweepPolicy<384, 11, DominantT, BLOCK_LOAD_TRANSPOSE, LOAD_DEFAULT, RADIX_RANK_MEMOIZE, BLOCK_SCAN_WARP_SCANS, SEGMENTED_RADIX_BITS - 1>;
  };
  using MaxPolicy = Policy;
};

__global__ void fill(uint32_t* keys, uint64_t n, uint32_t range)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  uint64_t z = i + 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  keys[i] = (uint32_t)(z % range);
}

__global__ void check(const uint32_t* keys, uint64_t n, int* bad)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i + 1 < n && keys[i] > keys[i + 1]) { *bad = 1; }
}

template<class Hub>
void run(const char* name, uint32_t* keys, uint32_t* alt, uint64_t n, int bits, int* d_bad)
{
  using Dispatch = DispatchRadixSort<false, uint32_t, NullType, unsigned long long, Hub>;
  float best = 1e30f;
  for(int it = 0; it < 3; it++)
  {
    fill<<<(unsigned)((n + 255) / 256), 256>>>(keys, n, 1510000001u);
    DoubleBuffer<uint32_t> kb(keys, alt); DoubleBuffer<NullType> vb;
    size_t bytes = 0;
    Dispatch::Dispatch(nullptr, bytes, kb, vb, (unsigned long long)n, 0, bits, true, 0);
    void* temp; cudaMalloc(&temp, bytes);
    cudaEvent_t b, e; cudaEventCreate(&b); cudaEventCreate(&e);
    cudaEventRecord(b);
    cudaError_t err = Dispatch::Dispatch(temp, bytes, kb, vb, (unsigned long long)n, 0, bits, true, 0);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, b, e); if(ms < best) best = ms;
    if(it == 0)
    {
      cudaMemset(d_bad, 0, 4);
      check<<<(unsigned)((n + 255) / 256), 256>>>(kb.Current(), n, d_bad);
      int bad = 0; cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
      if(bad || err != cudaSuccess) printf("  !! %s: sorted=%d err=%s\n", name, !bad, cudaGetErrorString(err));
    }
    cudaFree(temp);
  }
  printf("%-34s %8.2f ms  %6.1f Gkeys/s (%s)\n", name, best, n / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  uint64_t n = 1510000000ull; int bits = 31;
  uint32_t *keys, *alt; int* d_bad;
  cudaMalloc(&keys, n * 4); cudaMalloc(&alt, n * 4); cudaMalloc(&d_bad, 4);
  {
    // library default for reference
    float best = 1e30f;
    for(int it = 0; it < 3; it++)
    {
      fill<<<(unsigned)((n + 255) / 256), 256>>>(keys, n, 1510000001u);
      DoubleBuffer<uint32_t> kb(keys, alt); size_t bytes = 0;
      DeviceRadixSort::SortKeys(nullptr, bytes, kb, (long long)n, 0, bits);
      void* temp; cudaMalloc(&temp, bytes);
      cudaEvent_t b, e; cudaEventCreate(&b); cudaEventCreate(&e);
      cudaEventRecord(b); DeviceRadixSort::SortKeys(temp, bytes, kb, (long long)n, 0, bits); cudaEventRecord(e); cudaEventSynchronize(e);
      float ms; cudaEventElapsedTime(&ms, b, e); if(ms < best) best = ms; cudaFree(temp);
    }
    printf("%-34s %8.2f ms\n", "DeviceRadixSort::SortKeys default", best);
  }
  run<hub<384, 19, true>>("onesweep 384x19 (default)", keys, alt, n, bits, d_bad);
  run<hub<256, 30, true>>("onesweep 256x30", keys, alt, n, bits, d_bad);
  run<hub<384, 24, true>>("onesweep 384x24", keys, alt, n, bits, d_bad);
  run<hub<512, 19, true>>("onesweep 512x19", keys, alt, n, bits, d_bad);
  run<hub<128, 40, true>>("onesweep 128x40", keys, alt, n, bits, d_bad);
  run<hub<384, 19, false, RADIX_RANK_MATCH_EARLY_COUNTS_ANY, 7>>("legacy 7 bits 512x23", keys, alt, n, bits, d_bad);
  return 0;
}
