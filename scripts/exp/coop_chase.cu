// Experiment (not product): which shape of dependent random record reads gives the most records/s on B200?
// LANES lanes cooperate on one GRANULE-byte record (one instruction covers the record), CHAINS independent
// walkers per lane group, optional ld.global.nc.L2::64B hint.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template<int HINT> __device__ __forceinline__ uint4 load16(const uint4* p)
{
  uint4 r;
  if(HINT == 0) { r = __ldg(p); }
  else { asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); }
  return r;
}

template<int GRANULE, int LANES, int CHAINS, int HINT>
__global__ void __launch_bounds__(256) chase(const uint4* __restrict__ table, uint64_t granules, uint64_t steps, uint64_t seed, uint32_t* sink)
{
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t group = tid / LANES; int sub = (int)(tid % LANES);
  constexpr int VEC = GRANULE / 16;          // uint4 per record
  constexpr int PER_LANE = VEC / LANES;      // uint4 per lane
  uint64_t state[CHAINS];
#pragma unroll
  for(int c = 0; c < CHAINS; c++) { state[c] = mix64(seed + group * CHAINS + c); }
  uint32_t acc = 0;
  for(uint64_t k = 0; k < steps; k++)
  {
    uint4 q[CHAINS][PER_LANE];
#pragma unroll
    for(int c = 0; c < CHAINS; c++)
    {
      uint64_t g = __umul64hi(state[c], granules);
      const uint4* p = table + g * VEC + sub * PER_LANE;
#pragma unroll
      for(int v = 0; v < PER_LANE; v++) { q[c][v] = load16<HINT>(p + v); }
    }
#pragma unroll
    for(int c = 0; c < CHAINS; c++)
    {
      uint32_t x = 0;
#pragma unroll
      for(int v = 0; v < PER_LANE; v++) { x ^= q[c][v].x ^ q[c][v].y ^ q[c][v].z ^ q[c][v].w; }
#pragma unroll
      for(int o = 1; o < LANES; o <<= 1) { x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o); }
      acc ^= x;
      state[c] = state[c] * 6364136223846793005ull + 1442695040888963407ull + x;
    }
  }
  if(acc == 0x12345678u) { sink[0] = acc; }
}

template<int GRANULE, int LANES, int CHAINS, int HINT>
void run(const uint4* table, uint64_t bytes, uint32_t* sink, int threads_per_sm)
{
  uint64_t threads = 148ull * threads_per_sm;
  uint64_t walkers = threads * CHAINS / LANES;
  uint64_t steps = (64ull << 20) / walkers; if(steps < 16) steps = 16;
  cudaEvent_t b, e; cudaEventCreate(&b); cudaEventCreate(&e);
  float best = 1e30f;
  for(int it = 0; it < 3; it++)
  {
    cudaEventRecord(b);
    chase<GRANULE, LANES, CHAINS, HINT><<<(unsigned)(threads / 256), 256>>>(table, bytes / GRANULE, steps, 7 + it, sink);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, b, e); if(ms < best) best = ms;
  }
  printf("granule %3d lanes %d chains %d hint %d threads/SM %4d walkers/SM %5d : %7.2f G records/s  (%s)\n", GRANULE, LANES, CHAINS, HINT,
         threads_per_sm, (int)(walkers / 148), walkers * steps / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
  uint64_t gb = (argc > 1 ? atoi(argv[1]) : 2);
  cudaSetDevice(0);
  uint64_t bytes = gb << 30;
  uint4* table; uint32_t* sink;
  cudaMalloc(&table, bytes); cudaMalloc(&sink, 16); cudaMemset(table, 1, bytes);
  printf("table %llu GB\n", (unsigned long long)gb);
  for(int tps : {1024, 2048})
  {
    run<64, 1, 1, 0>(table, bytes, sink, tps);
    run<64, 1, 1, 1>(table, bytes, sink, tps);
    run<64, 1, 2, 1>(table, bytes, sink, tps);
    run<64, 2, 1, 0>(table, bytes, sink, tps);
    run<64, 2, 1, 1>(table, bytes, sink, tps);
    run<64, 2, 2, 1>(table, bytes, sink, tps);
    run<64, 4, 1, 0>(table, bytes, sink, tps);
    run<64, 4, 1, 1>(table, bytes, sink, tps);
    run<64, 4, 2, 1>(table, bytes, sink, tps);
    run<64, 4, 4, 1>(table, bytes, sink, tps);
    run<32, 1, 1, 1>(table, bytes, sink, tps);
    run<32, 2, 1, 1>(table, bytes, sink, tps);
    run<32, 2, 2, 1>(table, bytes, sink, tps);
    run<128, 4, 1, 0>(table, bytes, sink, tps);
    run<128, 8, 1, 0>(table, bytes, sink, tps);
    run<128, 8, 2, 0>(table, bytes, sink, tps);
  }
  return 0;
}
