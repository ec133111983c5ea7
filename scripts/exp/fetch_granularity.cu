// Experiment (not product): how much DRAM traffic does a random 32/64-byte record read cost on B200,
// for different load flavours and cudaLimitMaxL2FetchGranularity settings?  Run under ncu with
//   --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum
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

template<int V> __device__ __forceinline__ uint4 load16(const uint4* p);
template<> __device__ __forceinline__ uint4 load16<0>(const uint4* p) { return __ldg(p); }
template<> __device__ __forceinline__ uint4 load16<1>(const uint4* p) { uint4 r; asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template<> __device__ __forceinline__ uint4 load16<2>(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template<> __device__ __forceinline__ uint4 load16<3>(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template<> __device__ __forceinline__ uint4 load16<4>(const uint4* p) { uint4 r; asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template<> __device__ __forceinline__ uint4 load16<5>(const uint4* p) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
template<> __device__ __forceinline__ uint4 load16<6>(const uint4* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

// Dependent chain of random GRANULE-byte reads, load flavour V.
template<int GRANULE, int V>
__global__ void __launch_bounds__(256) chase(const uint4* __restrict__ table, uint64_t granules, uint64_t steps, uint64_t seed, uint32_t* sink)
{
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t state = mix64(seed + tid);
  uint32_t acc = 0;
  constexpr int VEC = GRANULE / 16;
  for(uint64_t k = 0; k < steps; k++)
  {
    uint64_t g = __umul64hi(state, granules);
    const uint4* p = table + g * VEC;
    uint32_t x = 0;
#pragma unroll
    for(int v = 0; v < VEC; v++) { uint4 q = load16<V>(p + v); x ^= q.x ^ q.y ^ q.z ^ q.w; }
    acc ^= x;
    state = state * 6364136223846793005ull + 1442695040888963407ull + x;
  }
  if(acc == 0x12345678u) { sink[0] = acc; }
}

// 256-bit loads (sm_100): one instruction per 32-byte sector.
template<int GRANULE>
__global__ void __launch_bounds__(256) chase256(const uint4* __restrict__ table, uint64_t granules, uint64_t steps, uint64_t seed, uint32_t* sink)
{
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t state = mix64(seed + tid);
  uint32_t acc = 0;
  constexpr int VEC = GRANULE / 32;
  for(uint64_t k = 0; k < steps; k++)
  {
    uint64_t g = __umul64hi(state, granules);
    const char* p = reinterpret_cast<const char*>(table) + g * GRANULE;
    uint32_t x = 0;
#pragma unroll
    for(int v = 0; v < VEC; v++)
    {
      uint32_t r[8];
      asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + 32 * v));
      x ^= r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
    }
    acc ^= x;
    state = state * 6364136223846793005ull + 1442695040888963407ull + x;
  }
  if(acc == 0x12345678u) { sink[0] = acc; }
}

template<int GRANULE, int V>
void run(const char* name, const uint4* table, uint64_t bytes, uint32_t* sink)
{
  uint64_t threads = 148ull * 1024, steps = 512;
  cudaEvent_t b, e; cudaEventCreate(&b); cudaEventCreate(&e);
  cudaEventRecord(b);
  chase<GRANULE, V><<<(unsigned)(threads / 256), 256>>>(table, bytes / GRANULE, steps, 7, sink);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, b, e);
  printf("%-28s granule %3d: %8.3f ms  %7.2f G loads/s  %s\n", name, GRANULE, ms, threads * steps / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

template<int GRANULE>
void run256(const uint4* table, uint64_t bytes, uint32_t* sink)
{
  uint64_t threads = 148ull * 1024, steps = 512;
  cudaEvent_t b, e; cudaEventCreate(&b); cudaEventCreate(&e);
  cudaEventRecord(b);
  chase256<GRANULE><<<(unsigned)(threads / 256), 256>>>(table, bytes / GRANULE, steps, 7, sink);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, b, e);
  printf("%-28s granule %3d: %8.3f ms  %7.2f G loads/s  %s\n", "ld.nc.v8.u32 (256-bit)", GRANULE, ms, threads * steps / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
  size_t limit = (argc > 1 ? atoi(argv[1]) : 0);
  cudaSetDevice(0);
  size_t before = 0, after = 0;
  if(limit) { printf("set limit: %s\n", cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, limit))); }
  cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
  printf("cudaLimitMaxL2FetchGranularity = %zu (requested %zu)\n", after, limit); (void)before;
  uint64_t bytes = 2ull << 30;
  uint4* table; uint32_t* sink;
  cudaMalloc(&table, bytes); cudaMalloc(&sink, 16); cudaMemset(table, 1, bytes);
  run<32, 0>("__ldg (ld.global.nc)", table, bytes, sink);
  run<64, 0>("__ldg (ld.global.nc)", table, bytes, sink);
  run<32, 1>("ld.global", table, bytes, sink);
  run<64, 1>("ld.global", table, bytes, sink);
  run<32, 2>("ld.global.nc.L2::64B", table, bytes, sink);
  run<64, 2>("ld.global.nc.L2::64B", table, bytes, sink);
  run<64, 3>("ld.global.nc.L2::128B", table, bytes, sink);
  run<64, 4>("ld.global.L1::no_allocate", table, bytes, sink);
  run<64, 5>("ld.global.cg", table, bytes, sink);
  run<64, 6>("ld.global.nc.L1::no_allocate", table, bytes, sink);
  run256<32>(table, bytes, sink);
  run256<64>(table, bytes, sink);
  cudaDeviceSynchronize();
  return 0;
}
