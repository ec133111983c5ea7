// Shared declarations of the B200 rank-array path: error handling, the device index layout and
// the rank / LF device functions every kernel uses.
//
// Device layout of one BWT (replaces BWT::data + samples[6] + block_boundaries, bwt.h:172-178):
//
//   rle      : the reference's run-length bytes, unchanged (needed for download and as the
//              interleave's definition of the sequence).
//   records  : one 64-byte record per 128 sequence positions, 64-byte aligned.  Record r covers
//              positions [128 r, 128 r + 128).  It is four 16-byte chunks; chunk j covers 32
//              positions and holds {p0, p1, p2, h_j}: bit t of p_k is bit k of the comp value at
//              position 128 r + 32 j + t.  The four h words form a 128-bit little-endian field
//              with five 25-bit counters: counter c-1 (c = 1..5) = number of c's in
//              [start of the superblock, 128 r).
//   super    : one row of 8 u64 per superblock of 2^25 positions: absolute count of comp c before
//              the superblock (c = 0..5).  450 rows for a 15 G symbol BWT: L1/L2 resident.
//
// Because blocks are addressed by POSITION (not by encoded byte offset as in the reference), a
// query needs no block directory: rank(i, c) reads exactly one aligned 64-byte record (two 32-byte
// sectors) plus one cached superblock counter.  The in-block work is branch-free popcount
// arithmetic instead of a sequential decode of up to 64 run bytes (bwt.cpp:329-338).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/bwtm.h"

namespace bwtm
{

constexpr int      SIGMA            = 6;
constexpr int      RLE_BLOCK        = 64;            // Run::BLOCK_SIZE, support.h:227
constexpr int      MAX_RUN          = 42;            // Run::MAX_RUN, support.h:229
constexpr int      RECORD_SYMBOLS   = 128;
constexpr int      RECORD_SHIFT     = 7;
constexpr int      SUPER_SHIFT      = 25;            // positions per superblock = 2^25
constexpr int      SUPER_RECORD_SHIFT = SUPER_SHIFT - RECORD_SHIFT;
constexpr uint32_t FIELD_MASK       = (1u << 25) - 1;
constexpr int      SUPER_STRIDE     = 8;             // u64 per superblock row

struct DeviceIndex
{
  const uint4*    records;   // n_records * 4 chunks
  const uint64_t* super;     // n_super * SUPER_STRIDE
  uint64_t        size;      // sequence length n
  uint64_t        sequences;
  uint64_t        C[SIGMA + 1];
};

//------------------------------------------------------------------------------
// Errors

void set_error(const char* fmt, ...);
int  cuda_failed(cudaError_t err, const char* what, const char* file, int line);
void count_launch(uint64_t n = 1);

#define BWTM_CUDA(call) do { cudaError_t err__ = (call); \
  if(err__ != cudaSuccess) { return ::bwtm::cuda_failed(err__, #call, __FILE__, __LINE__); } } while(0)

#define BWTM_TRY(call) do { int rc__ = (call); if(rc__ != BWTM_OK) { return rc__; } } while(0)

#define BWTM_LAUNCH_CHECK() do { ::bwtm::count_launch(); BWTM_CUDA(cudaGetLastError()); } while(0)

#ifdef __CUDACC__
#define BWTM_HD __host__ __device__
#else
#define BWTM_HD
#endif
BWTM_HD inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

//------------------------------------------------------------------------------
// Device functions

#ifdef __CUDACC__

__device__ __forceinline__ uint4 load_chunk(const uint4* p)
{
  return __ldg(p);
}

// Bits of the chunk's 32 positions whose comp value equals c.
__device__ __forceinline__ uint32_t match_mask(const uint4& chunk, uint32_t c)
{
  uint32_t m0 = (c & 1u) ? chunk.x : ~chunk.x;
  uint32_t m1 = (c & 2u) ? chunk.y : ~chunk.y;
  uint32_t m2 = (c & 4u) ? chunk.z : ~chunk.z;
  return m0 & m1 & m2;
}

// Mask of the first `k` bits, k in [0, 32].
__device__ __forceinline__ uint32_t low_mask(int k)
{
  return (k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u));
}

// 25-bit counter of comp c (1..5) from the four header words.
__device__ __forceinline__ uint32_t header_field(uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t c)
{
  uint64_t lo = (uint64_t)h0 | ((uint64_t)h1 << 32);
  uint64_t hi = (uint64_t)h2 | ((uint64_t)h3 << 32);
  uint32_t s = 25u * (c - 1u);
  uint64_t v;
  if(s == 0)       { v = lo; }
  else if(s < 64)  { v = (lo >> s) | (hi << (64 - s)); }
  else             { v = hi >> (s - 64); }
  return (uint32_t)v & FIELD_MASK;
}

struct Record
{
  uint4 q[4];
};

__device__ __forceinline__ Record load_record(const DeviceIndex& idx, uint64_t record)
{
  Record r;
  const uint4* p = idx.records + 4 * record;
  r.q[0] = load_chunk(p); r.q[1] = load_chunk(p + 1); r.q[2] = load_chunk(p + 2); r.q[3] = load_chunk(p + 3);
  return r;
}

// Number of c's (c = 1..5) among the first `offset` (0..127) positions of the record.
__device__ __forceinline__ uint32_t record_rank(const Record& r, uint32_t offset, uint32_t c)
{
  uint32_t res = 0;
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    int k = (int)offset - 32 * j;
    k = (k < 0 ? 0 : k);
    res += __popc(match_mask(r.q[j], c) & low_mask(k));
  }
  return res;
}

// comp value at position `offset` (0..127) of the record.
__device__ __forceinline__ uint32_t record_symbol(const Record& r, uint32_t offset)
{
  uint32_t j = offset >> 5, t = offset & 31u;
  uint4 q = (j == 0 ? r.q[0] : (j == 1 ? r.q[1] : (j == 2 ? r.q[2] : r.q[3])));
  return ((q.x >> t) & 1u) | (((q.y >> t) & 1u) << 1) | (((q.z >> t) & 1u) << 2);
}

// Count of comp c (1..5) before the record (absolute).
__device__ __forceinline__ uint64_t record_base(const DeviceIndex& idx, const Record& r, uint64_t record, uint32_t c)
{
  uint64_t sb = record >> SUPER_RECORD_SHIFT;
  return __ldg(idx.super + sb * SUPER_STRIDE + c) + header_field(r.q[0].w, r.q[1].w, r.q[2].w, r.q[3].w, c);
}

// BWT::rank(i, c) (bwt.cpp:318-341) for c = 1..5 and i <= size.
__device__ __forceinline__ uint64_t rank_nonzero(const DeviceIndex& idx, uint64_t i, uint32_t c)
{
  uint64_t record = i >> RECORD_SHIFT;
  Record r = load_record(idx, record);
  return record_base(idx, r, record, c) + record_rank(r, (uint32_t)(i & (RECORD_SYMBOLS - 1)), c);
}

// BWT::rank(i, c) for any comp value; i is clamped to size as in the reference (bwt.cpp:321-322).
__device__ __forceinline__ uint64_t rank_any(const DeviceIndex& idx, uint64_t i, uint32_t c)
{
  if(c >= SIGMA) { return 0; }
  if(i > idx.size) { i = idx.size; }
  uint64_t record = i >> RECORD_SHIFT;
  Record r = load_record(idx, record);
  uint32_t offset = (uint32_t)(i & (RECORD_SYMBOLS - 1));
  if(c != 0) { return record_base(idx, r, record, c) + record_rank(r, offset, c); }
  uint64_t others = 0;
  for(uint32_t d = 1; d < SIGMA; d++) { others += record_base(idx, r, record, d) + record_rank(r, offset, d); }
  return i - others;
}

// FMI::LF(i) (fmi.h:147-150, utils.h:335-341, BWT::inverse_select bwt.cpp:445-464):
// comp = BWT[i]; returns C[comp] + rank(i, comp). For comp == 0 the position is not computed
// (LF is undefined for the endmarker, paper.tex:141) and 0 is returned.
__device__ __forceinline__ uint64_t lf_step(const DeviceIndex& idx, uint64_t i, uint32_t& comp)
{
  uint64_t record = i >> RECORD_SHIFT;
  Record r = load_record(idx, record);
  uint32_t offset = (uint32_t)(i & (RECORD_SYMBOLS - 1));
  comp = record_symbol(r, offset);
  if(comp == 0) { return 0; }
  return idx.C[comp] + record_base(idx, r, record, comp) + record_rank(r, offset, comp);
}

#endif // __CUDACC__

} // namespace bwtm
