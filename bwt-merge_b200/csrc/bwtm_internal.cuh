// Host-side internals shared by the translation units of libbwtm_b200.so.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <vector>

#include "bwtm_common.cuh"

struct bwtm_index
{
  int       device;
  uint8_t*  d_rle;        // rle_bytes + padding, zero padded
  uint64_t  rle_bytes;
  uint4*    d_records;    // n_records * 4
  uint64_t  n_records;
  uint64_t* d_super;      // n_super * SUPER_STRIDE
  uint64_t  n_super;
  uint64_t  size, sequences;
  uint64_t  counts[bwtm::SIGMA];
  uint64_t  C[bwtm::SIGMA + 1];
  uint64_t  device_bytes;
  // Pair records for the two-step walk (bwtm_pairs.cu); built on first use as a merge input, may be absent.
  uint4*    d_pairs;      // n_pair_records * 8
  uint64_t  n_pair_records;
  uint64_t* d_pair_super; // n_pair_super * 32
  uint64_t  n_pair_super;
  uint64_t  pair_bytes;
};

namespace bwtm
{

constexpr uint64_t RLE_PADDING = 128;

inline DeviceIndex device_view(const bwtm_index* index)
{
  DeviceIndex v;
  v.records = index->d_records; v.super = index->d_super;
  v.size = index->size; v.sequences = index->sequences;
  for(int c = 0; c <= SIGMA; c++) { v.C[c] = index->C[c]; }
  return v;
}

uint64_t device_total_bytes();   // of the current device, cached

// Stream-ordered pool allocation (bwtm_index.cu).
int device_alloc(void** ptr, uint64_t bytes, bool resident = false);   // resident: outlives the call (an index's buffers)
int device_alloc_plain(void** ptr, uint64_t bytes);                     // cudaMalloc; freed by device_free as well
void device_free(void* ptr);

// RAII device buffer on the pool.
struct DeviceBuffer
{
  void*    ptr;
  uint64_t bytes;
  DeviceBuffer() : ptr(nullptr), bytes(0) {}
  ~DeviceBuffer() { this->release(); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  int allocate(uint64_t n, bool resident = false);   // BWTM_OK or BWTM_ERR_MEMORY
  void release();
  void* detach() { void* p = ptr; ptr = nullptr; bytes = 0; return p; }
  template<class T> T* as() const { return static_cast<T*>(ptr); }
};

// Index construction from RLE bytes already on the device. Takes ownership of d_rle on success.
int index_from_device_rle(uint8_t* d_rle, uint64_t rle_bytes, cudaStream_t stream, bwtm_index** out);
void index_free(bwtm_index* index);
// The same when the plane words of the records are already filled (headers zero): skips the run decoding.
int index_from_planes(uint8_t* d_rle, uint64_t rle_bytes, uint4* d_records, uint64_t size, cudaStream_t stream, bwtm_index** out);
// Fills the plane words of the records covering [first_position, first_position + count) from one-symbol-per-byte data.
int planes_from_symbols(const uint8_t* d_symbols, uint64_t first_position, uint64_t count, uint4* d_records, cudaStream_t stream);

// Pair records (bwtm_pairs.cu): two backward steps per record read.
uint64_t pair_index_bytes(uint64_t size);
int ensure_pair_index(bwtm_index* index, cudaStream_t stream, bool built_ahead = false);   // built_ahead: plain allocation
void release_pair_index(bwtm_index* index);
void build_pairs_ahead(bwtm_index* a, uint64_t expected_b_size, cudaStream_t stream);

// Per-64-byte-block symbol counts and their exclusive scan (block start positions).
int rle_block_starts(const uint8_t* d_rle, uint64_t rle_bytes, uint64_t* d_starts /* blocks + 1 */, cudaStream_t stream);

// Byte-exact Run::write of a device run list (bwtm_encode.cu).
struct EncoderState;

struct MergeContext;

} // namespace bwtm
