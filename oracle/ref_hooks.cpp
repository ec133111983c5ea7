/*
  TEST INFRASTRUCTURE -- not part of the product.

  C entry points into the UNMODIFIED reference classes (compiled from /root/reference
  against oracle/sdsl_shim by oracle/Makefile, linked into oracle/_ref/libref_hooks.so).
  Used only by tests/ to pin oracle/bwtm_oracle.c: every function below is a thin call
  into reference code (the class and member it calls is named), no algorithm lives here.
*/
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "fmi.h"

using namespace bwtmerge;

namespace
{

struct ByteVector  // the ByteArray concept of support.h: size(), push_back(), operator[]
{
  std::vector<byte_type> data;
  size_type size() const { return data.size(); }
  void push_back(byte_type b) { data.push_back(b); }
  byte_type operator[](size_type i) const { return data[i]; }
};

}

extern "C"
{

// Run::write (support.h:256-282) appended to a buffer that already holds `size` bytes.
// Returns the new size; `buffer` must have room.
uint64_t ref_run_write(uint8_t* buffer, uint64_t size, uint8_t comp, uint64_t length)
{
  ByteVector array; array.data.assign(buffer, buffer + size);
  Run::write(array, comp, length);
  std::memcpy(buffer + size, array.data.data() + size, array.data.size() - size);
  return array.data.size();
}

// Run::read (support.h:244-250)
void ref_run_read(const uint8_t* buffer, uint64_t size, uint64_t* i, uint8_t* comp, uint64_t* length)
{
  ByteVector array; array.data.assign(buffer, buffer + size);
  size_type pos = *i;
  range_type run = Run::read(array, pos);
  *i = pos; *comp = run.first; *length = run.second;
}

// ByteCode::write / ByteCode::read (support.h:172-212)
uint64_t ref_bytecode_write(uint8_t* buffer, uint64_t value)
{
  ByteVector array;
  ByteCode::write(array, value);
  std::memcpy(buffer, array.data.data(), array.data.size());
  return array.data.size();
}

uint64_t ref_bytecode_read(const uint8_t* buffer, uint64_t size, uint64_t* i)
{
  ByteVector array; array.data.assign(buffer, buffer + size);
  size_type pos = *i;
  uint64_t value = ByteCode::read(array, pos);
  *i = pos;
  return value;
}

// load(fmi, filename, format) (fmi.cpp:411-447)
void* ref_fmi_load(const char* filename, const char* format)
{
  FMI* fmi = new FMI();
  load(*fmi, filename, format);
  return fmi;
}

void ref_fmi_free(void* handle) { delete static_cast<FMI*>(handle); }

// serialize(fmi, filename, format) (fmi.cpp:373-409)
void ref_fmi_serialize(void* handle, const char* filename, const char* format)
{
  serialize(*static_cast<FMI*>(handle), filename, format);
}

uint64_t ref_fmi_size(void* handle) { return static_cast<FMI*>(handle)->size(); }
uint64_t ref_fmi_sequences(void* handle) { return static_cast<FMI*>(handle)->sequences(); }
uint64_t ref_fmi_bytes(void* handle) { return static_cast<FMI*>(handle)->bwt.bytes(); }
uint64_t ref_fmi_hash(void* handle) { return static_cast<FMI*>(handle)->bwt.hash(); }

void ref_fmi_C(void* handle, uint64_t* C7)
{
  FMI* fmi = static_cast<FMI*>(handle);
  for(size_type c = 0; c <= fmi->alpha.sigma; c++) { C7[c] = fmi->alpha.C[c]; }
}

// BWT::data bytes (BlockArray, bwt.h:173)
void ref_fmi_rle(void* handle, uint8_t* out)
{
  FMI* fmi = static_cast<FMI*>(handle);
  for(size_type i = 0; i < fmi->bwt.bytes(); i++) { out[i] = fmi->bwt.data[i]; }
}

// BWT::rank / inverse_select / ranks / operator[] (bwt.cpp:318-464)
uint64_t ref_rank(void* handle, uint64_t i, uint8_t c) { return static_cast<FMI*>(handle)->bwt.rank(i, c); }

void ref_inverse_select(void* handle, uint64_t i, uint64_t* rank, uint8_t* comp)
{
  range_type res = static_cast<FMI*>(handle)->bwt.inverse_select(i);
  *rank = res.first; *comp = res.second;
}

void ref_ranks(void* handle, uint64_t i, uint64_t* results6)
{
  BWT::ranks_type results; results[0] = 0;
  static_cast<FMI*>(handle)->bwt.ranks(i, results);
  for(size_type c = 0; c < BWT::SIGMA; c++) { results6[c] = results[c]; }
}

void ref_ranks_range(void* handle, uint64_t sp, uint64_t ep, uint64_t* first6, uint64_t* second6)
{
  BWT::rank_ranges_type results; results[0] = range_type(0, 0);
  static_cast<FMI*>(handle)->bwt.ranks(range_type(sp, ep), results);
  for(size_type c = 0; c < BWT::SIGMA; c++) { first6[c] = results[c].first; second6[c] = results[c].second; }
}

uint8_t ref_access(void* handle, uint64_t i) { return static_cast<FMI*>(handle)->bwt[i]; }

// Block samples as the reference's structures report them (bwt.cpp:324-327, support.h:338-343).
uint64_t ref_blocks(void* handle)
{
  FMI* fmi = static_cast<FMI*>(handle);
  return (fmi->bwt.bytes() + BWT::SAMPLE_RATE - 1) / BWT::SAMPLE_RATE;
}

void ref_samples(void* handle, uint64_t* block_end_out, uint64_t* cum_out)
{
  FMI* fmi = static_cast<FMI*>(handle);
  size_type blocks = ref_blocks(handle);
  for(size_type k = 0; k < blocks; k++)
  {
    block_end_out[k] = fmi->bwt.block_select(k + 1);
    for(size_type c = 0; c < BWT::SIGMA; c++) { cum_out[c * blocks + k] = fmi->bwt.samples[c].sum(k + 1); }
  }
}

// FMI::find (fmi.h:195-209) on a raw character pattern.
void ref_find(void* handle, const char* pattern, uint64_t length, uint64_t* sp, uint64_t* ep)
{
  range_type res = static_cast<FMI*>(handle)->find(pattern, length);
  *sp = res.first; *ep = res.second;
}

// FMI::FMI(a, b, parameters) (fmi.cpp:336-369). Destroys a and b (their handles must still be freed).
void* ref_merge(void* a, void* b, uint64_t threads, uint64_t sequence_blocks, const char* temp_dir)
{
  MergeParameters parameters;
  parameters.setT(threads); parameters.setSB(sequence_blocks);
  parameters.setRB(1); parameters.setTB(1);
  parameters.setTemp(temp_dir);
  parameters.sanitize();
  FMI* result = new FMI(*static_cast<FMI*>(a), *static_cast<FMI*>(b), parameters);
  return result;
}

// FMI copy constructor (fmi.h:94): lets a benchmark merge the same inputs repeatedly.
void* ref_fmi_copy(void* handle) { return new FMI(*static_cast<FMI*>(handle)); }

// FMI::FMI(a, b, parameters) with explicit MergeParameters; 0 keeps the reference default (fmi.h:49-52).
void* ref_merge_params(void* a, void* b, uint64_t threads, uint64_t sequence_blocks, uint64_t run_buffer_mb,
                       uint64_t thread_buffer_mb, uint64_t merge_buffers, const char* temp_dir)
{
  MergeParameters parameters;
  parameters.setT(threads);
  parameters.setSB(sequence_blocks > 0 ? sequence_blocks : threads * MergeParameters::BLOCKS_PER_THREAD);
  if(run_buffer_mb > 0) { parameters.setRB(run_buffer_mb); }
  if(thread_buffer_mb > 0) { parameters.setTB(thread_buffer_mb); }
  if(merge_buffers > 0) { parameters.setMB(merge_buffers); }
  parameters.setTemp(temp_dir);
  parameters.sanitize();
  Parallel::max_threads = parameters.threads;   // bwt_merge.cpp:136-137
  return new FMI(*static_cast<FMI*>(a), *static_cast<FMI*>(b), parameters);
}

} // extern "C"
