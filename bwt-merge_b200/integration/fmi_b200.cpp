// The reference-side binding of libbwtm_b200.so: a replacement body for the merging constructor
//
//     FMI::FMI(FMI& a, FMI& b, MergeParameters parameters)            fmi.h:107-110, fmi.cpp:336-369
//
// written against the reference's PUBLIC interface only, so that it compiles next to the unmodified
// reference sources (oracle/Makefile, target ref_b200: the reference's own bwt_merge.cpp, bwt.cpp,
// formats.cpp, support.cpp, utils.cpp and fmi.cpp are compiled where they lie; the one symbol this file
// defines is made weak in the object file of fmi.cpp and replaced at link time).  Everything else of
// bwt_merge -- option parsing, the seven file formats, the sequential multi-input loop, -v verification
// and the report -- is the reference's code, unchanged.
//
// Data crossing the seam: per input the run-length bytes BWT::data and the comp counts;
// back come the merged run-length bytes and counts.  The host-side rank/select samples of the result are
// rebuilt by the reference's own BWT::load<Format> (bwt.h:91-106: Format::read -> setHeader -> build),
// fed from memory by the small Format class below.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "fmi.h"       // the reference's header (-I/root/reference)
#include "bwtm.h"      // include/bwtm.h of this repository

namespace bwtmerge
{

namespace
{

// What Format::read hands to BWT::load: the merged bytes and counts produced by the device.
struct DeviceResult
{
  const uint8_t* bytes;
  uint64_t       size;
  uint64_t       counts[BWTM_SIGMA];
};

DeviceResult* pending_result = nullptr;

// A "file format" in the sense of formats.h:64-86 whose file is the result of bwtm_merge.
struct DeviceFormat
{
  static void read(std::ifstream&, BlockArray& data, sdsl::int_vector<64>& counts)
  {
    data.clear();
    for(uint64_t i = 0; i < pending_result->size; i++) { data.push_back(pending_result->bytes[i]); }
    counts = sdsl::int_vector<64>(BWT::SIGMA, 0);
    for(size_type c = 0; c < BWT::SIGMA; c++) { counts[c] = pending_result->counts[c]; }
  }
  inline static AlphabeticOrder order() { return AO_ANY; }
};

// BWT::data (8 MiB blocks, support.h:90-150) as one contiguous byte range, plus the comp counts.
void flatten(const FMI& fmi, std::vector<uint8_t>& bytes, uint64_t* counts)
{
  bytes.resize(fmi.bwt.bytes());
  for(size_type block = 0, offset = 0; offset < bytes.size(); block++, offset += BlockArray::BLOCK_SIZE)
  {
    size_type n = std::min<size_type>(BlockArray::BLOCK_SIZE, bytes.size() - offset);
    std::memcpy(bytes.data() + offset, fmi.bwt.data.data[block], n);
  }
  for(size_type c = 0; c < BWTM_SIGMA; c++) { counts[c] = fmi.alpha.C[c + 1] - fmi.alpha.C[c]; }
}

void deviceFailure()
{
  std::cerr << "FMI::FMI(): " << bwtm_last_error() << std::endl;
  std::exit(EXIT_FAILURE);
}

} // namespace

FMI::FMI(FMI& a, FMI& b, MergeParameters parameters)
{
  if(a.alpha != b.alpha)   // fmi.cpp:338-342
  {
    std::cerr << "FMI::FMI(): Cannot merge BWTs with different alphabets" << std::endl;
    std::exit(EXIT_FAILURE);
  }

#ifdef VERBOSE_STATUS_INFO
  std::cerr << "bwt_merge: " << a.sequences() << " sequences of total length " << a.size() << std::endl;
  std::cerr << "bwt_merge: Adding " << b.sequences() << " sequences of total length " << b.size() << std::endl;
#endif

  bwtm_index *device_a = nullptr, *device_b = nullptr;
  {
    std::vector<uint8_t> bytes_a, bytes_b;
    uint64_t counts_a[BWTM_SIGMA], counts_b[BWTM_SIGMA];
    flatten(a, bytes_a, counts_a); flatten(b, bytes_b, counts_b);
    if(bwtm_index_create_pair(bytes_a.data(), bytes_a.size(), counts_a, bytes_b.data(), bytes_b.size(), counts_b,
                              &device_a, &device_b) != BWTM_OK) { deviceFailure(); }
  }
  AlphabeticOrder order = a.bwt.header.order();
  Alphabet merged_alpha = a.alpha;
  for(size_type c = 0; c <= merged_alpha.sigma; c++) { merged_alpha.C[c] += b.alpha.C[c]; }   // fmi.cpp:367-368
  a.bwt.data.clear(); b.bwt.data.clear();   // the constructor consumes its inputs (fmi.h:107-109)

  bwtm_merge_options options; std::memset(&options, 0, sizeof(options));
  options.run_buffer_size = parameters.run_buffer_size;
  options.thread_buffer_size = parameters.thread_buffer_size;
  options.merge_buffers = parameters.merge_buffers;
  options.threads = parameters.threads;
  options.sequence_blocks = 0;             // -s counts CPU work units; the device chooses its own batches
  options.temp_dir = parameters.temp_dir.c_str();
  options.skip_index = 1;                  // the host rebuilds its own samples below

  bwtm_index* merged = nullptr; bwtm_timings timings;
  if(bwtm_merge(device_a, device_b, &options, &merged, &timings) != BWTM_OK) { deviceFailure(); }
#ifdef VERBOSE_STATUS_INFO
  std::cerr << "bwt_merge: RA built in " << (timings.search_seconds + timings.sort_seconds) << " seconds" << std::endl;
  std::cerr << "bwt_merge: BWTs merged in " << (timings.interleave_seconds + timings.encode_seconds) << " seconds" << std::endl;
#endif

  bwtm_index_info info;
  if(bwtm_index_get_info(merged, &info) != BWTM_OK) { deviceFailure(); }
  std::vector<uint8_t> bytes(info.rle_bytes);
  if(bwtm_index_download(merged, bytes.data(), bytes.size(), nullptr) != BWTM_OK) { deviceFailure(); }
  bwtm_index_destroy(merged);

  DeviceResult result; result.bytes = bytes.data(); result.size = bytes.size();
  for(size_type c = 0; c < BWTM_SIGMA; c++) { result.counts[c] = info.counts[c]; }
  pending_result = &result;
  sdsl::int_vector<64> counts;
  this->bwt.load<DeviceFormat>("/dev/null", counts);   // Format::read -> setHeader -> build (bwt.h:91-106)
  pending_result = nullptr;
  this->bwt.header.setOrder(order);                    // bwt.cpp:307
  this->alpha = merged_alpha;
}

} // namespace bwtmerge
