// bwtm_fixture: writes the BWT of a synthetic read collection to a file.
//
//   bwtm_fixture --genome G --genome-seed S --read-len L --error E --segment seed:reads[:first] [--segment ...]
//                [--format native|plain_default|...|rle] [--device D] [--chunk-reads N] --output FILE
//
// The collection is built on the GPU by the library's fixture builder (bwtm_tools_build_synthetic: counter-based
// reads, suffixes sorted by radix passes) and written in one of the reference's file formats, or as the raw
// run-length bytes ("rle").  The reference has no counterpart (README.md:21: it only merges); this exists so that
// Collections beyond the builder's memory (about 32 bytes per symbol) are built in chunks of --chunk-reads reads that
// are merged with bwtm_merge (default: chunks of at most 2^31 symbols).  This exists so that
// benchmark inputs of the named sizes can be handed to OTHER processes as files -- bench.py's reference arm runs
// the unmodified oracle/_ref/bwt_merge binary on files written by this tool and never maps the CUDA library itself.
// Prints one JSON line describing the collection.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "bwtm_host.hpp"
#include "../../include/bwtm.h"

using namespace bwtm_host;

static void die(const std::string& message)
{
  std::cerr << "bwtm_fixture: " << message << std::endl;
  std::exit(EXIT_FAILURE);
}

int main(int argc, char** argv)
{
  uint64_t genome = 0, genome_seed = 42, read_len = 0, chunk_reads = 0;
  double error = 0.0;
  int device = 0;
  std::string format = "native", output, rle_output;
  std::vector<bwtm_read_segment> segments;
  for(int k = 1; k < argc; k++)
  {
    std::string flag = argv[k];
    if(k + 1 >= argc) { die("missing value after " + flag); }
    std::string value = argv[++k];
    if(flag == "--genome") { genome = std::stoull(value); }
    else if(flag == "--genome-seed") { genome_seed = std::stoull(value); }
    else if(flag == "--read-len") { read_len = std::stoull(value); }
    else if(flag == "--error") { error = std::stod(value); }
    else if(flag == "--device") { device = std::stoi(value); }
    else if(flag == "--chunk-reads") { chunk_reads = std::stoull(value); }
    else if(flag == "--format") { format = value; }
    else if(flag == "--output") { output = value; }
    else if(flag == "--rle-output") { rle_output = value; }   // the raw run-length bytes as well
    else if(flag == "--segment")
    {
      std::istringstream fields(value); std::string field; std::vector<uint64_t> numbers;
      while(std::getline(fields, field, ':')) { numbers.push_back(std::stoull(field)); }
      if(numbers.size() < 2 || numbers.size() > 3) { die("--segment takes seed:reads[:first_read]"); }
      bwtm_read_segment segment; segment.seed = numbers[0]; segment.reads = numbers[1]; segment.first_read = (numbers.size() > 2 ? numbers[2] : 0);
      segments.push_back(segment);
    }
    else { die("unknown option " + flag); }
  }
  if(genome == 0 || read_len == 0 || segments.empty() || output.empty()) { die("--genome, --read-len, --segment and --output are required"); }
  if(format != "rle" && !formatExists(format)) { die("unknown format " + format); }

  // Same threshold as bwtm_b200/synth.py: substitute a base iff (rnd >> 11) < floor(e * 2^53).
  uint64_t threshold = (uint64_t)(error * 9007199254740992.0);
  if(bwtm_set_device(device) != BWTM_OK) { die(bwtm_last_error()); }
  if(chunk_reads == 0) { chunk_reads = std::max<uint64_t>(1, (1ull << 31) / (read_len + 1)); }
  // The segments in order, cut into chunks; every chunk is merged into what has been built so far.
  bwtm_index* index = nullptr;
  std::vector<bwtm_read_segment> chunk; uint64_t chunk_size = 0;
  auto flush = [&]()
  {
    if(chunk.empty()) { return; }
    bwtm_index* part = nullptr;
    if(bwtm_tools_build_synthetic(genome, genome_seed, read_len, threshold, chunk.data(), chunk.size(), &part) != BWTM_OK) { die(bwtm_last_error()); }
    if(index == nullptr) { index = part; }
    else
    {
      bwtm_index* merged = nullptr;
      if(bwtm_merge(index, part, nullptr, &merged, nullptr) != BWTM_OK) { die(bwtm_last_error()); }
      index = merged;
    }
    chunk.clear(); chunk_size = 0;
  };
  for(const bwtm_read_segment& segment : segments)
  {
    uint64_t done = 0;
    while(done < segment.reads)
    {
      uint64_t take = std::min(segment.reads - done, chunk_reads - chunk_size);
      bwtm_read_segment piece; piece.seed = segment.seed; piece.reads = take; piece.first_read = segment.first_read + done;
      chunk.push_back(piece); chunk_size += take; done += take;
      if(chunk_size == chunk_reads) { flush(); }
    }
  }
  flush();
  bwtm_index_info info;
  if(bwtm_index_get_info(index, &info) != BWTM_OK) { die(bwtm_last_error()); }

  HostBWT host;
  host.rle.resize(info.rle_bytes);
  if(bwtm_index_download(index, host.rle.data(), host.rle.size(), nullptr) != BWTM_OK) { die(bwtm_last_error()); }
  bwtm_index_destroy(index);
  host.sequences = info.sequences; host.bases = info.bases;
  for(size_type c = 0; c < SIGMA; c++) { host.counts[c] = info.counts[c]; }
  host.alpha = Alphabet(); host.alpha.setCounts(host.counts);

  if(!rle_output.empty())
  {
    std::ofstream out(rle_output.c_str(), std::ios_base::binary);
    if(!out) { die("cannot open " + rle_output); }
    out.write(reinterpret_cast<const char*>(host.rle.data()), host.rle.size());
  }
  if(format == "rle")
  {
    std::ofstream out(output.c_str(), std::ios_base::binary);
    if(!out) { die("cannot open " + output); }
    out.write(reinterpret_cast<const char*>(host.rle.data()), host.rle.size());
  }
  else if(!serializeBWT(host, output, format)) { return EXIT_FAILURE; }

  std::cout << "{\"file\": \"" << output << "\", \"format\": \"" << format << "\", \"sequences\": " << info.sequences
            << ", \"bases\": " << info.bases << ", \"rle_bytes\": " << info.rle_bytes << ", \"counts\": [";
  for(size_type c = 0; c < SIGMA; c++) { std::cout << (c ? ", " : "") << info.counts[c]; }
  std::cout << "]}" << std::endl;
  return 0;
}
