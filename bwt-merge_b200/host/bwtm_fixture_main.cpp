// bwtm_fixture: writes the BWT of a synthetic read collection to a file.
//
//   bwtm_fixture --genome G --genome-seed S --read-len L --error E --segment seed:reads[:first] [--segment ...]
//                [--format native|plain_default|...|rle] [--device D] --output FILE
//
// The collection is built on the GPU by the library's fixture builder (bwtm_tools_build_synthetic: counter-based
// reads, suffixes sorted by radix passes) and written in one of the reference's file formats, or as the raw
// run-length bytes ("rle").  The reference has no counterpart (README.md:21: it only merges); this exists so that
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
  uint64_t genome = 0, genome_seed = 42, read_len = 0;
  double error = 0.0;
  int device = 0;
  std::string format = "native", output;
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
    else if(flag == "--format") { format = value; }
    else if(flag == "--output") { output = value; }
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
  bwtm_index* index = nullptr;
  if(bwtm_tools_build_synthetic(genome, genome_seed, read_len, threshold, segments.data(), segments.size(), &index) != BWTM_OK) { die(bwtm_last_error()); }
  bwtm_index_info info;
  if(bwtm_index_get_info(index, &info) != BWTM_OK) { die(bwtm_last_error()); }

  HostBWT host;
  host.rle.resize(info.rle_bytes);
  if(bwtm_index_download(index, host.rle.data(), host.rle.size(), nullptr) != BWTM_OK) { die(bwtm_last_error()); }
  bwtm_index_destroy(index);
  host.sequences = info.sequences; host.bases = info.bases;
  for(size_type c = 0; c < SIGMA; c++) { host.counts[c] = info.counts[c]; }
  host.alpha = Alphabet(); host.alpha.setCounts(host.counts);

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
