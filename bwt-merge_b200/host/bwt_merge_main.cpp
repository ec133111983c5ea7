// bwt_merge_b200: the reference's bwt_merge command line (bwt_merge.cpp:47-203) on top of the C ABI.
// Options, file formats, report lines and the sequential multi-input loop are the reference's; the merge,
// the rank structures and the -v queries run on the GPU (include/bwtm.h). The merged index stays on the
// device between merges; only the final result is downloaded.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "bwtm_host.hpp"
#include "../../include/bwtm.h"

using namespace bwtm_host;

namespace
{

struct Parameters   // MergeParameters, fmi.h:45-80
{
  size_type run_buffer_size, thread_buffer_size, merge_buffers, threads, sequence_blocks;
  std::string temp_dir;
  Parameters()
  {
    size_type hardware = std::max(1u, std::thread::hardware_concurrency());
    run_buffer_size = 8 * MEGABYTE; thread_buffer_size = 256 * MEGABYTE; merge_buffers = 6;
    threads = hardware; sequence_blocks = threads * 4; temp_dir = ".";
  }
  void sanitize()   // fmi.cpp:462-468
  {
    size_type hardware = std::max(1u, std::thread::hardware_concurrency());
    threads = std::max<size_type>(std::min(threads, hardware), 1);
    sequence_blocks = std::max<size_type>(sequence_blocks, 1);
    threads = std::min(threads, sequence_blocks);
  }
  void setTemp(const std::string& directory)   // fmi.cpp:470-476
  {
    if(directory.length() == 0) { temp_dir = "."; }
    else if(directory[directory.length() - 1] != '/') { temp_dir = directory; }
    else { temp_dir = directory.substr(0, directory.length() - 1); }
  }
};

std::ostream& operator<<(std::ostream& stream, const Parameters& p)   // fmi.cpp:484-495
{
  stream << "Run buffers:      " << ((p.run_buffer_size * 16) / 1048576.0) << " MB" << std::endl;
  stream << "Thread buffers:   " << (p.thread_buffer_size / 1048576.0) << " MB" << std::endl;
  stream << "Merge buffers:    " << p.merge_buffers << std::endl;
  stream << "Threads:          " << p.threads << std::endl;
  stream << "Sequence blocks:  " << p.sequence_blocks << std::endl;
  stream << "Temp directory:   " << p.temp_dir << std::endl;
  return stream;
}

void fail(const char* where)
{
  std::cerr << where << ": " << bwtm_last_error() << std::endl;
  std::exit(EXIT_FAILURE);
}

void printUsage()
{
  Parameters defaults;
  std::cerr << "Usage: bwt_merge [options] input1 input2 [input3 ...] output" << std::endl << std::endl;
  std::cerr << "Options:" << std::endl;
  std::cerr << "  -b N          Set thread buffer size to N megabytes / thread (default: " << 256 << ")" << std::endl;
  std::cerr << "  -m N          Set the number of merge buffers to N (default: " << 6 << ")" << std::endl;
  std::cerr << "  -r N          Set run buffer size to N megabytes / thread (default: " << 128 << ")" << std::endl;
  std::cerr << "  -s N          Set the number of sequence blocks to N (default: " << 4 << " / thread)" << std::endl;
  std::cerr << "  -t N          Use N parallel threads (default: " << defaults.threads << " on this system)" << std::endl;
  std::cerr << std::endl;
  std::cerr << "  -d directory  Use the given directory for temporary files (default: .)" << std::endl;
  std::cerr << "  -v filename   Verify by querying with patterns from the given file" << std::endl;
  std::cerr << std::endl;
  std::cerr << "  -i formats    Read the inputs in the given formats (default: native)" << std::endl;
  std::cerr << "                Multiple comma-separated formats can be provided." << std::endl;
  std::cerr << "  -o format     Write the output in the given format (default: native)" << std::endl;
  std::cerr << std::endl;
  printFormats(std::cerr);
}

// An index on the device plus what only the host knows about it.
struct DeviceFMI
{
  bwtm_index* handle;
  Alphabet    alpha;
  size_type   native_size;   // sdsl::size_in_bytes(fmi) of the equivalent reference object
  DeviceFMI() : handle(0), native_size(0) {}
};

DeviceFMI upload(HostBWT& host)
{
  DeviceFMI fmi;
  fmi.alpha = host.alpha; fmi.native_size = host.nativeSize();
  if(bwtm_index_create(host.rle.data(), host.rle.size(), host.counts, &fmi.handle) != BWTM_OK) { fail("load()"); }
  std::vector<byte_type>().swap(host.rle);
  return fmi;
}

// load(fmi, filename, format), fmi.cpp:373-409. RopeBWT and SGA files go to the device as they are and are decoded
// there; a file the device reader refuses (zero-length codes) takes the host reader, which treats it
// the way the reference does. The other formats are read on the host (PlainData::read joins equal characters, not
// equal comp values, which only the host sees).
DeviceFMI loadInput(const std::string& filename, const std::string& format)
{
  if(format == "ropebwt" || format == "sga")
  {
    std::vector<byte_type> runs;
    if(!loadRunBytes(filename, format, runs)) { std::exit(EXIT_FAILURE); }
    DeviceFMI fmi;
    int rc = (runs.empty() ? BWTM_ERR_ALPHABET : bwtm_index_create_runs(runs.data(), runs.size(), (format == "sga" ? BWTM_RUNS_SGA : BWTM_RUNS_ROPEBWT), 0, &fmi.handle));
    if(rc == BWTM_OK)
    {
      bwtm_index_info info; bwtm_index_get_info(fmi.handle, &info);
      fmi.alpha = Alphabet::create(formatOrder(format)); fmi.alpha.setCounts(info.counts);
      fmi.native_size = nativeSize(info.rle_bytes, info.bases, info.counts);
      return fmi;
    }
    if(rc != BWTM_ERR_ALPHABET) { fail("load()"); }
  }
  HostBWT host;
  if(!loadBWT(host, filename, format)) { std::exit(EXIT_FAILURE); }
  return upload(host);
}

// verifyFMI + queryFMI (bwt_merge.cpp:240-285): adds the occurrences of every pattern to results.
void verifyFMI(const DeviceFMI& fmi, const std::string& name, const std::vector<std::string>& patterns, std::vector<size_type>& results)
{
  bwtm_index_info info; bwtm_index_get_info(fmi.handle, &info);
  size_type chars = 0;
  for(const std::string& p : patterns) { chars += p.length(); }
  printSize(name, fmi.native_size, info.bases);
  if(chars > 0)
  {
    double start = readTimer();
    std::vector<size_type> offsets(patterns.size() + 1, 0), counts(patterns.size(), 0);
    std::string flat; flat.reserve(chars);
    for(size_type i = 0; i < patterns.size(); i++) { flat += patterns[i]; offsets[i + 1] = flat.size(); }
    if(bwtm_count(fmi.handle, reinterpret_cast<const uint8_t*>(flat.data()), offsets.data(), patterns.size(),
                  fmi.alpha.char2comp, counts.data()) != BWTM_OK) { fail("verifyFMI()"); }
    size_type matches = 0;
    for(size_type i = 0; i < patterns.size(); i++) { results[i] += counts[i]; matches += counts[i]; }
    double seconds = readTimer() - start;
    // bwt_merge.cpp:255 tests the emptiness of the LOOP range, so "found" is always the number of patterns.
    printTime(name, patterns.size(), matches, chars, seconds);
  }
  std::cout << std::endl;
}

void tokenize(const std::string& source, std::vector<std::string>& tokens, char delim)
{
  std::istringstream ss(source);
  std::string token;
  while(std::getline(ss, token, delim)) { tokens.push_back(token); }
}

} // namespace

int main(int argc, char** argv)
{
  if(argc < 2) { printUsage(); std::exit(EXIT_SUCCESS); }

  double start = readTimer();
  std::cout << "BWT-merge" << std::endl << std::endl;

  int c = 0;
  bool verify = false;
  Parameters parameters;
  std::string pattern_name, output_format;
  std::vector<std::string> input_formats;
  while((c = getopt(argc, argv, "b:m:r:s:t:d:v:i:o:")) != -1)
  {
    switch(c)
    {
    case 'b': parameters.thread_buffer_size = std::stoul(optarg) * MEGABYTE; break;
    case 'm': parameters.merge_buffers = std::stoul(optarg); break;
    case 'r': parameters.run_buffer_size = std::stoul(optarg) * MEGABYTE / 16; break;
    case 's': parameters.sequence_blocks = std::stoul(optarg); break;
    case 't': parameters.threads = std::stoul(optarg); break;
    case 'd': parameters.setTemp(optarg); break;
    case 'v': pattern_name = optarg; verify = true; break;
    case 'i':
      tokenize(optarg, input_formats, ',');
      for(const std::string& format : input_formats)
      {
        if(!formatExists(format)) { std::cerr << "bwt_merge: Invalid input format: " << format << std::endl; std::exit(EXIT_FAILURE); }
      }
      break;
    case 'o':
      output_format = optarg;
      if(!formatExists(output_format)) { std::cerr << "bwt_merge: Invalid output format: " << output_format << std::endl; std::exit(EXIT_FAILURE); }
      break;
    default: std::exit(EXIT_FAILURE);
    }
  }

  int inputs = (argc - 1) - optind;
  if(inputs < 2) { std::cerr << "bwt_merge: Output file not specified" << std::endl; std::exit(EXIT_FAILURE); }
  if(input_formats.empty()) { input_formats.assign(inputs, "native"); }
  if(input_formats.size() == 1) { input_formats.resize(inputs, input_formats[0]); }
  if(input_formats.size() != (unsigned)inputs)
  {
    std::cerr << "bwt_merge: Specified " << input_formats.size() << " formats for " << inputs << " inputs" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  if(output_format.length() == 0) { output_format = "native"; }
  parameters.sanitize();

  for(int i = optind; i < argc - 1; i++)
  {
    std::cout << "Input:            " << argv[i] << " (" << input_formats[i - optind] << ")" << std::endl;
  }
  std::cout << "Output:           " << argv[argc - 1] << " (" << output_format << ")" << std::endl;
  if(verify) { std::cout << "Patterns:         " << pattern_name << std::endl; }
  std::cout << std::endl << parameters << std::endl;

  std::vector<std::string> patterns;
  std::vector<size_type> pre_results, post_results;
  if(verify)
  {
    size_type chars = readRows(pattern_name, patterns, true);
    pre_results.assign(patterns.size(), 0); post_results.assign(patterns.size(), 0);
    std::cout << "Read " << patterns.size() << " patterns of total length " << chars << std::endl << std::endl;
  }

  if(bwtm_set_device(0) != BWTM_OK) { fail("bwt_merge"); }

  DeviceFMI index = loadInput(argv[optind], input_formats[0]);
  verifyFMI(index, "Input", patterns, pre_results);

  size_type bytes_added = 0;
  for(int input = 1; input < inputs; input++)
  {
    DeviceFMI increment = loadInput(argv[optind + input], input_formats[input]);
    size_type increment_size = 0;
    { bwtm_index_info info; bwtm_index_get_info(increment.handle, &info); increment_size = info.bases; }
    bytes_added += increment_size;
    verifyFMI(increment, "Input", patterns, pre_results);

    // merge(), bwt_merge.cpp:287-299
    double merge_start = readTimer();
    if(!index.alpha.sameMaps(increment.alpha))   // fmi.cpp:338-342
    {
      std::cerr << "FMI::FMI(): Cannot merge BWTs with different alphabets" << std::endl;
      std::exit(EXIT_FAILURE);
    }
    bwtm_merge_options options; std::memset(&options, 0, sizeof(options));
    options.run_buffer_size = parameters.run_buffer_size; options.thread_buffer_size = parameters.thread_buffer_size;
    options.merge_buffers = parameters.merge_buffers; options.threads = parameters.threads;
    options.sequence_blocks = parameters.sequence_blocks; options.temp_dir = parameters.temp_dir.c_str();
    bwtm_index* merged = 0; bwtm_timings timings;
    if(bwtm_merge(index.handle, increment.handle, &options, &merged, &timings) != BWTM_OK) { fail("FMI::FMI()"); }
#ifdef VERBOSE_STATUS_INFO
    std::cerr << "bwt_merge: RA built in " << (timings.search_seconds + timings.sort_seconds) << " seconds" << std::endl;
    std::cerr << "bwt_merge: BWTs merged in " << (timings.interleave_seconds + timings.encode_seconds) << " seconds" << std::endl;
    std::cerr << "bwt_merge: rank/select built in " << timings.index_seconds << " seconds" << std::endl;
#endif
    index.handle = merged;
    for(size_type k = 0; k <= SIGMA; k++) { index.alpha.C[k] += increment.alpha.C[k]; }   // fmi.cpp:367-368
    {
      bwtm_index_info info; bwtm_index_get_info(merged, &info);
      index.native_size = nativeSize(info.rle_bytes, info.bases, info.counts);
    }
    double seconds = readTimer() - merge_start;
    std::cout << "BWTs merged in " << seconds << " seconds (" << ((increment_size / 1048576.0) / seconds) << " MB/s)" << std::endl << std::endl;
  }

  // serialize(index, ...), bwt_merge.cpp:175
  bwtm_index_info info; bwtm_index_get_info(index.handle, &info);
  HostBWT result;
  result.rle.resize(info.rle_bytes);
  if(bwtm_index_download(index.handle, result.rle.data(), result.rle.size(), 0) != BWTM_OK) { fail("serialize()"); }
  result.sequences = info.sequences; result.bases = info.bases;
  for(size_type k = 0; k < SIGMA; k++) { result.counts[k] = info.counts[k]; }
  result.alpha = index.alpha;
  serializeBWT(result, argv[argc - 1], output_format);
  verifyFMI(index, "Output", patterns, post_results);

  if(verify)
  {
    size_type errors = 0;
    for(size_type i = 0; i < patterns.size(); i++) { if(pre_results[i] != post_results[i]) { errors++; } }
    if(errors > 0) { std::cout << "Verification failed for " << errors << " patterns" << std::endl; }
    else { std::cout << "Verification successful" << std::endl; }
    std::cout << std::endl;
  }

  double seconds = readTimer() - start;
  std::cout << "Total time:       " << seconds << " seconds (" << ((bytes_added / 1048576.0) / seconds) << " MB/s)" << std::endl;
  std::cout << "Peak memory:      " << (memoryUsage() / 1073741824.0) << " GB" << std::endl;
  std::cout << std::endl;

  bwtm_index_destroy(index.handle);
  return 0;
}
