// bwt_merge_b200: a bwt_merge-compatible command line on top of the C ABI (include/bwtm.h).
//
// Same options, file formats and report text as the reference's tool (bwt_merge.cpp; the text is the
// contract that tests/test_gpu_cli.py compares), organised around three objects instead of one main():
//
//   CommandLine  a table of option specifications, scanned without getopt;
//   PatternSet   the -v patterns, flattened once for bwtm_count, with the before/after tallies;
//   Session      the device-resident index: inputs are folded into it one by one with bwtm_merge and only
//                the final result is downloaded.
//
// The merge, the rank structures and the -v queries run on the GPU.  The reference's own bwt_merge.cpp bound
// to the same library is oracle/_ref/bwt_merge_b200 (bwt-merge_b200/integration/fmi_b200.cpp).
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bwtm_host.hpp"
#include "../../include/bwtm.h"

using namespace bwtm_host;

namespace
{

[[noreturn]] void die(const std::string& message)
{
  std::cerr << message << std::endl;
  std::exit(EXIT_FAILURE);
}

[[noreturn]] void dieWithDeviceError(const char* where)
{
  die(std::string(where) + ": " + bwtm_last_error());
}

size_type hardwareThreads() { return std::max(1u, std::thread::hardware_concurrency()); }

//------------------------------------------------------------------------------
// Command line

struct CommandLine
{
  // MergeParameters (fmi.h:45-80): the CPU buffer knobs are parsed, reported and handed to the library, which
  // ignores them (there are no run, thread or merge buffers and no temporary files on the device).
  size_type run_buffer_runs = 8 * MEGABYTE, thread_buffer_bytes = 256 * MEGABYTE, merge_buffers = 6;
  size_type threads = hardwareThreads(), sequence_blocks = 4 * hardwareThreads();
  std::string temp_dir = ".", pattern_file, output_format = "native", output;
  std::vector<std::string> inputs, input_formats;

  struct Spec { char letter; const char* argument; std::string help; std::function<void(const std::string&)> apply; };
  std::vector<Spec> specs;
  std::vector<char> group_breaks;   // letters after which the usage text has an empty line

  CommandLine()
  {
    auto number = [](const std::string& text) -> size_type { return std::stoul(text); };
    specs = {
      { 'b', "N", "Set thread buffer size to N megabytes / thread (default: 256)",
        [=](const std::string& v) { thread_buffer_bytes = number(v) * MEGABYTE; } },
      { 'm', "N", "Set the number of merge buffers to N (default: 6)",
        [=](const std::string& v) { merge_buffers = number(v); } },
      { 'r', "N", "Set run buffer size to N megabytes / thread (default: 128)",
        [=](const std::string& v) { run_buffer_runs = number(v) * MEGABYTE / 16; } },
      { 's', "N", "Set the number of sequence blocks to N (default: 4 / thread)",
        [=](const std::string& v) { sequence_blocks = number(v); } },
      { 't', "N", "Use N parallel threads (default: " + std::to_string(hardwareThreads()) + " on this system)",
        [=](const std::string& v) { threads = number(v); } },
      { 'd', "directory", "Use the given directory for temporary files (default: .)",
        [=](const std::string& v) { temp_dir = (v.empty() ? "." : (v.back() == '/' ? v.substr(0, v.size() - 1) : v)); } },
      { 'v', "filename", "Verify by querying with patterns from the given file",
        [=](const std::string& v) { pattern_file = v; } },
      { 'i', "formats", "Read the inputs in the given formats (default: native)\n                Multiple comma-separated formats can be provided.",
        [=](const std::string& v)
        {
          std::istringstream list(v); std::string tag;
          while(std::getline(list, tag, ','))
          {
            if(!formatExists(tag)) { die("bwt_merge: Invalid input format: " + tag); }
            input_formats.push_back(tag);
          }
        } },
      { 'o', "format", "Write the output in the given format (default: native)",
        [=](const std::string& v)
        {
          if(!formatExists(v)) { die("bwt_merge: Invalid output format: " + v); }
          output_format = v;
        } },
    };
    group_breaks = { 't', 'v', 'o' };
  }

  void usage() const
  {
    std::cerr << "Usage: bwt_merge [options] input1 input2 [input3 ...] output" << std::endl << std::endl;
    std::cerr << "Options:" << std::endl;
    for(const Spec& spec : specs)
    {
      std::string head = std::string("  -") + spec.letter + " " + spec.argument;
      head.resize(16, ' ');
      std::cerr << head << spec.help << std::endl;
      for(char c : group_breaks) { if(c == spec.letter) { std::cerr << std::endl; } }
    }
    printFormats(std::cerr);
  }

  void parse(int argc, char** argv)
  {
    std::vector<std::string> files;
    bool options_done = false;
    for(int k = 1; k < argc; k++)
    {
      std::string word = argv[k];
      if(options_done || word.size() < 2 || word[0] != '-') { files.push_back(word); continue; }
      if(word == "--") { options_done = true; continue; }
      const Spec* found = nullptr;
      for(const Spec& spec : specs) { if(spec.letter == word[1]) { found = &spec; } }
      if(found == nullptr) { die(std::string(argv[0]) + ": invalid option -- '" + word[1] + "'"); }
      std::string value = word.substr(2);
      if(value.empty())
      {
        if(k + 1 >= argc) { die(std::string(argv[0]) + ": option requires an argument -- '" + word[1] + "'"); }
        value = argv[++k];
      }
      found->apply(value);
    }

    if(files.size() < 3) { die("bwt_merge: Output file not specified"); }
    output = files.back(); files.pop_back();
    inputs.swap(files);
    if(input_formats.empty()) { input_formats.assign(inputs.size(), "native"); }
    else if(input_formats.size() == 1) { input_formats.resize(inputs.size(), input_formats.front()); }
    else if(input_formats.size() != inputs.size())
    {
      die("bwt_merge: Specified " + std::to_string(input_formats.size()) + " formats for " + std::to_string(inputs.size()) + " inputs");
    }

    // MergeParameters::sanitize, fmi.cpp:462-468
    threads = std::max<size_type>(std::min(threads, hardwareThreads()), 1);
    sequence_blocks = std::max<size_type>(sequence_blocks, 1);
    threads = std::min(threads, sequence_blocks);
  }

  void report() const   // bwt_merge.cpp:144-152 and operator<<(MergeParameters), fmi.cpp:484-495
  {
    for(size_type k = 0; k < inputs.size(); k++) { std::cout << "Input:            " << inputs[k] << " (" << input_formats[k] << ")" << std::endl; }
    std::cout << "Output:           " << output << " (" << output_format << ")" << std::endl;
    if(!pattern_file.empty()) { std::cout << "Patterns:         " << pattern_file << std::endl; }
    std::cout << std::endl;
    std::cout << "Run buffers:      " << ((run_buffer_runs * 16) / 1048576.0) << " MB" << std::endl;
    std::cout << "Thread buffers:   " << (thread_buffer_bytes / 1048576.0) << " MB" << std::endl;
    std::cout << "Merge buffers:    " << merge_buffers << std::endl;
    std::cout << "Threads:          " << threads << std::endl;
    std::cout << "Sequence blocks:  " << sequence_blocks << std::endl;
    std::cout << "Temp directory:   " << temp_dir << std::endl;
    std::cout << std::endl;
  }

  bwtm_merge_options mergeOptions() const
  {
    bwtm_merge_options options; std::memset(&options, 0, sizeof(options));
    options.run_buffer_size = run_buffer_runs; options.thread_buffer_size = thread_buffer_bytes;
    options.merge_buffers = merge_buffers; options.threads = threads; options.temp_dir = temp_dir.c_str();
    // -s counts CPU work units (4 per thread by default): it is reported but not forwarded. With 0 the library
    // splits the search into as many batches as the free device memory requires (include/bwtm.h).
    options.sequence_blocks = 0;
    return options;
  }
};

//------------------------------------------------------------------------------
// -v patterns (verifyFMI / queryFMI, bwt_merge.cpp:240-285)

struct PatternSet
{
  bool active = false;
  std::vector<std::string> rows;
  std::string flat;
  std::vector<size_type> offsets, before, after;
  size_type chars = 0;

  void read(const std::string& filename)
  {
    active = true;
    chars = readRows(filename, rows, true);
    offsets.assign(1, 0);
    for(const std::string& row : rows) { flat += row; offsets.push_back(flat.size()); }
    before.assign(rows.size(), 0); after.assign(rows.size(), 0);
    std::cout << "Read " << rows.size() << " patterns of total length " << chars << std::endl << std::endl;
  }

  void verdict() const
  {
    if(!active) { return; }
    size_type errors = 0;
    for(size_type k = 0; k < rows.size(); k++) { errors += (before[k] != after[k]); }
    if(errors > 0) { std::cout << "Verification failed for " << errors << " patterns" << std::endl; }
    else { std::cout << "Verification successful" << std::endl; }
    std::cout << std::endl;
  }
};

//------------------------------------------------------------------------------
// The device-resident index

class Session
{
  struct Loaded { bwtm_index* handle; Alphabet alpha; size_type native_size; };

public:
  explicit Session(const CommandLine& cl) : options(cl), handle(nullptr), native_size(0), bases_added(0)
  {
    if(bwtm_set_device(0) != BWTM_OK) { dieWithDeviceError("bwt_merge"); }
  }
  ~Session() { if(handle != nullptr) { bwtm_index_destroy(handle); } }

  size_type added() const { return bases_added; }

  // load(fmi, filename, format), fmi.cpp:373-409. RopeBWT and SGA files go to the device as they are and are
  // decoded there; a file with zero-length codes takes the host reader, which treats it the way the reference
  // does. The other formats are read on the host (PlainData::read joins equal characters, not equal comp values,
  // which only the host sees).
  void add(const std::string& filename, const std::string& format, PatternSet& patterns)
  {
    Loaded next = load(filename, format);
    if(handle == nullptr)
    {
      handle = next.handle; alpha = next.alpha; native_size = next.native_size;
      query("Input", next, patterns, patterns.before);
      return;
    }
    bwtm_index* increment = next.handle;
    size_type increment_size = info(increment).bases;
    bases_added += increment_size;
    query("Input", next, patterns, patterns.before);

    double start = readTimer();
    if(!alpha.sameMaps(next.alpha)) { die("FMI::FMI(): Cannot merge BWTs with different alphabets"); }   // fmi.cpp:338-342
    bwtm_merge_options merge_options = options.mergeOptions();
    bwtm_index* merged = nullptr; bwtm_timings timings;
    if(bwtm_merge(handle, increment, &merge_options, &merged, &timings) != BWTM_OK) { handle = nullptr; dieWithDeviceError("FMI::FMI()"); }
#ifdef VERBOSE_STATUS_INFO
    std::cerr << "bwt_merge: RA built in " << (timings.search_seconds + timings.sort_seconds) << " seconds" << std::endl;
    std::cerr << "bwt_merge: BWTs merged in " << (timings.interleave_seconds + timings.encode_seconds) << " seconds" << std::endl;
    std::cerr << "bwt_merge: rank/select built in " << timings.index_seconds << " seconds" << std::endl;
#endif
    handle = merged;
    for(size_type c = 0; c <= SIGMA; c++) { alpha.C[c] += next.alpha.C[c]; }   // fmi.cpp:367-368
    bwtm_index_info merged_info = info(handle);
    native_size = nativeSize(merged_info.rle_bytes, merged_info.bases, merged_info.counts);
    double seconds = readTimer() - start;
    std::cout << "BWTs merged in " << seconds << " seconds (" << ((increment_size / 1048576.0) / seconds) << " MB/s)" << std::endl << std::endl;
  }

  // serialize(index, filename, format), then the queries on the result.
  void finish(const std::string& filename, const std::string& format, PatternSet& patterns)
  {
    bwtm_index_info result_info = info(handle);
    HostBWT result;
    result.rle.resize(result_info.rle_bytes);
    if(bwtm_index_download(handle, result.rle.data(), result.rle.size(), nullptr) != BWTM_OK) { dieWithDeviceError("serialize()"); }
    result.sequences = result_info.sequences; result.bases = result_info.bases;
    for(size_type c = 0; c < SIGMA; c++) { result.counts[c] = result_info.counts[c]; }
    result.alpha = alpha;
    serializeBWT(result, filename, format);
    Loaded current; current.handle = handle; current.alpha = alpha; current.native_size = native_size;
    query("Output", current, patterns, patterns.after);
  }

private:
  static bwtm_index_info info(const bwtm_index* index)
  {
    bwtm_index_info result;
    if(bwtm_index_get_info(index, &result) != BWTM_OK) { dieWithDeviceError("bwt_merge"); }
    return result;
  }

  Loaded load(const std::string& filename, const std::string& format)
  {
    Loaded result; result.handle = nullptr; result.native_size = 0;
    if(format == "ropebwt" || format == "sga")
    {
      std::vector<byte_type> runs;
      if(!loadRunBytes(filename, format, runs)) { std::exit(EXIT_FAILURE); }
      int rc = (runs.empty() ? BWTM_ERR_ALPHABET
                             : bwtm_index_create_runs(runs.data(), runs.size(), (format == "sga" ? BWTM_RUNS_SGA : BWTM_RUNS_ROPEBWT), 0, &result.handle));
      if(rc == BWTM_OK)
      {
        bwtm_index_info loaded = info(result.handle);
        result.alpha = Alphabet::create(formatOrder(format)); result.alpha.setCounts(loaded.counts);
        result.native_size = nativeSize(loaded.rle_bytes, loaded.bases, loaded.counts);
        return result;
      }
      if(rc != BWTM_ERR_ALPHABET) { dieWithDeviceError("load()"); }
    }
    HostBWT host;
    if(!loadBWT(host, filename, format)) { std::exit(EXIT_FAILURE); }
    result.alpha = host.alpha; result.native_size = host.nativeSize();
    if(bwtm_index_create(host.rle.data(), host.rle.size(), host.counts, &result.handle) != BWTM_OK) { dieWithDeviceError("load()"); }
    return result;
  }

  // One "Input:" / "Output:" block of the report: size line, then the pattern counts added to `tally`.
  void query(const std::string& name, const Loaded& index, const PatternSet& patterns, std::vector<size_type>& tally) const
  {
    printSize(name, index.native_size, info(index.handle).bases);
    if(patterns.chars > 0)
    {
      double start = readTimer();
      std::vector<size_type> counts(patterns.rows.size(), 0);
      if(bwtm_count(index.handle, reinterpret_cast<const uint8_t*>(patterns.flat.data()), patterns.offsets.data(), patterns.rows.size(),
                    index.alpha.char2comp, counts.data()) != BWTM_OK) { dieWithDeviceError("verifyFMI()"); }
      size_type matches = 0;
      for(size_type k = 0; k < counts.size(); k++) { tally[k] += counts[k]; matches += counts[k]; }
      // bwt_merge.cpp:255 tests the emptiness of the LOOP range, so "found" is always the number of patterns.
      printTime(name, patterns.rows.size(), matches, patterns.chars, readTimer() - start);
    }
    std::cout << std::endl;
  }

  const CommandLine& options;
  bwtm_index* handle;
  Alphabet    alpha;
  size_type   native_size;   // sdsl::size_in_bytes(fmi) of the equivalent reference object
  size_type   bases_added;
};

} // namespace

int main(int argc, char** argv)
{
  CommandLine command_line;
  if(argc < 2) { command_line.usage(); return EXIT_SUCCESS; }

  double start = readTimer();
  std::cout << "BWT-merge" << std::endl << std::endl;
  command_line.parse(argc, argv);
  command_line.report();

  PatternSet patterns;
  if(!command_line.pattern_file.empty()) { patterns.read(command_line.pattern_file); }

  Session session(command_line);
  for(size_type k = 0; k < command_line.inputs.size(); k++) { session.add(command_line.inputs[k], command_line.input_formats[k], patterns); }
  session.finish(command_line.output, command_line.output_format, patterns);
  patterns.verdict();

  double seconds = readTimer() - start;
  std::cout << "Total time:       " << seconds << " seconds (" << ((session.added() / 1048576.0) / seconds) << " MB/s)" << std::endl;
  std::cout << "Peak memory:      " << (memoryUsage() / 1073741824.0) << " GB" << std::endl;
  std::cout << std::endl;
  return 0;
}
