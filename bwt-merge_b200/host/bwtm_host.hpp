// Host side of the B200 rank-array path: BWT file formats and alphabets, exactly as the reference
// defines them (formats.h, formats.cpp, support.cpp), above the C ABI of include/bwtm.h.
//
// This is I/O and transcoding only: files <-> the run-length byte code the device
// consumes and produces. Nothing here ranks, searches or merges.
#ifndef BWTM_HOST_HPP
#define BWTM_HOST_HPP

#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace bwtm_host
{

typedef std::uint64_t size_type;
typedef std::uint8_t  byte_type;

const size_type SIGMA = 6;              // Run::SIGMA, support.h:228
const size_type BLOCK_SIZE = 64;        // Run::BLOCK_SIZE, support.h:227
const size_type MAX_RUN = 42;           // Run::MAX_RUN, support.h:229
const size_type MEGABYTE = 1048576;
const size_type ARRAY_BLOCK = 8 * MEGABYTE;   // BlockArray::BLOCK_SIZE, support.h:95

enum AlphabeticOrder { AO_DEFAULT = 0, AO_SORTED = 1, AO_ANY = 254, AO_UNKNOWN = 255 };   // formats.h:35

// Alphabet (support.h:41-84): char <-> comp maps and the C array.
struct Alphabet
{
  byte_type char2comp[256];
  byte_type comp2char[SIGMA];
  size_type C[SIGMA + 1];
  size_type sigma;

  Alphabet();                                   // default alphabet $ACGTN, support.cpp:40-63
  static Alphabet create(AlphabeticOrder order); // createAlphabet, formats.cpp:34-53
  static Alphabet identity(size_type sigma);    // Alphabet(size_type), support.cpp:93-113
  void setCounts(const size_type* counts);      // support.cpp:90
  bool sorted() const;                          // support.cpp:182-190
  bool sameMaps(const Alphabet& another) const; // operator==, support.cpp:192-205
  AlphabeticOrder identify() const;             // identifyAlphabet, formats.cpp:55-64
};

std::string alphabetName(AlphabeticOrder order);
bool compatible(const Alphabet& alpha, AlphabeticOrder order);

// One BWT on the host: the run-length bytes plus what the native header and the alphabet carry.
struct HostBWT
{
  std::vector<byte_type> rle;        // BWT::data
  size_type sequences, bases;        // NativeHeader
  size_type counts[SIGMA];
  Alphabet  alpha;

  HostBWT();
  AlphabeticOrder order() const { return alpha.identify(); }
  size_type nativeSize() const;      // sdsl::size_in_bytes(fmi): bytes the native serialization takes
};

// The same from the shape alone (RLE byte count, sequence length, per-comp counts).
size_type nativeSize(size_type rle_bytes, size_type bases, const size_type* counts);

// Run code (support.h:160-286), host copy used by the transcoders.
struct RunWriter
{
  std::vector<byte_type>& out;
  size_type value, length;           // RunBuffer state, utils.h:121-142
  size_type counts[SIGMA];
  explicit RunWriter(std::vector<byte_type>& target);
  void add(size_type comp, size_type n = 1);     // RunBuffer::add + Run::write of completed runs
  void finish();
private:
  void flushRun();
};

void writeRun(std::vector<byte_type>& out, size_type comp, size_type length);      // Run::write
bool readRun(const std::vector<byte_type>& in, size_type& pos, size_type& comp, size_type& length); // Run::read

// Block samples of BWT::build (bwt.cpp:476-512): last position and six cumulative counts per 64-byte block.
struct Samples
{
  size_type blocks;
  std::vector<size_type> block_ends;           // blocks
  std::vector<size_type> cumulative[SIGMA];    // blocks each, counts through the block
};
Samples computeSamples(const std::vector<byte_type>& rle, size_type* counts_out, size_type* bases_out);

// Formats (formats.h:64-156). Tags as in the reference.
bool formatExists(const std::string& tag);
void printFormats(std::ostream& stream);
AlphabeticOrder formatOrder(const std::string& tag);
std::string formatName(const std::string& tag);

// load(fmi, filename, format) / serialize(fmi, filename, format), fmi.cpp:373-447. Return false after
// printing the reference's message to stderr when the file cannot be opened or has a bad header.
bool loadBWT(HostBWT& bwt, const std::string& filename, const std::string& format);
bool serializeBWT(const HostBWT& bwt, const std::string& filename, const std::string& format);
// Raw run bytes of a RopeBWT / SGA file (header checked and skipped) for the device-side reader.
bool loadRunBytes(const std::string& filename, const std::string& format, std::vector<byte_type>& runs);

// Reporting helpers with the reference's formatting (utils.cpp:38-96).
void printHeader(const std::string& header, size_type indent = 18);
void printSize(const std::string& header, size_type bytes, size_type data_size);
void printTime(const std::string& header, size_type found, size_type matches, size_type bytes, double seconds);
double readTimer();
size_type memoryUsage();
size_type readRows(const std::string& filename, std::vector<std::string>& rows, bool skip_empty_rows);

} // namespace bwtm_host

#endif // BWTM_HOST_HPP
