// File formats, alphabets and report helpers of the host driver. Behaviour follows the reference
// (formats.cpp, support.cpp, utils.cpp); the code is this repository's own.
#include "bwtm_host.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include <sys/resource.h>

namespace bwtm_host
{

//------------------------------------------------------------------------------
// Alphabet

Alphabet::Alphabet()
{
  // \0 and $ are endmarkers, ACGT / acgt are bases, everything else is N (support.cpp:35-63).
  for(size_type c = 0; c < 256; c++) { char2comp[c] = 5; }
  char2comp[0] = 0; char2comp[(unsigned)'$'] = 0;
  const char* bases = "ACGT";
  for(size_type k = 0; k < 4; k++)
  {
    char2comp[(unsigned)bases[k]] = (byte_type)(k + 1);
    char2comp[(unsigned)(bases[k] - 'A' + 'a')] = (byte_type)(k + 1);
  }
  const char* chars = "$ACGTN";
  for(size_type c = 0; c < SIGMA; c++) { comp2char[c] = (byte_type)chars[c]; }
  for(size_type c = 0; c <= SIGMA; c++) { C[c] = 0; }
  sigma = SIGMA;
}

Alphabet Alphabet::create(AlphabeticOrder order)
{
  Alphabet alpha;
  if(order == AO_SORTED)   // $ACGNT, formats.cpp:43-47
  {
    std::swap(alpha.comp2char[4], alpha.comp2char[5]);
    std::swap(alpha.char2comp[(unsigned)'N'], alpha.char2comp[(unsigned)'T']);
    std::swap(alpha.char2comp[(unsigned)'n'], alpha.char2comp[(unsigned)'t']);
  }
  return alpha;
}

Alphabet Alphabet::identity(size_type _sigma)
{
  Alphabet alpha;
  alpha.sigma = _sigma;
  for(size_type c = 0; c < 256; c++) { alpha.char2comp[c] = (byte_type)(c < _sigma ? c : 0); }
  for(size_type c = 0; c < SIGMA; c++) { alpha.comp2char[c] = (byte_type)c; }
  return alpha;
}

void Alphabet::setCounts(const size_type* counts)
{
  C[0] = 0;
  for(size_type c = 0; c < SIGMA; c++) { C[c + 1] = C[c] + counts[c]; }
}

bool Alphabet::sorted() const
{
  for(size_type c = 1; c < sigma; c++) { if(comp2char[c - 1] >= comp2char[c]) { return false; } }
  return true;
}

bool Alphabet::sameMaps(const Alphabet& another) const
{
  if(sigma != another.sigma) { return false; }
  return std::memcmp(char2comp, another.char2comp, 256) == 0 && std::memcmp(comp2char, another.comp2char, SIGMA) == 0;
}

AlphabeticOrder Alphabet::identify() const
{
  if(this->sorted()) { return AO_SORTED; }
  Alphabet default_alpha;
  if(this->sameMaps(default_alpha)) { return AO_DEFAULT; }
  return AO_UNKNOWN;
}

std::string alphabetName(AlphabeticOrder order)
{
  switch(order)
  {
    case AO_DEFAULT: return "default";
    case AO_SORTED:  return "sorted";
    case AO_ANY:     return "any";
    default:         return "unknown";
  }
}

bool compatible(const Alphabet& alpha, AlphabeticOrder order)
{
  Alphabet default_alpha;
  switch(order)
  {
    case AO_DEFAULT: return alpha.sameMaps(default_alpha);
    case AO_SORTED:  return alpha.sorted();
    case AO_ANY:     return true;
    default:         return false;
  }
}

//------------------------------------------------------------------------------
// Run code

static size_type bitLength(size_type value)   // utils.h:146-151, bit_length(0) == 1
{
  size_type bits = 1;
  while(value > 1) { value >>= 1; bits++; }
  return bits;
}

static void writeByteCode(std::vector<byte_type>& out, size_type value)   // support.h:203-212
{
  while(value > 0x7F) { out.push_back((byte_type)((value & 0x7F) | 0x80)); value >>= 7; }
  out.push_back((byte_type)value);
}

void writeRun(std::vector<byte_type>& out, size_type comp, size_type length)   // support.h:256-282
{
  while(length > 0)
  {
    if(length < MAX_RUN) { out.push_back((byte_type)(comp + SIGMA * (length - 1))); return; }
    size_type room = BLOCK_SIZE - (out.size() % BLOCK_SIZE);
    size_type head = (room > 1 ? MAX_RUN : MAX_RUN - 1);
    out.push_back((byte_type)(comp + SIGMA * (head - 1)));
    length -= head; room--;
    if(room > 0)
    {
      size_type extension = length;
      if(bitLength(length) > 7 * room) { extension = (((size_type)1) << (7 * room)) - 1; }
      writeByteCode(out, extension);
      length -= extension;
    }
  }
}

bool readRun(const std::vector<byte_type>& in, size_type& pos, size_type& comp, size_type& length)   // support.h:244-250
{
  if(pos >= in.size()) { return false; }
  byte_type code = in[pos++];
  comp = code % SIGMA; length = code / SIGMA + 1;
  if(length >= MAX_RUN)
  {
    size_type shift = 0;
    while(pos < in.size())
    {
      byte_type b = in[pos++];
      length += ((size_type)(b & 0x7F)) << shift; shift += 7;
      if(!(b & 0x80)) { break; }
    }
  }
  return true;
}

RunWriter::RunWriter(std::vector<byte_type>& target) : out(target), value(0), length(0)
{
  for(size_type c = 0; c < SIGMA; c++) { counts[c] = 0; }
}

void RunWriter::flushRun()
{
  if(length > 0) { writeRun(out, value, length); counts[value] += length; }
}

void RunWriter::add(size_type comp, size_type n)
{
  if(comp == value) { length += n; return; }
  this->flushRun();
  value = comp; length = n;
}

void RunWriter::finish() { this->flushRun(); length = 0; }

//------------------------------------------------------------------------------
// Samples

Samples computeSamples(const std::vector<byte_type>& rle, size_type* counts_out, size_type* bases_out)
{
  Samples s;
  s.blocks = (rle.size() + BLOCK_SIZE - 1) / BLOCK_SIZE;
  s.block_ends.reserve(s.blocks);
  for(size_type c = 0; c < SIGMA; c++) { s.cumulative[c].reserve(s.blocks); }
  size_type cumulative[SIGMA] = { 0, 0, 0, 0, 0, 0 };
  size_type seq_pos = 0, pos = 0, comp = 0, length = 0;
  while(readRun(rle, pos, comp, length))
  {
    seq_pos += length; cumulative[comp] += length;
    if(pos >= rle.size() || pos % BLOCK_SIZE == 0)      // bwt.cpp:494
    {
      s.block_ends.push_back(seq_pos - 1);
      for(size_type c = 0; c < SIGMA; c++) { s.cumulative[c].push_back(cumulative[c]); }
    }
  }
  if(counts_out != 0) { for(size_type c = 0; c < SIGMA; c++) { counts_out[c] = cumulative[c]; } }
  if(bases_out != 0) { *bases_out = seq_pos; }
  return s;
}

//------------------------------------------------------------------------------
// Sparse bitvector serialization. The reference stores its samples in SDSL sd_vectors, whose byte layout
// SDSL does not specify and the reference does not pin. This writer produces the
// layout of oracle/sdsl_shim: {u64 size, u64 ones, u8 low width, low words, high words}, so files are
// byte-comparable with oracle/_ref; with real SDSL only this function and its reader would change.

static size_type highBit(size_type x) { size_type h = 0; while(x > 1) { x >>= 1; h++; } return h; }

struct SparseLayout
{
  size_type size, ones, low_width, high_bits, low_words, high_words;
  SparseLayout(size_type n, size_type m) : size(n), ones(m)
  {
    size_type logm = highBit(m) + 1, logn = highBit(n) + 1;
    if(logm == logn) { logm--; }
    low_width = logn - logm;
    high_bits = m + (((size_type)1) << logm);
    low_words = (m * low_width + 63) / 64;
    high_words = (high_bits + 63) / 64;
  }
  size_type bytes() const { return 8 + 8 + 1 + 8 * (low_words + high_words); }
};

template<class T> static void writePod(std::ostream& out, const T& value) { out.write(reinterpret_cast<const char*>(&value), sizeof(T)); }
template<class T> static bool readPod(std::istream& in, T& value) { in.read(reinterpret_cast<char*>(&value), sizeof(T)); return (bool)in; }

// positions[k] + shift * k are the (strictly increasing) one bits.
static void writeSparse(std::ostream& out, size_type universe, const std::vector<size_type>& positions, size_type shift)
{
  SparseLayout layout(universe, positions.size());
  std::vector<size_type> low(layout.low_words + 1, 0), high(layout.high_words + 1, 0);
  size_type mask = (layout.low_width == 0 ? 0 : (~(size_type)0) >> (64 - layout.low_width));
  for(size_type k = 0; k < positions.size(); k++)
  {
    size_type value = positions[k] + shift * k;
    if(layout.low_width > 0)
    {
      size_type bit = k * layout.low_width, part = value & mask;
      low[bit / 64] |= part << (bit % 64);
      if(bit % 64 + layout.low_width > 64) { low[bit / 64 + 1] |= part >> (64 - bit % 64); }
    }
    size_type high_pos = (value >> layout.low_width) + k;
    high[high_pos / 64] |= ((size_type)1) << (high_pos % 64);
  }
  writePod(out, layout.size); writePod(out, layout.ones);
  byte_type width = (byte_type)layout.low_width; writePod(out, width);
  out.write(reinterpret_cast<const char*>(low.data()), 8 * layout.low_words);
  out.write(reinterpret_cast<const char*>(high.data()), 8 * layout.high_words);
}


//------------------------------------------------------------------------------
// HostBWT

HostBWT::HostBWT() : sequences(0), bases(0)
{
  for(size_type c = 0; c < SIGMA; c++) { counts[c] = 0; }
}

static size_type paddedVector(size_type elements, size_type element_bytes)   // int_vector file: u64 bits + data padded to 8
{
  size_type bytes = elements * element_bytes;
  return 8 + ((bytes + 7) / 8) * 8;
}

size_type nativeSize(size_type rle_bytes, size_type bases, const size_type* counts)
{
  size_type blocks = (rle_bytes + BLOCK_SIZE - 1) / BLOCK_SIZE;
  size_type total = 24;                                                           // NativeHeader
  total += 8 + ((rle_bytes + ARRAY_BLOCK - 1) / ARRAY_BLOCK) * ARRAY_BLOCK;       // BlockArray
  for(size_type c = 0; c < SIGMA; c++) { total += SparseLayout(counts[c] + blocks, blocks).bytes() + 8; }
  total += SparseLayout(bases, blocks).bytes();
  total += paddedVector(256, 1) + paddedVector(SIGMA, 1) + paddedVector(SIGMA + 1, 8) + 8;   // Alphabet
  return total;
}

size_type HostBWT::nativeSize() const { return bwtm_host::nativeSize(rle.size(), bases, counts); }

//------------------------------------------------------------------------------
// Formats

struct FormatInfo { const char* tag; const char* name; AlphabeticOrder order; };

static const FormatInfo FORMATS[] =
{
  { "native",        "Native format",                    AO_ANY },
  { "plain_default", "Plain format (default alphabet)",  AO_DEFAULT },
  { "plain_sorted",  "Plain format (sorted alphabet)",   AO_SORTED },
  { "rfm",           "RFM format",                       AO_SORTED },
  { "sdsl",          "SDSL format",                      AO_SORTED },
  { "ropebwt",       "RopeBWT format",                   AO_DEFAULT },
  { "sga",           "SGA format",                       AO_DEFAULT },
};

static const FormatInfo* findFormat(const std::string& tag)
{
  for(const FormatInfo& f : FORMATS) { if(tag == f.tag) { return &f; } }
  return 0;
}

bool formatExists(const std::string& tag) { return findFormat(tag) != 0; }
AlphabeticOrder formatOrder(const std::string& tag) { const FormatInfo* f = findFormat(tag); return (f ? f->order : AO_UNKNOWN); }
std::string formatName(const std::string& tag) { const FormatInfo* f = findFormat(tag); return (f ? f->name : ""); }

static void printFormat(std::ostream& stream, const char* tag)
{
  std::string t(tag), padding;
  if(t.length() < 15) { padding = std::string(15 - t.length(), ' '); }
  stream << "  " << t << padding << formatName(t) << std::endl;
}

void printFormats(std::ostream& stream)   // formats.cpp:462-480
{
  stream << "Formats supporting any alphabetic order:" << std::endl;
  printFormat(stream, "native");
  stream << std::endl;
  stream << "Formats using the default alphabet:" << std::endl;
  printFormat(stream, "plain_default"); printFormat(stream, "ropebwt"); printFormat(stream, "sga");
  stream << std::endl;
  stream << "Formats using sorted alphabet:" << std::endl;
  printFormat(stream, "plain_sorted"); printFormat(stream, "rfm"); printFormat(stream, "sdsl");
  stream << std::endl;
}

static size_type remainingBytes(std::ifstream& in)
{
  std::streamoff here = in.tellg();
  in.seekg(0, std::ios::end);
  std::streamoff end = in.tellg();
  in.seekg(here, std::ios::beg);
  return (size_type)(end - here);
}

// PlainData::read (formats.cpp:133-161): maximal runs of equal CHARACTERS, mapped to comps run by run.
static void readSymbols(std::ifstream& in, size_type bytes, const Alphabet& map, HostBWT& bwt)
{
  RunWriter writer(bwt.rle);
  std::vector<byte_type> buffer(MEGABYTE);
  size_type run_char = 0, run_length = 0;
  for(size_type offset = 0; offset < bytes; offset += MEGABYTE)
  {
    size_type n = std::min(MEGABYTE, bytes - offset);
    in.read(reinterpret_cast<char*>(buffer.data()), n);
    for(size_type i = 0; i < n; i++)
    {
      if(buffer[i] == run_char) { run_length++; continue; }
      if(run_length > 0) { writeRun(bwt.rle, map.char2comp[run_char], run_length); writer.counts[map.char2comp[run_char]] += run_length; }
      run_char = buffer[i]; run_length = 1;
    }
  }
  if(run_length > 0) { writeRun(bwt.rle, map.char2comp[run_char], run_length); writer.counts[map.char2comp[run_char]] += run_length; }
  for(size_type c = 0; c < SIGMA; c++) { bwt.counts[c] = writer.counts[c]; }
}

static void writeSymbols(std::ofstream& out, const HostBWT& bwt, const Alphabet& map, bool pad_to_words)
{
  std::vector<byte_type> buffer; buffer.reserve(MEGABYTE + 64);
  size_type pos = 0, comp = 0, length = 0, written = 0;
  while(readRun(bwt.rle, pos, comp, length))
  {
    byte_type ch = map.comp2char[comp];
    while(length > 0)
    {
      size_type n = std::min(length, MEGABYTE - buffer.size());
      buffer.insert(buffer.end(), n, ch); length -= n;
      if(buffer.size() >= MEGABYTE) { out.write(reinterpret_cast<const char*>(buffer.data()), buffer.size()); written += buffer.size(); buffer.clear(); }
    }
  }
  out.write(reinterpret_cast<const char*>(buffer.data()), buffer.size()); written += buffer.size();
  if(pad_to_words && written % 8 != 0) { const char zeros[8] = { 0 }; out.write(zeros, 8 - written % 8); }
}

// RopeData::read (formats.cpp:286-310): (comp, length) codes, coalesced into maximal runs.
// Returns false on a comp value above 5: the reference would index past its count array with it (RunBuffer::add,
// Run::write), here it is a malformed file.
static bool readRuns(std::ifstream& in, size_type bytes, bool sga, HostBWT& bwt)
{
  RunWriter writer(bwt.rle);
  std::vector<byte_type> buffer(MEGABYTE);
  for(size_type offset = 0; offset < bytes; offset += MEGABYTE)
  {
    size_type n = std::min(MEGABYTE, bytes - offset);
    in.read(reinterpret_cast<char*>(buffer.data()), n);
    for(size_type i = 0; i < n; i++)
    {
      size_type comp = (sga ? buffer[i] >> 5 : buffer[i] & 0x07), length = (sga ? buffer[i] & 0x1F : buffer[i] >> 3);
      if(comp >= SIGMA)
      {
        std::cerr << (sga ? "SGAFormat" : "RopeFormat") << "::load(): Invalid character value " << comp << " at run " << (offset + i) << std::endl;
        return false;
      }
      // RunBuffer::add(v, n) (utils.h:125-134): a run of another value flushes the pending one, even if n == 0.
      writer.add(comp, length);
    }
  }
  writer.finish();
  for(size_type c = 0; c < SIGMA; c++) { bwt.counts[c] = writer.counts[c]; }
  return true;
}

static size_type countShortRuns(const HostBWT& bwt)   // RopeData::countRuns, formats.cpp:343-363
{
  size_type pos = 0, comp = 0, length = 0, runs = 0;
  while(readRun(bwt.rle, pos, comp, length)) { runs += (length + 30) / 31; }
  return runs;
}

static void writeRuns(std::ofstream& out, const HostBWT& bwt, bool sga)   // RopeData::write, formats.cpp:312-338
{
  std::vector<byte_type> buffer; buffer.reserve(MEGABYTE + 8);
  size_type pos = 0, comp = 0, length = 0;
  while(readRun(bwt.rle, pos, comp, length))
  {
    while(length > 0)
    {
      size_type n = std::min<size_type>(length, 31);
      buffer.push_back((byte_type)(sga ? ((comp << 5) | n) : ((n << 3) | comp)));
      length -= n;
      if(buffer.size() >= MEGABYTE) { out.write(reinterpret_cast<const char*>(buffer.data()), buffer.size()); buffer.clear(); }
    }
  }
  out.write(reinterpret_cast<const char*>(buffer.data()), buffer.size());
}

static void finishLoaded(HostBWT& bwt, AlphabeticOrder order)   // BWT::setHeader + FMI::load<Format>, fmi.h:126-134
{
  bwt.sequences = bwt.counts[0];
  bwt.bases = 0;
  for(size_type c = 0; c < SIGMA; c++) { bwt.bases += bwt.counts[c]; }
  bwt.alpha = Alphabet::create(order);
  bwt.alpha.setCounts(bwt.counts);
}

// Native files (bwt.cpp:111-148, fmi.cpp: serialize): NativeHeader, BlockArray, seven sparse vectors with their
// support structures, Alphabet. Only the header, the run-length bytes and the alphabet are needed here (the rank
// structure is rebuilt on the device), and the sparse vectors are the one part whose byte layout belongs to SDSL
// (unpinned). So they are not parsed at all: the Alphabet is the LAST section and has a fixed
// size (int_vector<8>[256], int_vector<8>[6], int_vector<64>[7], u64 sigma = 352 bytes, support.cpp:160-171) and is
// read from the end of the file. A file written by a reference built with the real SDSL loads the same way.
static bool loadNative(HostBWT& bwt, std::ifstream& in)
{
  std::uint32_t tag = 0, flags = 0;
  readPod(in, tag); readPod(in, flags); readPod(in, bwt.sequences); readPod(in, bwt.bases);
  if(!in || tag != 0x54574221u) { std::cerr << "BWT::load(): Invalid header!" << std::endl; return false; }
  size_type bytes = 0; readPod(in, bytes);
  const size_type stored = ((bytes + ARRAY_BLOCK - 1) / ARRAY_BLOCK) * ARRAY_BLOCK;
  const size_type ALPHABET_BYTES = (8 + 256) + (8 + 8) + (8 + 8 * (SIGMA + 1)) + 8;
  const std::streampos payload = in.tellg();
  in.seekg(0, std::ios::end);
  const size_type file_size = (size_type)in.tellg();
  if(!in || (size_type)payload + stored + ALPHABET_BYTES > file_size)
  {
    std::cerr << "BWT::load(): Truncated native file (" << bytes << " bytes of run-length code announced)" << std::endl;
    return false;
  }
  in.seekg(payload);
  bwt.rle.resize(bytes);
  in.read(reinterpret_cast<char*>(bwt.rle.data()), bytes);

  in.seekg(file_size - ALPHABET_BYTES);
  size_type bits_char2comp = 0, bits_comp2char = 0, bits_C = 0;
  byte_type padded[8];
  readPod(in, bits_char2comp); in.read(reinterpret_cast<char*>(bwt.alpha.char2comp), 256);
  readPod(in, bits_comp2char); in.read(reinterpret_cast<char*>(padded), 8);
  readPod(in, bits_C); in.read(reinterpret_cast<char*>(bwt.alpha.C), 8 * (SIGMA + 1));
  readPod(in, bwt.alpha.sigma);
  for(size_type c = 0; c < SIGMA; c++) { bwt.alpha.comp2char[c] = padded[c]; }
  bool plausible = (bool)in && bits_char2comp == 8 * 256 && bits_comp2char == 8 * SIGMA && bits_C == 64 * (SIGMA + 1) && bwt.alpha.sigma == SIGMA &&
                   bwt.alpha.C[0] == 0 && bwt.alpha.C[SIGMA] == bwt.bases;
  for(size_type c = 0; plausible && c < SIGMA; c++) { plausible = (bwt.alpha.C[c] <= bwt.alpha.C[c + 1]); }
  if(!plausible)
  {
    std::cerr << "BWT::load(): Cannot find the alphabet at the end of the native file: not written by bwt_merge / bwt_convert, or with another alphabet size" << std::endl;
    return false;
  }
  for(size_type c = 0; c < SIGMA; c++) { bwt.counts[c] = bwt.alpha.C[c + 1] - bwt.alpha.C[c]; }
  return true;
}

static void serializeNative(const HostBWT& bwt, std::ofstream& out)
{
  std::uint32_t tag = 0x54574221u, flags = (std::uint32_t)bwt.order() & 0xFF;    // formats.cpp:488-533
  writePod(out, tag); writePod(out, flags); writePod(out, bwt.sequences); writePod(out, bwt.bases);

  size_type bytes = bwt.rle.size();                                               // BlockArray::serialize, support.cpp:296-309
  writePod(out, bytes);
  out.write(reinterpret_cast<const char*>(bwt.rle.data()), bytes);
  size_type stored = ((bytes + ARRAY_BLOCK - 1) / ARRAY_BLOCK) * ARRAY_BLOCK;
  std::vector<char> zeros(std::min<size_type>(stored - bytes, MEGABYTE), 0);
  for(size_type left = stored - bytes; left > 0; ) { size_type n = std::min<size_type>(left, zeros.size()); out.write(zeros.data(), n); left -= n; }

  size_type counts[SIGMA], bases = 0;
  Samples samples = computeSamples(bwt.rle, counts, &bases);
  for(size_type c = 0; c < SIGMA; c++)                                            // CumulativeArray::serialize, support.cpp:442-454
  {
    writeSparse(out, counts[c] + samples.blocks, samples.cumulative[c], 1);       // bit cumulative_c(k) + k, bwt.cpp:497-500
    writePod(out, samples.blocks);
  }
  writeSparse(out, bases, samples.block_ends, 0);                                 // block_boundaries, bwt.cpp:496

  size_type bits = 256 * 8; writePod(out, bits); out.write(reinterpret_cast<const char*>(bwt.alpha.char2comp), 256);
  bits = SIGMA * 8; writePod(out, bits); out.write(reinterpret_cast<const char*>(bwt.alpha.comp2char), SIGMA);
  const char pad[8] = { 0 }; out.write(pad, 8 - SIGMA);
  bits = (SIGMA + 1) * 64; writePod(out, bits); out.write(reinterpret_cast<const char*>(bwt.alpha.C), 8 * (SIGMA + 1));
  writePod(out, bwt.alpha.sigma);
}

bool loadBWT(HostBWT& bwt, const std::string& filename, const std::string& format)
{
  bwt = HostBWT();
  std::ifstream in(filename.c_str(), std::ios_base::binary);
  if(!in)
  {
    std::cerr << (format == "native" ? "FMI::load(): " : "BWT::load(): ") << "Cannot open input file " << filename << std::endl;
    return false;
  }
  if(format == "native") { return loadNative(bwt, in); }
  if(format == "plain_default" || format == "plain_sorted")
  {
    readSymbols(in, remainingBytes(in), Alphabet::create(formatOrder(format)), bwt);
  }
  else if(format == "rfm" || format == "sdsl")   // int_vector<8> file: u64 bit count, data (formats.cpp:248-277)
  {
    size_type bits = 0; readPod(in, bits);
    readSymbols(in, bits / 8, (format == "rfm" ? Alphabet::identity(SIGMA) : Alphabet::create(AO_SORTED)), bwt);
  }
  else if(format == "ropebwt")
  {
    std::uint32_t tag = 0; readPod(in, tag);
    if(!in || tag != 0x06454C52u) { std::cerr << "RopeFormat::load(): Invalid header!" << std::endl; return false; }
    if(!readRuns(in, remainingBytes(in), false, bwt)) { return false; }
  }
  else if(format == "sga")
  {
    std::uint16_t tag = 0; size_type sequences = 0, bases = 0, bytes = 0; std::uint32_t flags = 0;
    readPod(in, tag); readPod(in, sequences); readPod(in, bases); readPod(in, bytes); readPod(in, flags);
    if(!in || tag != 0xCACA || flags != 0) { std::cerr << "SGAFormat::load(): Invalid header!" << std::endl; return false; }
    if(!readRuns(in, std::min(bytes, remainingBytes(in)), true, bwt)) { return false; }
  }
  else { std::cerr << "load(): Invalid BWT format: " << format << std::endl; return false; }
  finishLoaded(bwt, formatOrder(format));
  return true;
}

// The run bytes of a RopeBWT or SGA file as they are (one (comp, length) code per byte), for the device-side
// reader bwtm_index_create_runs. Same header checks and messages as loadBWT.
bool loadRunBytes(const std::string& filename, const std::string& format, std::vector<byte_type>& runs)
{
  runs.clear();
  std::ifstream in(filename.c_str(), std::ios_base::binary);
  if(!in) { std::cerr << "BWT::load(): Cannot open input file " << filename << std::endl; return false; }
  size_type bytes = 0;
  if(format == "ropebwt")
  {
    std::uint32_t tag = 0; readPod(in, tag);
    if(!in || tag != 0x06454C52u) { std::cerr << "RopeFormat::load(): Invalid header!" << std::endl; return false; }
    bytes = remainingBytes(in);
  }
  else if(format == "sga")
  {
    std::uint16_t tag = 0; size_type sequences = 0, bases = 0; std::uint32_t flags = 0;
    readPod(in, tag); readPod(in, sequences); readPod(in, bases); readPod(in, bytes); readPod(in, flags);
    if(!in || tag != 0xCACA || flags != 0) { std::cerr << "SGAFormat::load(): Invalid header!" << std::endl; return false; }
    bytes = std::min(bytes, remainingBytes(in));   // the header is not trusted with the allocation
  }
  else { std::cerr << "load(): Invalid BWT format: " << format << std::endl; return false; }
  runs.resize(bytes);
  in.read(reinterpret_cast<char*>(runs.data()), bytes);
  runs.resize(in.gcount() > 0 ? (size_type)in.gcount() : 0);
  return true;
}

bool serializeBWT(const HostBWT& bwt, const std::string& filename, const std::string& format)
{
  if(!formatExists(format)) { std::cerr << "serialize(): Invalid BWT format: " << format << std::endl; return false; }
  if(format != "native" && !compatible(bwt.alpha, formatOrder(format)))   // fmi.h:117-122
  {
    std::cerr << "FMI::serialize(): Warning: " << formatName(format) << " is not compatible with "
              << alphabetName(bwt.alpha.identify()) << " alphabets!" << std::endl;
  }
  std::ofstream out(filename.c_str(), std::ios_base::binary);
  if(!out)
  {
    std::cerr << (format == "native" ? "FMI::serialize(): " : "BWT::serialize(): ") << "Cannot open output file " << filename << std::endl;
    return false;
  }
  if(format == "native") { serializeNative(bwt, out); }
  else if(format == "plain_default" || format == "plain_sorted") { writeSymbols(out, bwt, Alphabet::create(formatOrder(format)), false); }
  else if(format == "rfm" || format == "sdsl")
  {
    size_type bits = bwt.bases * 8; writePod(out, bits);
    writeSymbols(out, bwt, (format == "rfm" ? Alphabet::identity(SIGMA) : Alphabet::create(AO_SORTED)), true);
  }
  else if(format == "ropebwt")
  {
    std::uint32_t tag = 0x06454C52u; writePod(out, tag);
    writeRuns(out, bwt, false);
  }
  else   // sga: formats.cpp:431-445
  {
    std::uint16_t tag = 0xCACA; std::uint32_t flags = 0; size_type runs = countShortRuns(bwt);
    writePod(out, tag); writePod(out, bwt.sequences); writePod(out, bwt.bases); writePod(out, runs); writePod(out, flags);
    writeRuns(out, bwt, true);
  }
  return true;
}

//------------------------------------------------------------------------------
// Reports (utils.cpp:38-96)

void printHeader(const std::string& header, size_type indent)
{
  std::string padding;
  if(header.length() + 1 < indent) { padding = std::string(indent - 1 - header.length(), ' '); }
  std::cout << header << ":" << padding;
}

void printSize(const std::string& header, size_type bytes, size_type data_size)
{
  printHeader(header);
  std::cout << (bytes / 1048576.0) << " MB (" << ((8.0 * bytes) / data_size) << " bpc)" << std::endl;
}

void printTime(const std::string& header, size_type found, size_type matches, size_type bytes, double seconds)
{
  printHeader(header);
  std::cout << "Found " << found << " patterns with " << matches << " occ in "
            << seconds << " seconds (" << ((bytes / 1048576.0) / seconds) << " MB/s)" << std::endl;
}

double readTimer()
{
  std::chrono::duration<double> now = std::chrono::steady_clock::now().time_since_epoch();
  return now.count();
}

size_type memoryUsage()
{
  rusage usage;
  getrusage(RUSAGE_SELF, &usage);
  return 1024 * (size_type)usage.ru_maxrss;
}

size_type readRows(const std::string& filename, std::vector<std::string>& rows, bool skip_empty_rows)   // utils.cpp:100-122
{
  std::ifstream input(filename.c_str(), std::ios_base::binary);
  if(!input) { std::cerr << "readRows(): Cannot open input file " << filename << std::endl; return 0; }
  size_type chars = 0;
  while(input)
  {
    std::string row;
    std::getline(input, row);
    if(skip_empty_rows && row.length() == 0) { continue; }
    rows.push_back(row); chars += row.length();
  }
  return chars;
}

} // namespace bwtm_host
