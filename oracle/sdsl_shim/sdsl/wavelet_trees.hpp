/*
  TEST INFRASTRUCTURE -- not part of the product.

  Header-only stand-in for the subset of SDSL 2.x (simongog/sdsl-lite, version
  not pinned by the reference: Makefile:1 expects ../sdsl-lite, README.md:17
  says "SDSL 2.0") that the unmodified bwt-merge sources use through
  `#include <sdsl/wavelet_trees.hpp>` (utils.h:37).  SDSL itself is not
  available in this environment (no network), so the reference is compiled
  against this file into oracle/_ref/ by oracle/Makefile.

  What is faithful:
    * semantics of every call the reference makes (rank/select on a sparse
      bitvector, int_vector, int_vector_buffer<8>, write_member/read_member);
    * int_vector<8/64> serialization = u64 bit count + data padded to 8 bytes,
      which the reference itself documents in utils.h:374-407.
  What is NOT pinned: the serialized byte layout of sd_vector and its
  supports.  The layout written here is {u64 size, u64 ones, u8 wl, low words,
  high words}; select supports serialize to zero bytes and are rebuilt on
  load.  Native files are therefore byte-comparable only between programs
  that use this same layout (oracle/_ref and this repo's host writer).

  sd_vector is a real Elias-Fano encoding (Elias 1974, Fano 1971; Okanohara &
  Sadakane 2007) with sampled constant-time select on the high bits, written
  from the published algorithm, so that the CPU baseline timed through it is
  not handicapped by a toy bitvector.
*/
#ifndef BWTM_ORACLE_SDSL_SHIM_WAVELET_TREES_HPP
#define BWTM_ORACLE_SDSL_SHIM_WAVELET_TREES_HPP

#include <algorithm>
#include <array>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <initializer_list>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include <unistd.h>

namespace sdsl
{

//------------------------------------------------------------------------------
// bits

struct bits
{
  static const std::uint64_t lo_set[65];

  static inline std::uint32_t hi(std::uint64_t x)
  {
    return (x == 0 ? 0 : 63 - __builtin_clzll(x));
  }

  static inline std::uint32_t cnt(std::uint64_t x) { return __builtin_popcountll(x); }

  // Position of the k-th (1-based) set bit of x; x must have at least k set bits.
  static inline std::uint32_t sel(std::uint64_t x, std::uint32_t k)
  {
#if defined(__BMI2__)
    return __builtin_ctzll(__builtin_ia32_pdep_di(std::uint64_t(1) << (k - 1), x));
#else
    for(std::uint32_t i = 1; i < k; i++) { x &= x - 1; }
    return __builtin_ctzll(x);
#endif
  }
};

// Header-only definition of bits::lo_set: a weak symbol, so every translation unit may carry it.
__attribute__((weak)) const std::uint64_t bits::lo_set[65] =
{
  0x0ULL,
  0x1ULL, 0x3ULL, 0x7ULL, 0xFULL, 0x1FULL, 0x3FULL, 0x7FULL, 0xFFULL,
  0x1FFULL, 0x3FFULL, 0x7FFULL, 0xFFFULL, 0x1FFFULL, 0x3FFFULL, 0x7FFFULL, 0xFFFFULL,
  0x1FFFFULL, 0x3FFFFULL, 0x7FFFFULL, 0xFFFFFULL, 0x1FFFFFULL, 0x3FFFFFULL, 0x7FFFFFULL, 0xFFFFFFULL,
  0x1FFFFFFULL, 0x3FFFFFFULL, 0x7FFFFFFULL, 0xFFFFFFFULL, 0x1FFFFFFFULL, 0x3FFFFFFFULL, 0x7FFFFFFFULL, 0xFFFFFFFFULL,
  0x1FFFFFFFFULL, 0x3FFFFFFFFULL, 0x7FFFFFFFFULL, 0xFFFFFFFFFULL,
  0x1FFFFFFFFFULL, 0x3FFFFFFFFFULL, 0x7FFFFFFFFFULL, 0xFFFFFFFFFFULL,
  0x1FFFFFFFFFFULL, 0x3FFFFFFFFFFULL, 0x7FFFFFFFFFFULL, 0xFFFFFFFFFFFULL,
  0x1FFFFFFFFFFFULL, 0x3FFFFFFFFFFFULL, 0x7FFFFFFFFFFFULL, 0xFFFFFFFFFFFFULL,
  0x1FFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFULL,
  0x1FFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFULL,
  0x1FFFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFULL,
  0x1FFFFFFFFFFFFFFFULL, 0x3FFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL
};

//------------------------------------------------------------------------------
// structure tree (size reporting): no-ops

struct structure_tree_node {};

struct structure_tree
{
  static structure_tree_node* add_child(structure_tree_node*, const std::string&, const std::string&) { return nullptr; }
  static void add_size(structure_tree_node*, std::uint64_t) {}
};

//------------------------------------------------------------------------------
// raw POD members

template<class T>
inline std::uint64_t
write_member(const T& t, std::ostream& out, structure_tree_node* = nullptr, const std::string& = "")
{
  out.write(reinterpret_cast<const char*>(&t), sizeof(T));
  return sizeof(T);
}

template<class T>
inline void
read_member(T& t, std::istream& in)
{
  in.read(reinterpret_cast<char*>(&t), sizeof(T));
}

//------------------------------------------------------------------------------
// int_vector<8>, int_vector<64>

template<std::uint8_t t_width> struct int_vector_trait;
template<> struct int_vector_trait<8>  { typedef std::uint8_t  value_type; };
template<> struct int_vector_trait<16> { typedef std::uint16_t value_type; };
template<> struct int_vector_trait<32> { typedef std::uint32_t value_type; };
template<> struct int_vector_trait<64> { typedef std::uint64_t value_type; };

template<std::uint8_t t_width>
class int_vector
{
public:
  typedef typename int_vector_trait<t_width>::value_type value_type;
  typedef std::uint64_t                                  size_type;
  typedef typename std::vector<value_type>::iterator       iterator;
  typedef typename std::vector<value_type>::const_iterator const_iterator;

  int_vector() {}
  int_vector(size_type n, value_type value = 0) : m_data(n, value) {}

  template<class T>
  int_vector(std::initializer_list<T> il)
  {
    m_data.reserve(il.size());
    for(auto x : il) { m_data.push_back(static_cast<value_type>(x)); }
  }

  size_type size() const { return m_data.size(); }
  bool empty() const { return m_data.empty(); }
  void resize(size_type n) { m_data.resize(n); }

  value_type& operator[](size_type i) { return m_data[i]; }
  const value_type& operator[](size_type i) const { return m_data[i]; }

  iterator begin() { return m_data.begin(); }
  iterator end() { return m_data.end(); }
  const_iterator begin() const { return m_data.begin(); }
  const_iterator end() const { return m_data.end(); }

  value_type* data() { return m_data.data(); }
  const value_type* data() const { return m_data.data(); }

  void swap(int_vector& another) { m_data.swap(another.m_data); }

  // u64 size in bits, then the data padded with zeros to a multiple of 8 bytes (utils.h:374-407).
  size_type serialize(std::ostream& out, structure_tree_node* = nullptr, const std::string& = "") const
  {
    size_type bit_size = m_data.size() * t_width;
    size_type written = write_member(bit_size, out);
    size_type bytes = m_data.size() * sizeof(value_type);
    out.write(reinterpret_cast<const char*>(m_data.data()), bytes);
    written += bytes;
    static const char zeros[8] = {};
    if(bytes % 8 != 0) { out.write(zeros, 8 - bytes % 8); written += 8 - bytes % 8; }
    return written;
  }

  void load(std::istream& in)
  {
    size_type bit_size = 0;
    read_member(bit_size, in);
    m_data.assign(bit_size / t_width, 0);
    size_type bytes = m_data.size() * sizeof(value_type);
    in.read(reinterpret_cast<char*>(m_data.data()), bytes);
    if(bytes % 8 != 0) { char pad[8]; in.read(pad, 8 - bytes % 8); }
  }

private:
  std::vector<value_type> m_data;
};

//------------------------------------------------------------------------------
// int_vector_buffer<8>: a file in int_vector<8> format with buffered sequential access.

template<std::uint8_t t_width>
class int_vector_buffer
{
public:
  typedef typename int_vector_trait<t_width>::value_type value_type;
  typedef std::uint64_t                                  size_type;

  int_vector_buffer() { this->init(); }

  explicit int_vector_buffer(const std::string& filename, std::ios::openmode mode = std::ios::in)
  {
    this->init();
    m_filename = filename;
    if(mode & std::ios::out)
    {
      m_file = std::fopen(filename.c_str(), "wb");
      if(m_file == nullptr) { throw std::runtime_error("int_vector_buffer: cannot open " + filename); }
      m_writing = true;
      size_type bit_size = 0;
      std::fwrite(&bit_size, sizeof(bit_size), 1, m_file);
      m_buffer.resize(BUFFER_ELEMENTS);
    }
    else
    {
      m_file = std::fopen(filename.c_str(), "rb");
      if(m_file == nullptr) { throw std::runtime_error("int_vector_buffer: cannot open " + filename); }
      size_type bit_size = 0;
      if(std::fread(&bit_size, sizeof(bit_size), 1, m_file) != 1) { bit_size = 0; }
      m_size = bit_size / t_width;
      m_buffer.resize(BUFFER_ELEMENTS);
    }
  }

  int_vector_buffer(const int_vector_buffer&) = delete;
  int_vector_buffer& operator=(const int_vector_buffer&) = delete;

  int_vector_buffer(int_vector_buffer&& another) { this->init(); this->swap(another); }
  int_vector_buffer& operator=(int_vector_buffer&& another)
  {
    if(this != &another) { this->close(); this->swap(another); }
    return *this;
  }

  ~int_vector_buffer() { this->close(); }

  size_type size() const { return m_size; }

  // Read access (read mode). Returns by value through a proxy-free fast path.
  inline value_type operator[](size_type i)
  {
    if(i < m_begin || i >= m_end) { this->fill(i); }
    return m_buffer[i - m_begin];
  }

  inline void push_back(value_type value)
  {
    m_buffer[m_end - m_begin] = value;
    m_end++; m_size++;
    if(m_end - m_begin >= BUFFER_ELEMENTS) { this->flush(); }
  }

  void close()
  {
    if(m_file == nullptr) { return; }
    if(m_writing)
    {
      this->flush();
      size_type bytes = m_size * sizeof(value_type);
      static const char zeros[8] = {};
      if(bytes % 8 != 0) { std::fwrite(zeros, 1, 8 - bytes % 8, m_file); }
      size_type bit_size = m_size * t_width;
      std::fseek(m_file, 0, SEEK_SET);
      std::fwrite(&bit_size, sizeof(bit_size), 1, m_file);
    }
    std::fclose(m_file);
    m_buffer = std::vector<value_type>();
    this->init();
  }

  void swap(int_vector_buffer& another)
  {
    std::swap(m_file, another.m_file);
    std::swap(m_filename, another.m_filename);
    std::swap(m_writing, another.m_writing);
    std::swap(m_size, another.m_size);
    std::swap(m_begin, another.m_begin);
    std::swap(m_end, another.m_end);
    m_buffer.swap(another.m_buffer);
  }

private:
  const static size_type BUFFER_ELEMENTS = 1048576;

  void init()
  {
    m_file = nullptr; m_writing = false; m_size = 0; m_begin = 0; m_end = 0;
  }

  void fill(size_type i)
  {
    m_begin = i - i % BUFFER_ELEMENTS;
    size_type limit = std::min(m_size, m_begin + BUFFER_ELEMENTS);
    std::fseek(m_file, sizeof(size_type) + m_begin * sizeof(value_type), SEEK_SET);
    size_type got = std::fread(m_buffer.data(), sizeof(value_type), limit - m_begin, m_file);
    m_end = m_begin + got;
    if(i >= m_end) { m_buffer[i - m_begin] = 0; m_end = i + 1; }  // Out of range: behave like zeros.
  }

  void flush()
  {
    if(m_end > m_begin) { std::fwrite(m_buffer.data(), sizeof(value_type), m_end - m_begin, m_file); }
    m_begin = m_end;
  }

  std::FILE*              m_file;
  std::string             m_filename;
  bool                    m_writing;
  size_type               m_size;
  size_type               m_begin, m_end; // Buffered range.
  std::vector<value_type> m_buffer;
};

//------------------------------------------------------------------------------
// Elias-Fano sparse bitvector

class sd_vector_builder
{
public:
  typedef std::uint64_t size_type;

  sd_vector_builder() : m_size(0), m_capacity(0), m_items(0), m_tail(0), m_wl(0), m_high_bits(0) {}

  sd_vector_builder(size_type n, size_type m) :
    m_size(n), m_capacity(m), m_items(0), m_tail(0)
  {
    if(m > n) { throw std::runtime_error("sd_vector_builder: requested capacity is larger than vector size."); }
    std::uint32_t logm = bits::hi(m) + 1, logn = bits::hi(n) + 1;
    if(logm == logn) { logm--; }
    m_wl = logn - logm;
    m_low.assign((m * m_wl + 63) / 64 + 1, 0);
    m_high.assign((m + (size_type(1) << logm) + 63) / 64 + 1, 0);
    m_high_bits = m + (size_type(1) << logm);
  }

  size_type size() const { return m_size; }
  size_type capacity() const { return m_capacity; }
  size_type tail() const { return m_tail; }
  size_type items() const { return m_items; }

  inline void set(size_type i)
  {
    if(m_items >= m_capacity || i < m_tail || i >= m_size)
    {
      throw std::runtime_error("sd_vector_builder: invalid set()");
    }
    if(m_wl > 0)
    {
      size_type low = i & bits::lo_set[m_wl];
      size_type bit = m_items * m_wl;
      m_low[bit / 64] |= low << (bit % 64);
      if(bit % 64 + m_wl > 64) { m_low[bit / 64 + 1] |= low >> (64 - bit % 64); }
    }
    size_type high_pos = (i >> m_wl) + m_items;
    m_high[high_pos / 64] |= size_type(1) << (high_pos % 64);
    m_items++; m_tail = i + 1;
  }

private:
  friend class sd_vector_impl;

  size_type     m_size, m_capacity, m_items, m_tail;
  std::uint32_t m_wl;
  size_type     m_high_bits;
  std::vector<std::uint64_t> m_low, m_high;
};

class sd_vector_impl
{
public:
  typedef std::uint64_t size_type;

  sd_vector_impl() : m_size(0), m_ones(0), m_wl(0), m_high_bits(0) {}

  explicit sd_vector_impl(sd_vector_builder& builder)
  {
    if(builder.items() != builder.capacity())
    {
      throw std::runtime_error("sd_vector: builder is not full.");
    }
    m_size = builder.m_size; m_ones = builder.m_items; m_wl = builder.m_wl;
    m_high_bits = builder.m_high_bits;
    m_low.swap(builder.m_low); m_high.swap(builder.m_high);
    builder = sd_vector_builder();
    this->buildSelect();
  }

  template<class Iterator>
  sd_vector_impl(Iterator begin, Iterator end)
  {
    size_type m = std::distance(begin, end);
    size_type n = (m > 0 ? *(end - 1) + 1 : 0);
    sd_vector_builder builder(n, m);
    for(Iterator iter = begin; iter != end; ++iter) { builder.set(*iter); }
    *this = sd_vector_impl(builder);
  }

  size_type size() const { return m_size; }
  size_type ones() const { return m_ones; }

  void swap(sd_vector_impl& another)
  {
    std::swap(m_size, another.m_size); std::swap(m_ones, another.m_ones);
    std::swap(m_wl, another.m_wl); std::swap(m_high_bits, another.m_high_bits);
    m_low.swap(another.m_low); m_high.swap(another.m_high);
    m_sel1.swap(another.m_sel1); m_sel0.swap(another.m_sel0);
  }

  size_type serialize(std::ostream& out, structure_tree_node* = nullptr, const std::string& = "") const
  {
    size_type written = 0;
    written += write_member(m_size, out);
    written += write_member(m_ones, out);
    std::uint8_t wl = m_wl;
    written += write_member(wl, out);
    size_type low_words = (m_ones * m_wl + 63) / 64, high_words = (m_high_bits + 63) / 64;
    out.write(reinterpret_cast<const char*>(m_low.data()), low_words * 8);
    out.write(reinterpret_cast<const char*>(m_high.data()), high_words * 8);
    written += (low_words + high_words) * 8;
    return written;
  }

  void load(std::istream& in)
  {
    read_member(m_size, in);
    read_member(m_ones, in);
    std::uint8_t wl = 0; read_member(wl, in); m_wl = wl;
    std::uint32_t logm = bits::hi(m_ones) + 1, logn = bits::hi(m_size) + 1;
    if(logm == logn) { logm--; }
    m_high_bits = m_ones + (size_type(1) << logm);
    size_type low_words = (m_ones * m_wl + 63) / 64, high_words = (m_high_bits + 63) / 64;
    m_low.assign(low_words + 1, 0); m_high.assign(high_words + 1, 0);
    in.read(reinterpret_cast<char*>(m_low.data()), low_words * 8);
    in.read(reinterpret_cast<char*>(m_high.data()), high_words * 8);
    this->buildSelect();
  }

  inline size_type low(size_type k) const
  {
    if(m_wl == 0) { return 0; }
    size_type bit = k * m_wl;
    size_type res = m_low[bit / 64] >> (bit % 64);
    if(bit % 64 + m_wl > 64) { res |= m_low[bit / 64 + 1] << (64 - bit % 64); }
    return res & bits::lo_set[m_wl];
  }

  inline bool highBit(size_type i) const { return (m_high[i / 64] >> (i % 64)) & 1; }

  // Position of the k-th (1-based) one / zero in the high bitvector.
  inline size_type highSelect1(size_type k) const
  {
    size_type pos = m_sel1[(k - 1) / SAMPLE];
    size_type skip = (k - 1) % SAMPLE + 1;  // Find the skip-th one at or after pos.
    size_type word = pos / 64;
    std::uint64_t w = m_high[word] & ~bits::lo_set[pos % 64];
    while(true)
    {
      size_type c = bits::cnt(w);
      if(c >= skip) { return word * 64 + bits::sel(w, skip); }
      skip -= c; word++; w = m_high[word];
    }
  }

  inline size_type highSelect0(size_type k) const
  {
    size_type pos = m_sel0[(k - 1) / SAMPLE];
    size_type skip = (k - 1) % SAMPLE + 1;
    size_type word = pos / 64;
    std::uint64_t w = ~m_high[word] & ~bits::lo_set[pos % 64];
    while(true)
    {
      size_type c = bits::cnt(w);
      if(c >= skip) { return word * 64 + bits::sel(w, skip); }
      skip -= c; word++; w = ~m_high[word];
    }
  }

  // Number of ones in [0, i).
  inline size_type rank1(size_type i) const
  {
    if(i >= m_size) { return m_ones; }
    if(m_ones == 0) { return 0; }
    size_type high_val = i >> m_wl;
    size_type sel_high = this->highSelect0(high_val + 1);
    size_type rank_low = sel_high - high_val; // Ones with high part <= high_val.
    size_type val_low = i & bits::lo_set[m_wl];
    while(rank_low > 0 && sel_high > 0 && this->highBit(sel_high - 1) && this->low(rank_low - 1) >= val_low)
    {
      sel_high--; rank_low--;
    }
    return rank_low;
  }

  // Position of the k-th (1-based) one.
  inline size_type select1(size_type k) const
  {
    return ((this->highSelect1(k) - (k - 1)) << m_wl) | this->low(k - 1);
  }

  // Position of the k-th (1-based) zero.
  size_type select0(size_type k) const
  {
    // Find the number of ones before the k-th zero: largest j with select1(j) - (j - 1) < k.
    size_type lo = 0, hi = m_ones;
    while(lo < hi)
    {
      size_type mid = lo + (hi - lo + 1) / 2;
      if(this->select1(mid) - (mid - 1) < k) { lo = mid; } else { hi = mid - 1; }
    }
    return k - 1 + lo;
  }

  inline bool access(size_type i) const
  {
    return (this->rank1(i + 1) - this->rank1(i)) != 0;
  }

private:
  const static size_type SAMPLE = 256;

  void buildSelect()
  {
    m_sel1.clear(); m_sel0.clear();
    size_type ones = 0, zeros = 0;
    size_type words = (m_high_bits + 63) / 64;
    for(size_type w = 0; w < words; w++)
    {
      std::uint64_t word = m_high[w];
      size_type limit = std::min(size_type(64), m_high_bits - w * 64);
      std::uint64_t valid = bits::lo_set[limit];
      std::uint64_t one_bits = word & valid, zero_bits = ~word & valid;
      size_type c1 = bits::cnt(one_bits), c0 = bits::cnt(zero_bits);
      // Sample k = j * SAMPLE + 1 (1-based) for j >= 0.
      while(m_sel1.size() * SAMPLE + 1 <= ones + c1)
      {
        m_sel1.push_back(w * 64 + bits::sel(one_bits, m_sel1.size() * SAMPLE + 1 - ones));
      }
      while(m_sel0.size() * SAMPLE + 1 <= zeros + c0)
      {
        m_sel0.push_back(w * 64 + bits::sel(zero_bits, m_sel0.size() * SAMPLE + 1 - zeros));
      }
      ones += c1; zeros += c0;
    }
    // Sentinels: make sure word scans cannot run past the end (extra zero word is allocated).
    if(m_high.size() < words + 1) { m_high.resize(words + 1, 0); }
  }

  size_type     m_size, m_ones;
  std::uint32_t m_wl;
  size_type     m_high_bits;
  std::vector<std::uint64_t> m_low, m_high;
  std::vector<size_type>     m_sel1, m_sel0;
};

template<class t_vector> class sd_rank_support;
template<class t_vector> class sd_select1_support;
template<class t_vector> class sd_select0_support;

struct sd_default_tag {};

template<class t_tag = sd_default_tag>
class sd_vector : public sd_vector_impl
{
public:
  typedef std::uint64_t size_type;
  typedef sd_rank_support<sd_vector>    rank_1_type;
  typedef sd_select1_support<sd_vector> select_1_type;
  typedef sd_select0_support<sd_vector> select_0_type;

  sd_vector() {}
  explicit sd_vector(sd_vector_builder& builder) : sd_vector_impl(builder) {}
  template<class Iterator>
  sd_vector(Iterator begin, Iterator end) : sd_vector_impl(begin, end) {}

  void swap(sd_vector& another) { sd_vector_impl::swap(another); }
  inline bool operator[](size_type i) const { return this->access(i); }
};

template<class t_vector>
class sd_rank_support
{
public:
  typedef std::uint64_t size_type;
  sd_rank_support() : m_v(nullptr) {}
  explicit sd_rank_support(const t_vector* v) : m_v(v) {}
  inline size_type operator()(size_type i) const { return m_v->rank1(i); }
  inline size_type rank(size_type i) const { return m_v->rank1(i); }
  void set_vector(const t_vector* v) { m_v = v; }
  void swap(sd_rank_support&) {}
  size_type serialize(std::ostream&, structure_tree_node* = nullptr, const std::string& = "") const { return 0; }
  void load(std::istream&, const t_vector* v = nullptr) { m_v = v; }
private:
  const t_vector* m_v;
};

template<class t_vector>
class sd_select1_support
{
public:
  typedef std::uint64_t size_type;
  sd_select1_support() : m_v(nullptr) {}
  explicit sd_select1_support(const t_vector* v) : m_v(v) {}
  inline size_type operator()(size_type k) const { return m_v->select1(k); }
  inline size_type select(size_type k) const { return m_v->select1(k); }
  void set_vector(const t_vector* v) { m_v = v; }
  void swap(sd_select1_support&) {}
  size_type serialize(std::ostream&, structure_tree_node* = nullptr, const std::string& = "") const { return 0; }
  void load(std::istream&, const t_vector* v = nullptr) { m_v = v; }
private:
  const t_vector* m_v;
};

template<class t_vector>
class sd_select0_support
{
public:
  typedef std::uint64_t size_type;
  sd_select0_support() : m_v(nullptr) {}
  explicit sd_select0_support(const t_vector* v) : m_v(v) {}
  inline size_type operator()(size_type k) const { return m_v->select0(k); }
  inline size_type select(size_type k) const { return m_v->select0(k); }
  void set_vector(const t_vector* v) { m_v = v; }
  void swap(sd_select0_support&) {}
  size_type serialize(std::ostream&, structure_tree_node* = nullptr, const std::string& = "") const { return 0; }
  void load(std::istream&, const t_vector* v = nullptr) { m_v = v; }
private:
  const t_vector* m_v;
};

//------------------------------------------------------------------------------
// util

namespace util
{

template<class T>
inline std::string class_name(const T&) { return typeid(T).name(); }

template<class T>
inline std::string to_string(const T& t) { std::ostringstream ss; ss << t; return ss.str(); }

inline std::uint64_t pid() { return static_cast<std::uint64_t>(::getpid()); }

inline std::uint64_t id()
{
  static std::atomic<std::uint64_t> counter(0);
  return counter++;
}

template<class T>
inline void clear(T& t) { T empty; t.swap(empty); }

template<class Support, class Vector>
inline void init_support(Support& support, const Vector* v) { Support temp(v); support = temp; support.set_vector(v); }

template<class Support, class Vector>
inline void swap_support(Support& a, Support& b, const Vector* va, const Vector* vb)
{
  a.swap(b); a.set_vector(va); b.set_vector(vb);
}

} // namespace util

//------------------------------------------------------------------------------
// size_in_bytes: the number of bytes serialize() writes.

class counting_streambuf : public std::streambuf
{
public:
  counting_streambuf() : m_count(0) {}
  std::uint64_t count() const { return m_count; }
protected:
  std::streamsize xsputn(const char*, std::streamsize n) override { m_count += n; return n; }
  int overflow(int c) override { m_count++; return c; }
private:
  std::uint64_t m_count;
};

template<class T>
inline std::uint64_t size_in_bytes(const T& t)
{
  counting_streambuf buffer;
  std::ostream out(&buffer);
  return t.serialize(out);
}

//------------------------------------------------------------------------------
// Names that must exist for utils.h:413-425 (directConstruct is never instantiated).

inline std::string ram_file_name(const std::string& name) { return "@" + name; }

template<class T>
inline bool store_to_file(const T& t, const std::string& filename)
{
  std::ofstream out(filename.c_str(), std::ios_base::binary);
  if(!out) { return false; }
  t.serialize(out);
  return true;
}

namespace ram_fs
{
inline int remove(const std::string& filename) { return std::remove(filename.c_str()); }
}

} // namespace sdsl

#endif // BWTM_ORACLE_SDSL_SHIM_WAVELET_TREES_HPP
