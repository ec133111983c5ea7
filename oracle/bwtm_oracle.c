/*
  TEST INFRASTRUCTURE -- not part of the product.  See bwtm_oracle.h.

  Plain-C restatement of the bwt-merge rank-array path.  The sparse bitvectors of the
  reference (SDSL sd_vector; not vendored) are replaced by plain sorted
  arrays with the same rank/select semantics:
    block_rank(i)        = number of blocks whose last position is < i   (bwt.cpp:324)
    block_select(k) + 1  = first position of block k                     (bwt.cpp:327)
    samples[c].sum(k)    = number of c's in blocks 0..k-1                (support.h:338-343)
*/
#include "bwtm_oracle.h"

#include <stdlib.h>
#include <string.h>

/*----------------------------------------------------------------------------*/
/* growable byte array: the ByteArray concept (size(), push_back(), operator[]) */

void orc_bytes_init(orc_bytes* a) { a->data = NULL; a->size = 0; a->capacity = 0; }
void orc_bytes_free(orc_bytes* a) { free(a->data); orc_bytes_init(a); }

void orc_bytes_push(orc_bytes* a, uint8_t v)
{
  if(a->size >= a->capacity)
  {
    a->capacity = (a->capacity == 0 ? 1024 : 2 * a->capacity);
    a->data = (uint8_t*)realloc(a->data, a->capacity);
  }
  a->data[a->size++] = v;
}

void orc_free(void* p) { free(p); }

/*----------------------------------------------------------------------------*/
/* utils.h:146-151 with sdsl::bits::hi(0) == 0, so bit_length(0) == 1. */

static uint64_t bit_length(uint64_t v)
{
  uint64_t hi = 0;
  while(v > 1) { v >>= 1; hi++; }
  return hi + 1;
}

/* support.h:172-184: little-endian base-128, bit 7 = continues. */
uint64_t orc_bytecode_read(const uint8_t* array, uint64_t* i)
{
  uint64_t offset = 0;
  uint64_t res = array[*i] & 0x7F;
  while(array[*i] & 0x80)
  {
    (*i)++; offset += 7;
    res += ((uint64_t)(array[*i] & 0x7F)) << offset;
  }
  (*i)++;
  return res;
}

/* support.h:203-212 */
void orc_bytecode_write(orc_bytes* array, uint64_t value)
{
  while(value > 0x7F)
  {
    orc_bytes_push(array, (uint8_t)((value & 0x7F) | 0x80));
    value >>= 7;
  }
  orc_bytes_push(array, (uint8_t)value);
}

/* support.h:236-250: code = comp + 6 * (length - 1); length 42 is followed by a ByteCode. */
void orc_run_read(const uint8_t* array, uint64_t* i, uint8_t* comp, uint64_t* length)
{
  uint8_t code = array[*i]; (*i)++;
  *comp = code % ORC_SIGMA;
  *length = code / ORC_SIGMA + 1;
  if(*length >= ORC_MAX_RUN) { *length += orc_bytecode_read(array, i); }
}

/* support.h:256-282: a run never continues past a 64-byte block boundary. */
void orc_run_write(orc_bytes* array, uint8_t comp, uint64_t length)
{
  while(length > 0)
  {
    if(length < ORC_MAX_RUN)
    {
      orc_bytes_push(array, (uint8_t)(comp + ORC_SIGMA * (length - 1)));
      return;
    }

    uint64_t bytes_remaining = ORC_BLOCK_SIZE - (array->size % ORC_BLOCK_SIZE);
    uint64_t basic_length = (bytes_remaining > 1 ? ORC_MAX_RUN : ORC_MAX_RUN - 1);
    orc_bytes_push(array, (uint8_t)(comp + ORC_SIGMA * (basic_length - 1))); length -= basic_length;
    bytes_remaining--;

    if(bytes_remaining > 0)
    {
      uint64_t extension_length = length;
      if(bit_length(length) > 7 * bytes_remaining)
      {
        extension_length = (((uint64_t)1) << (7 * bytes_remaining)) - 1; /* lo_set[7 * rem], 7 * rem < 64 here */
      }
      orc_bytecode_write(array, extension_length); length -= extension_length;
    }
  }
}

/*----------------------------------------------------------------------------*/
/* utils.h:121-142 */

typedef struct
{
  uint64_t value, length;
  uint64_t run_value, run_length;
} run_buffer;

static void rb_init(run_buffer* rb) { rb->value = 0; rb->length = 0; rb->run_value = 0; rb->run_length = 0; }
static void rb_flush(run_buffer* rb) { rb->run_value = rb->value; rb->run_length = rb->length; }

static int rb_add(run_buffer* rb, uint64_t v, uint64_t n)
{
  if(v == rb->value) { rb->length += n; return 0; }
  rb_flush(rb);
  rb->value = v; rb->length = n;
  return (rb->run_length > 0);
}

/*----------------------------------------------------------------------------*/
/* BWT::setHeader (bwt.cpp:468-474) + BWT::build (bwt.cpp:476-512) */

static void bwt_build(orc_bwt* bwt)
{
  bwt->blocks = (bwt->bytes + ORC_BLOCK_SIZE - 1) / ORC_BLOCK_SIZE;
  bwt->block_end = (uint64_t*)calloc(bwt->blocks + 1, sizeof(uint64_t));
  for(int c = 0; c < ORC_SIGMA; c++)
  {
    bwt->cum[c] = (uint64_t*)calloc(bwt->blocks + 1, sizeof(uint64_t));
    bwt->counts[c] = 0;
  }

  uint64_t seq_pos = 0, rle_pos = 0, block = 0;
  uint64_t cumulative[ORC_SIGMA] = { 0, 0, 0, 0, 0, 0 };
  while(rle_pos < bwt->bytes)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length; cumulative[comp] += length;
    if(rle_pos >= bwt->bytes || rle_pos % ORC_BLOCK_SIZE == 0)
    {
      bwt->block_end[block] = seq_pos - 1;
      for(int c = 0; c < ORC_SIGMA; c++) { bwt->cum[c][block + 1] = cumulative[c]; }
      block++;
    }
  }

  bwt->size = seq_pos;
  for(int c = 0; c < ORC_SIGMA; c++) { bwt->counts[c] = cumulative[c]; }
  bwt->sequences = bwt->counts[0];
  bwt->C[0] = 0;
  for(int c = 0; c < ORC_SIGMA; c++) { bwt->C[c + 1] = bwt->C[c] + bwt->counts[c]; } /* support.cpp:90 */
}

orc_bwt* orc_bwt_from_rle(const uint8_t* rle, uint64_t bytes)
{
  orc_bwt* bwt = (orc_bwt*)calloc(1, sizeof(orc_bwt));
  bwt->rle = (uint8_t*)malloc(bytes + 16);
  memcpy(bwt->rle, rle, bytes);
  memset(bwt->rle + bytes, 0, 16);
  bwt->bytes = bytes;
  bwt_build(bwt);
  return bwt;
}

/* PlainData::read (formats.cpp:133-161) on comp values: RunBuffer -> Run::write. */
orc_bwt* orc_bwt_from_comps(const uint8_t* comps, uint64_t n)
{
  orc_bytes data; orc_bytes_init(&data);
  run_buffer rb; rb_init(&rb);
  for(uint64_t i = 0; i < n; i++)
  {
    if(rb_add(&rb, comps[i], 1)) { orc_run_write(&data, (uint8_t)rb.run_value, rb.run_length); }
  }
  rb_flush(&rb);
  orc_run_write(&data, (uint8_t)rb.run_value, rb.run_length);
  orc_bwt* bwt = orc_bwt_from_rle(data.data, data.size);
  orc_bytes_free(&data);
  return bwt;
}

void orc_bwt_free(orc_bwt* bwt)
{
  if(bwt == NULL) { return; }
  free(bwt->rle); free(bwt->block_end);
  for(int c = 0; c < ORC_SIGMA; c++) { free(bwt->cum[c]); }
  free(bwt);
}

uint64_t orc_bwt_bytes(const orc_bwt* bwt) { return bwt->bytes; }
uint64_t orc_bwt_size(const orc_bwt* bwt) { return bwt->size; }
uint64_t orc_bwt_sequences(const orc_bwt* bwt) { return bwt->sequences; }
uint64_t orc_bwt_blocks(const orc_bwt* bwt) { return bwt->blocks; }
const uint8_t* orc_bwt_rle(const orc_bwt* bwt) { return bwt->rle; }

void orc_bwt_counts(const orc_bwt* bwt, uint64_t* counts6)
{
  for(int c = 0; c < ORC_SIGMA; c++) { counts6[c] = bwt->counts[c]; }
}

void orc_bwt_samples(const orc_bwt* bwt, uint64_t* block_end_out, uint64_t* cum_out)
{
  for(uint64_t k = 0; k < bwt->blocks; k++)
  {
    block_end_out[k] = bwt->block_end[k];
    for(int c = 0; c < ORC_SIGMA; c++) { cum_out[c * bwt->blocks + k] = bwt->cum[c][k + 1]; }
  }
}

void orc_bwt_decode(const orc_bwt* bwt, uint8_t* comps_out)
{
  uint64_t rle_pos = 0, seq_pos = 0;
  while(rle_pos < bwt->bytes)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    memset(comps_out + seq_pos, comp, length);
    seq_pos += length;
  }
}

/* BWT::hash (bwt.cpp:538-549), FNV-1a constants utils.h:155-161. */
uint64_t orc_bwt_hash(const orc_bwt* bwt)
{
  uint64_t res = 0xcbf29ce484222325ULL;
  uint64_t rle_pos = 0;
  while(rle_pos < bwt->bytes)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    for(uint64_t i = 0; i < length; i++) { res = (res ^ comp) * 0x100000001b3ULL; }
  }
  return res;
}

/*----------------------------------------------------------------------------*/
/* Block lookup.  block_rank(i): ones in block_boundaries[0, i) = #blocks with block_end < i. */

static uint64_t block_rank(const orc_bwt* bwt, uint64_t i)
{
  uint64_t lo = 0, hi = bwt->blocks;
  while(lo < hi)
  {
    uint64_t mid = lo + (hi - lo) / 2;
    if(bwt->block_end[mid] < i) { lo = mid + 1; } else { hi = mid; }
  }
  return lo;
}

static uint64_t block_start(const orc_bwt* bwt, uint64_t block)
{
  return (block > 0 ? bwt->block_end[block - 1] + 1 : 0);
}

/* BWT::rank (bwt.cpp:318-341) */
uint64_t orc_rank(const orc_bwt* bwt, uint64_t i, uint8_t c)
{
  if(c >= ORC_SIGMA) { return 0; }
  if(i > bwt->size) { i = bwt->size; }

  uint64_t block = block_rank(bwt, i);
  uint64_t res = bwt->cum[c][block];
  uint64_t rle_pos = block * ORC_BLOCK_SIZE;
  uint64_t seq_pos = block_start(bwt, block);

  while(seq_pos < i)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    if(comp == c)
    {
      res += length;
      if(seq_pos > i) { res -= seq_pos - i; }
    }
  }
  return res;
}

/* BWT::ranks(i) (bwt.cpp:343-361).  results[0] is also filled here (the reference leaves it undefined). */
void orc_ranks(const orc_bwt* bwt, uint64_t i, uint64_t* results)
{
  if(i > bwt->size) { i = bwt->size; }

  uint64_t block = block_rank(bwt, i);
  for(int c = 0; c < ORC_SIGMA; c++) { results[c] = bwt->cum[c][block]; }
  uint64_t rle_pos = block * ORC_BLOCK_SIZE;
  uint64_t seq_pos = block_start(bwt, block);

  uint64_t prev = 0;
  while(seq_pos < i)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    results[comp] += length; prev = comp;
  }
  results[prev] -= seq_pos - i;
}

/* BWT::ranks(range) (bwt.cpp:363-403): (rank(sp), rank(ep + 1)) by a linear scan; entries of
   characters that do not occur in the range may be wrong, as in the reference (bwt.h:122-127). */
void orc_ranks_range(const orc_bwt* bwt, uint64_t sp, uint64_t ep, uint64_t* first, uint64_t* second)
{
  if(sp > bwt->size - 1) { sp = bwt->size - 1; }
  if(ep > bwt->size - 1) { ep = bwt->size - 1; }
  for(int c = 0; c < ORC_SIGMA; c++) { first[c] = 0; second[c] = 0; }

  uint64_t block = block_rank(bwt, sp);
  uint64_t rle_pos = block * ORC_BLOCK_SIZE;
  uint64_t seq_pos = block_start(bwt, block);

  uint8_t comp = 0; uint64_t length = 0;
  while(seq_pos < sp)
  {
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    first[comp] += length; second[comp] += length;
  }
  first[comp] -= seq_pos - sp;

  while(seq_pos <= ep)
  {
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    second[comp] += length;
  }
  second[comp] -= (seq_pos - 1) - ep;

  for(int c = 1; c < ORC_SIGMA; c++)
  {
    if(second[c] > first[c])
    {
      uint64_t temp = bwt->cum[c][block];
      first[c] += temp; second[c] += temp;
    }
  }
}

/* BWT::inverse_select (bwt.cpp:445-464): (rank(i, BWT[i]), BWT[i]). */
void orc_inverse_select(const orc_bwt* bwt, uint64_t i, uint64_t* rank_out, uint8_t* comp_out)
{
  *rank_out = 0; *comp_out = 0;
  if(i >= bwt->size) { return; }

  uint64_t block = block_rank(bwt, i);
  uint64_t rle_pos = block * ORC_BLOCK_SIZE;
  uint64_t seq_pos = block_start(bwt, block);

  uint64_t ranks[ORC_SIGMA] = { 0, 0, 0, 0, 0, 0 };
  uint8_t comp = 0; uint64_t length = 0;
  while(seq_pos <= i)
  {
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    ranks[comp] += length;
  }

  *rank_out = bwt->cum[comp][block] + ranks[comp] - (seq_pos - i);
  *comp_out = comp;
}

/* BWT::operator[] (bwt.cpp:429-443) */
uint8_t orc_access(const orc_bwt* bwt, uint64_t i)
{
  if(i >= bwt->size) { return 0; }
  uint64_t block = block_rank(bwt, i);
  uint64_t rle_pos = block * ORC_BLOCK_SIZE;
  uint64_t seq_pos = block_start(bwt, block);
  while(1)
  {
    uint8_t comp; uint64_t length;
    orc_run_read(bwt->rle, &rle_pos, &comp, &length);
    seq_pos += length;
    if(seq_pos > i) { return comp; }
  }
}

/* Range::empty (utils.h:80-83) */
static int range_empty(uint64_t first, uint64_t second) { return (first + 1 > second + 1); }

/* FMI::find (fmi.h:195-209) with charRange (utils.h:318-323) and LF(range, comp) (utils.h:350-355). */
void orc_find(const orc_bwt* bwt, const uint8_t* pattern, uint64_t length, uint64_t* sp, uint64_t* ep)
{
  if(length == 0) { *sp = 0; *ep = bwt->size - 1; return; }

  uint64_t end = length - 1;
  uint8_t c = pattern[end];
  uint64_t first = bwt->C[c], second = bwt->C[c + 1] - 1;
  while(!range_empty(first, second) && end != 0)
  {
    end--; c = pattern[end];
    uint64_t f = bwt->C[c] + orc_rank(bwt, first, c);
    uint64_t s = bwt->C[c] + orc_rank(bwt, second + 1, c) - 1;
    first = f; second = s;
  }
  *sp = first; *ep = second;
}

/* Range::length of the result (bwt_merge.cpp:253-254). */
uint64_t orc_count(const orc_bwt* bwt, const uint8_t* pattern, uint64_t length)
{
  uint64_t sp, ep;
  orc_find(bwt, pattern, length, &sp, &ep);
  return ep + 1 - sp;
}

/*----------------------------------------------------------------------------*/
/* buildRA (fmi.cpp:272-334) for one block of sequences [seq_first, seq_last]. */

typedef struct { uint64_t a_pos, sp, ep; } merge_position; /* fmi.cpp:261-270 */

typedef struct { merge_position* data; uint64_t size, capacity; } position_stack;

static void stack_push(position_stack* s, uint64_t a_pos, uint64_t sp, uint64_t ep)
{
  if(s->size >= s->capacity)
  {
    s->capacity = (s->capacity == 0 ? 1024 : 2 * s->capacity);
    s->data = (merge_position*)realloc(s->data, s->capacity * sizeof(merge_position));
  }
  s->data[s->size].a_pos = a_pos; s->data[s->size].sp = sp; s->data[s->size].ep = ep;
  s->size++;
}

typedef struct { orc_run* data; uint64_t size, capacity; } run_vector;

static void runs_push(run_vector* v, uint64_t pos, uint64_t len)
{
  if(v->size >= v->capacity)
  {
    v->capacity = (v->capacity == 0 ? 4096 : 2 * v->capacity);
    v->data = (orc_run*)realloc(v->data, v->capacity * sizeof(orc_run));
  }
  v->data[v->size].pos = pos; v->data[v->size].len = len;
  v->size++;
}

uint64_t orc_build_ra_dfs(const orc_bwt* a, const orc_bwt* b, uint64_t seq_first, uint64_t seq_last, orc_run** runs_out)
{
  position_stack positions = { NULL, 0, 0 };
  run_vector run_buf = { NULL, 0, 0 };
  uint64_t a_pos[ORC_SIGMA], b_sp[ORC_SIGMA], b_ep[ORC_SIGMA];

  stack_push(&positions, a->sequences, seq_first, seq_last);  /* fmi.cpp:286 */
  while(positions.size > 0)
  {
    merge_position curr = positions.data[--positions.size];
    uint64_t length = curr.ep + 1 - curr.sp;
    runs_push(&run_buf, curr.a_pos, length);                  /* fmi.cpp:290 */

    if(length == 1)                                           /* fmi.cpp:296-303 */
    {
      uint64_t rank; uint8_t comp;
      orc_inverse_select(b, curr.sp, &rank, &comp);
      if(comp != 0)
      {
        uint64_t next_b = rank + b->C[comp];                  /* utils.h:335-341 */
        uint64_t next_a = a->C[comp] + orc_rank(a, curr.a_pos, comp); /* utils.h:343-348 */
        stack_push(&positions, next_a, next_b, next_b);
      }
    }
    else if(length <= ORC_SHORT_RANGE)                        /* fmi.cpp:304-314 */
    {
      orc_ranks_range(b, curr.sp, curr.ep, b_sp, b_ep);
      for(int c = 1; c < ORC_SIGMA; c++)
      {
        uint64_t first = b_sp[c] + b->C[c], second = b_ep[c] + b->C[c] - 1; /* fmi.h:186-193 */
        if(!range_empty(first, second))
        {
          stack_push(&positions, a->C[c] + orc_rank(a, curr.a_pos, (uint8_t)c), first, second);
        }
      }
    }
    else                                                      /* fmi.cpp:315-322 */
    {
      orc_ranks(a, curr.a_pos, a_pos);
      orc_ranks(b, curr.sp, b_sp); orc_ranks(b, curr.ep + 1, b_ep);
      for(int c = 1; c < ORC_SIGMA; c++)
      {
        uint64_t sp = b_sp[c] + b->C[c], ep = b_ep[c] + b->C[c] - 1; /* fmi.h:174-181 */
        if(sp <= ep) { stack_push(&positions, a_pos[c] + a->C[c], sp, ep); }
      }
    }
  }

  free(positions.data);
  *runs_out = run_buf.data;
  return run_buf.size;
}

/*
  The same rank array computed the way the device computes it: one backward walk per
  sequence of B (finding 3).  Emits one A-position per suffix of B.
*/
uint64_t orc_build_ra_walk(const orc_bwt* a, const orc_bwt* b, uint64_t seq_first, uint64_t seq_last, uint64_t* out)
{
  uint64_t n = 0;
  for(uint64_t seq = seq_first; seq <= seq_last; seq++)
  {
    uint64_t b_pos = seq, a_pos = a->sequences;
    while(1)
    {
      out[n++] = a_pos;
      uint64_t rank; uint8_t comp;
      orc_inverse_select(b, b_pos, &rank, &comp);
      if(comp == 0) { break; }
      b_pos = rank + b->C[comp];
      a_pos = a->C[comp] + orc_rank(a, a_pos, comp);
    }
  }
  return n;
}

static int run_compare(const void* x, const void* y)
{
  const orc_run* a = (const orc_run*)x; const orc_run* b = (const orc_run*)y;
  if(a->pos != b->pos) { return (a->pos < b->pos ? -1 : 1); }
  if(a->len != b->len) { return (a->len < b->len ? -1 : 1); }
  return 0;
}

/* RLArray(std::vector&) (support.h:415-429): sort, then RunBuffer coalesces equal positions. In place. */
uint64_t orc_sort_compress(orc_run* runs, uint64_t n)
{
  if(n == 0) { return 0; }
  qsort(runs, n, sizeof(orc_run), run_compare);
  uint64_t out = 0;
  run_buffer rb; rb_init(&rb);
  for(uint64_t i = 0; i < n; i++)
  {
    if(rb_add(&rb, runs[i].pos, runs[i].len)) { runs[out].pos = rb.run_value; runs[out].len = rb.run_length; out++; }
  }
  rb_flush(&rb);
  runs[out].pos = rb.run_value; runs[out].len = rb.run_length; out++;
  return out;
}

/*----------------------------------------------------------------------------*/
/* mergeBWT (bwt.cpp:215-282) followed by BWT::BWT(a, b, ra) header/build (bwt.cpp:305-308). */

#define EMIT(rb, out, comp, len) \
  if(rb_add(&(rb), (comp), (len))) { orc_run_write(&(out), (uint8_t)(rb).run_value, (rb).run_length); }

orc_bwt* orc_interleave(const orc_bwt* a, const orc_bwt* b, const orc_run* ra, uint64_t ra_runs)
{
  orc_bytes result; orc_bytes_init(&result);
  run_buffer out_buffer; rb_init(&out_buffer);
  uint64_t a_rle_pos = 0, b_rle_pos = 0, a_seq_pos = 0;
  uint8_t a_comp = 0, b_comp = 0; uint64_t a_len = 0, b_len = 0;
  orc_run_read(a->rle, &a_rle_pos, &a_comp, &a_len);
  orc_run_read(b->rle, &b_rle_pos, &b_comp, &b_len);

  for(uint64_t i = 0; i < ra_runs; i++)
  {
    orc_run curr = ra[i];
    while(a_seq_pos < curr.pos)
    {
      uint64_t length = (curr.pos - a_seq_pos < a_len ? curr.pos - a_seq_pos : a_len);
      EMIT(out_buffer, result, a_comp, length);
      a_len -= length; a_seq_pos += length;
      if(a_len == 0 && a_rle_pos < a->bytes) { orc_run_read(a->rle, &a_rle_pos, &a_comp, &a_len); }
    }
    while(curr.len > 0)
    {
      uint64_t length = (curr.len < b_len ? curr.len : b_len);
      EMIT(out_buffer, result, b_comp, length);
      b_len -= length; curr.len -= length;
      if(b_len == 0 && b_rle_pos < b->bytes) { orc_run_read(b->rle, &b_rle_pos, &b_comp, &b_len); }
    }
  }

  while(a_len > 0)
  {
    EMIT(out_buffer, result, a_comp, a_len);
    if(a_rle_pos < a->bytes) { orc_run_read(a->rle, &a_rle_pos, &a_comp, &a_len); }
    else { a_len = 0; }
  }

  rb_flush(&out_buffer);
  orc_run_write(&result, (uint8_t)out_buffer.run_value, out_buffer.run_length);

  orc_bwt* merged = orc_bwt_from_rle(result.data, result.size);
  orc_bytes_free(&result);
  return merged;
}

/* FMI::FMI(a, b, parameters) (fmi.cpp:336-369) with one sequence block. */
orc_bwt* orc_merge(const orc_bwt* a, const orc_bwt* b, int use_dfs)
{
  orc_run* runs = NULL; uint64_t n = 0;
  if(use_dfs)
  {
    n = orc_build_ra_dfs(a, b, 0, b->sequences - 1, &runs);
  }
  else
  {
    uint64_t* positions = (uint64_t*)malloc((b->size + 1) * sizeof(uint64_t));
    n = orc_build_ra_walk(a, b, 0, b->sequences - 1, positions);
    runs = (orc_run*)malloc((n + 1) * sizeof(orc_run));
    for(uint64_t i = 0; i < n; i++) { runs[i].pos = positions[i]; runs[i].len = 1; }
    free(positions);
  }
  n = orc_sort_compress(runs, n);
  orc_bwt* merged = orc_interleave(a, b, runs, n);
  free(runs);
  return merged;
}

/*----------------------------------------------------------------------------*/
/*
  Input construction: the multi-string BWT of a read collection (paper/paper.tex:141-145):
  every read is terminated by its own endmarker $_i, $_i < $_j for i < j, all endmarkers are
  smaller than the bases.  Row i of the BWT is the suffix "$_i", so rows 0..reads-1 hold the
  last base of each read.  Suffixes are sorted by plain comparison (test sizes only).
*/

static const uint8_t*  sort_comps;
static const uint64_t* sort_starts;

typedef struct { uint32_t read; uint32_t offset; } suffix_type;

static int suffix_compare(const void* x, const void* y)
{
  const suffix_type* a = (const suffix_type*)x; const suffix_type* b = (const suffix_type*)y;
  const uint8_t* pa = sort_comps + sort_starts[a->read] + a->offset;
  const uint8_t* pb = sort_comps + sort_starts[b->read] + b->offset;
  uint64_t la = sort_starts[a->read + 1] - sort_starts[a->read] - a->offset;
  uint64_t lb = sort_starts[b->read + 1] - sort_starts[b->read] - b->offset;
  uint64_t l = (la < lb ? la : lb);
  int r = memcmp(pa, pb, l);
  if(r != 0) { return r; }
  if(la != lb) { return (la < lb ? -1 : 1); }  /* the shorter one hits its endmarker first */
  if(a->read != b->read) { return (a->read < b->read ? -1 : 1); }
  return 0;
}

int orc_build_bwt_from_reads(const uint8_t* comps, const uint64_t* read_starts, uint64_t reads, uint8_t* bwt_out)
{
  uint64_t total = read_starts[reads] + reads;
  suffix_type* suffixes = (suffix_type*)malloc(total * sizeof(suffix_type));
  if(suffixes == NULL) { return -1; }
  uint64_t n = 0;
  for(uint64_t r = 0; r < reads; r++)
  {
    uint64_t len = read_starts[r + 1] - read_starts[r];
    for(uint64_t p = 0; p <= len; p++) { suffixes[n].read = (uint32_t)r; suffixes[n].offset = (uint32_t)p; n++; }
  }
  sort_comps = comps; sort_starts = read_starts;
  qsort(suffixes, n, sizeof(suffix_type), suffix_compare);
  for(uint64_t i = 0; i < n; i++)
  {
    bwt_out[i] = (suffixes[i].offset > 0 ? comps[read_starts[suffixes[i].read] + suffixes[i].offset - 1] : 0);
  }
  free(suffixes);
  return 0;
}
