// Range-wise merge of sorted runs of rank-array values, shared by the single-GPU search in batches (bwtm_merge.cu) and
// the batched distributed merge (bwtm_dist.cu).
//
// The reference keeps its rank array as sorted, compressed runs in merge buffers and files and streams them back
// through a k-way merge while it interleaves (fmi.cpp:220-257, support.h:576-638, bwt.cpp:152-213). Here T sorted runs of
// plain values stay in HBM; the A positions are cut into ranges that hold about 2^30 merged positions each, and range
// by range the pieces of the T runs are gathered, merged pairwise and handed to the interleave and the encoder at
// once. Only two range-sized buffers are needed besides the runs.
#pragma once

#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "bwtm_merge.cuh"

namespace bwtm
{

// Range r = 0 .. ranges - 1 covers the A positions [splitters[r], splitters[r + 1]):
//   splitters[0] = x_lo, splitters[ranges] = x_hi, and for 0 < r < ranges the smallest x in [x_lo, x_hi] whose merged
//   position x + base + #{keys < x} reaches first_target + r * step (`base` = keys that precede all the runs);
//   bounds[r * T + k] = #{keys of run k below splitters[r]}.
template<class KeyT>
__global__ void batch_splitters(const KeyT* __restrict__ keys, const unsigned long long* __restrict__ run_offsets, int T,
                                unsigned long long x_lo, unsigned long long x_hi, unsigned long long base,
                                unsigned long long first_target, unsigned long long step, unsigned long long ranges,
                                unsigned long long* __restrict__ splitters, unsigned long long* __restrict__ bounds)
{
  unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(r > ranges) { return; }
  auto below = [&](unsigned long long x, int k) -> unsigned long long
  {
    unsigned long long lo = run_offsets[k], hi = run_offsets[k + 1];
    while(lo < hi)
    {
      unsigned long long mid = lo + (hi - lo) / 2;
      if((unsigned long long)keys[mid] < x) { lo = mid + 1; } else { hi = mid; }
    }
    return lo - run_offsets[k];
  };
  unsigned long long x = x_lo;
  if(r == ranges) { x = x_hi; }
  else if(r > 0)
  {
    unsigned long long target = first_target + r * step, lo = x_lo, hi = x_hi;
    while(lo < hi)
    {
      unsigned long long mid = lo + (hi - lo) / 2, placed = mid + base;
      for(int k = 0; k < T; k++) { placed += below(mid, k); }
      if(placed < target) { lo = mid + 1; } else { hi = mid; }
    }
    x = lo;
  }
  splitters[r] = x;
  for(int k = 0; k < T; k++) { bounds[r * T + k] = below(x, k); }
}

// Merges the sorted pieces [offsets[k], offsets[k + 1]) of `src` pairwise, ping-ponging between two buffers.
template<class KeyT>
static int merge_pieces(KeyT* src, KeyT* dst, std::vector<uint64_t> offsets, cudaStream_t stream, DeviceBuffer& temp, KeyT** result)
{
  while(offsets.size() > 2)
  {
    std::vector<uint64_t> next; next.push_back(0);
    for(size_t k = 0; k + 1 < offsets.size(); k += 2)
    {
      uint64_t begin = offsets[k], middle = offsets[k + 1], end = (k + 2 < offsets.size() ? offsets[k + 2] : offsets[k + 1]);
      if(end == middle || middle == begin)
      {
        if(end > begin) { BWTM_CUDA(cudaMemcpyAsync(dst + begin, src + begin, (end - begin) * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream)); }
      }
      else
      {
        size_t bytes = 0;
        BWTM_CUDA(cub::DeviceMerge::MergeKeys(nullptr, bytes, src + begin, (int)(middle - begin), src + middle, (int)(end - middle), dst + begin, ::cuda::std::less<>{}, stream));
        if(bytes > temp.bytes) { BWTM_CUDA(cudaStreamSynchronize(stream)); BWTM_TRY(temp.allocate(bytes)); }
        BWTM_CUDA(cub::DeviceMerge::MergeKeys(temp.ptr, bytes, src + begin, (int)(middle - begin), src + middle, (int)(end - middle), dst + begin, ::cuda::std::less<>{}, stream));
        count_launch(2);
      }
      next.push_back(end);
    }
    offsets.swap(next);
    std::swap(src, dst);
  }
  *result = src;
  return BWTM_OK;
}

// The T sorted runs [run_offsets[k], run_offsets[k + 1]) of `runs` hold all rank-array values in [x_lo, x_hi); `base`
// values precede them (smaller A positions), so the runs' values belong to b's positions base, base + 1, ... and the
// merged positions [begin, end). Interleaves and encodes that part of the merged BWT range by range, continuing the
// writer state in d_control; `finish` flushes the pending run after the last range.
template<class KeyT>
static int merge_ranges(const bwtm_index* a, const bwtm_index* b, const KeyT* runs, const std::vector<unsigned long long>& run_offsets,
                        uint64_t x_lo, uint64_t x_hi, uint64_t base, uint64_t begin, uint64_t end,
                        const bwtm_merge_options* options, OutputBuffer* out, EncodeControl* d_control, bool finish,
                        float* merge_ms, float* interleave_ms, float* encode_ms, unsigned long long* d_distinct, cudaStream_t stream)
{
  const int T = (int)run_offsets.size() - 1;
  const uint64_t total = end - begin;
  EventTimer timer(stream);
  timer.start();
  const uint64_t step = clamp_slab(options->slab_symbols, std::max<uint64_t>(total, 1));
  const uint64_t ranges = std::max<uint64_t>(1, div_up(total, step));
  DeviceBuffer d_offsets, d_splitters, d_bounds;
  BWTM_TRY(d_offsets.allocate((T + 1) * sizeof(unsigned long long)));
  BWTM_TRY(d_splitters.allocate((ranges + 1) * sizeof(unsigned long long)));
  BWTM_TRY(d_bounds.allocate((ranges + 1) * T * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_offsets.ptr, run_offsets.data(), (T + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  batch_splitters<KeyT><<<(unsigned)div_up(ranges + 1, 64), 64, 0, stream>>>(runs, d_offsets.as<unsigned long long>(), T, x_lo, x_hi, base, begin, step, ranges,
                                                                             d_splitters.as<unsigned long long>(), d_bounds.as<unsigned long long>());
  BWTM_LAUNCH_CHECK();
  std::vector<unsigned long long> splitters(ranges + 1), bounds((ranges + 1) * T);
  BWTM_CUDA(cudaMemcpyAsync(splitters.data(), d_splitters.ptr, (ranges + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds.ptr, (ranges + 1) * T * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t largest = 0;
  std::vector<uint64_t> keys_before(ranges + 1, 0);
  for(uint64_t r = 0; r <= ranges; r++)
  {
    for(int k = 0; k < T; k++) { keys_before[r] += bounds[r * T + k]; }
    if(r > 0) { largest = std::max(largest, keys_before[r] - keys_before[r - 1]); }
  }
  DeviceBuffer gathered, merged_keys, merge_temp;
  BWTM_TRY(gathered.allocate(std::max<uint64_t>(largest, 1) * sizeof(KeyT)));
  BWTM_TRY(merged_keys.allocate(std::max<uint64_t>(largest, 1) * sizeof(KeyT)));
  *merge_ms += timer.stop();

  for(uint64_t r = 0; r < ranges; r++)
  {
    const uint64_t count = keys_before[r + 1] - keys_before[r];
    const uint64_t range_begin = (r == 0 ? begin : splitters[r] + base + keys_before[r]);
    const uint64_t range_end = (r + 1 == ranges ? end : splitters[r + 1] + base + keys_before[r + 1]);
    const bool last = (r + 1 == ranges);
    if(range_end == range_begin && !(last && finish)) { continue; }
    KeyT* range_keys = gathered.as<KeyT>();
    timer.start();
    std::vector<uint64_t> piece_offsets(1, 0);
    for(int k = 0; k < T; k++)
    {
      uint64_t from = run_offsets[k] + bounds[r * T + k], piece = bounds[(r + 1) * T + k] - bounds[r * T + k];
      if(piece > 0) { BWTM_CUDA(cudaMemcpyAsync(gathered.as<KeyT>() + piece_offsets.back(), runs + from, piece * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream)); }
      piece_offsets.push_back(piece_offsets.back() + piece);
    }
    BWTM_TRY(merge_pieces<KeyT>(gathered.as<KeyT>(), merged_keys.as<KeyT>(), piece_offsets, stream, merge_temp, &range_keys));
    *merge_ms += timer.stop();
    BWTM_TRY(interleave_range<KeyT>(a, b, range_keys, base + keys_before[r], count, range_begin, range_end, options->slab_symbols, out, d_control,
                                    last && finish, interleave_ms, encode_ms, stream, d_distinct, nullptr));
  }
  return BWTM_OK;
}

} // namespace bwtm
