// Multi-GPU merge: one process per GPU, NCCL over NVLink.
//
//   search      : both indexes are replicated; rank r walks the sequences [m r / G, m (r + 1) / G) of B
//                 (the reference's own unit of parallelism, fmi.cpp:355) and sorts its RA values locally;
//   exchange    : G - 1 splitters on A positions are found by a distributed binary search so that every
//                 rank gets the same number of MERGED positions; one all-to-all (grouped ncclSend/ncclRecv)
//                 moves every RA value to the rank that owns its A-position range;
//   interleave  : rank r owns A positions [s_r, s_{r+1}) and the B symbols whose RA value lies there: a
//                 contiguous slice of the merged BWT. Symbols and maximal runs (K4, K3) are computed on all
//                 ranks at once; the byte writer (K5) needs the output offset modulo 64 and the pending run
//                 of the previous slice, so its 88-byte state travels down the ranks;
//   gather      : slices are broadcast into the complete run-length BWT on every rank, which rebuilds its
//                 replica of the rank structure (the next sequential merge needs it everywhere).
//
// There is no collective in the search itself. NCCL is loaded with dlopen so that single-GPU users do not
// need it.
//
// The two bulk exchanges (RA values, output slices) do not go through NCCL's staged send/recv: every rank
// owns two peer windows (cudaMalloc + CUDA IPC, opened once by all other ranks and kept in the communicator)
// and the senders store straight into the receiver's window over NVLink, one pass and no bounce buffer.
// NCCL carries the small control traffic (counts, writer state, barriers) and remains the fallback for
// the bulk data when IPC mappings are not available (BWTM_NCCL_EXCHANGE=1 forces it).
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <vector>

#include <nccl.h>

#include <cub/cub.cuh>

#include "bwtm_merge.cuh"
#include "bwtm_batches.cuh"

using namespace bwtm;

namespace
{

struct NcclApi
{
  void* handle = nullptr;
  decltype(&ncclGetUniqueId)    GetUniqueId = nullptr;
  decltype(&ncclCommInitRank)   CommInitRank = nullptr;
  decltype(&ncclCommDestroy)    CommDestroy = nullptr;
  decltype(&ncclAllGather)      AllGather = nullptr;
  decltype(&ncclAllReduce)      AllReduce = nullptr;
  decltype(&ncclBroadcast)      Broadcast = nullptr;
  decltype(&ncclSend)           Send = nullptr;
  decltype(&ncclRecv)           Recv = nullptr;
  decltype(&ncclGroupStart)     GroupStart = nullptr;
  decltype(&ncclGroupEnd)       GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};

NcclApi* nccl()
{
  static NcclApi api;
  static bool tried = false;
  if(tried) { return &api; }
  tried = true;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  for(const char* name : names)
  {
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if(api.handle != nullptr) { break; }
  }
  if(api.handle == nullptr) { return &api; }
#define BWTM_NCCL_SYM(field, symbol) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, symbol)); if(api.field == nullptr) { return &api; }
  BWTM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  BWTM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  BWTM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  BWTM_NCCL_SYM(AllGather, "ncclAllGather")
  BWTM_NCCL_SYM(AllReduce, "ncclAllReduce")
  BWTM_NCCL_SYM(Broadcast, "ncclBroadcast")
  BWTM_NCCL_SYM(Send, "ncclSend")
  BWTM_NCCL_SYM(Recv, "ncclRecv")
  BWTM_NCCL_SYM(GroupStart, "ncclGroupStart")
  BWTM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  BWTM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BWTM_NCCL_SYM
  api.ok = true;
  return &api;
}

int nccl_failed(ncclResult_t result, const char* what)
{
  set_error("NCCL error in %s: %s", what, nccl()->ok ? nccl()->GetErrorString(result) : "library not loaded");
  return BWTM_ERR_COMM;
}

#define BWTM_NCCL(call) do { ncclResult_t res__ = (call); if(res__ != ncclSuccess) { return nccl_failed(res__, #call); } } while(0)

} // namespace

// Device memory of this rank that every other rank of the communicator has mapped.
struct PeerWindow
{
  uint8_t* local = nullptr;
  uint64_t capacity = 0;            // identical on all ranks
  std::vector<uint8_t*> mapped;     // mapped[p]: rank p's window in this process (mapped[rank] == local)
};

struct bwtm_comm
{
  ncclComm_t comm;
  int        rank, world;
  int        peer_state;            // 0 not tried yet, 1 windows usable, -1 not available: NCCL moves the data
  PeerWindow key_window, rle_window, record_window;
  unsigned long long* d_flag;       // scratch of the barrier
  long long* d_status;              // scratch of agree_on_status
  cudaStream_t ship_stream;         // the planes of the result travel beside the writer chain
  cudaEvent_t  ship_ready, ship_done;
};

namespace bwtm
{

// First index with keys[index] >= probe, for each probe.
template<class KeyT>
__global__ void lower_bounds(const KeyT* __restrict__ keys, uint64_t n, const unsigned long long* __restrict__ probes, int n_probes,
                             unsigned long long* __restrict__ out)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n_probes) { return; }
  unsigned long long probe = probes[k];
  uint64_t lo = 0, hi = n;
  while(lo < hi)
  {
    uint64_t mid = lo + (hi - lo) / 2;
    if((unsigned long long)keys[mid] < probe) { lo = mid + 1; } else { hi = mid; }
  }
  out[k] = lo;
}

// The rank structure of the result without decoding the gathered bytes on every rank: every rank holds the plane
// chunks of its slice of the merged sequence (K4 wrote them, chunk 0 at the slice's first position) and stores them,
// shifted onto the global 32-position grid, straight into the record windows of ALL ranks over NVLink. A chunk that
// two slices share (a slice boundary inside 32 positions) is OR-ed in with atomics; the windows start zeroed.
constexpr int MAX_SHIP_PEERS = 16;
constexpr int MAX_SHIP_SLABS = 4;

struct ShipPlan
{
  const uint4* slabs[MAX_SHIP_SLABS];   // plane chunks of the slice's slabs
  uint64_t     slab_chunks;             // chunks per slab (all but the last slab are full)
  uint64_t     begin, end;              // the slice [begin, end) in merged positions
  uint4*       peers[MAX_SHIP_PEERS];   // record windows of all ranks (this one included)
  int          world;
};

__device__ __forceinline__ uint4 ship_local_chunk(const ShipPlan& plan, int64_t chunk, uint64_t chunks)
{
  if(chunk < 0 || (uint64_t)chunk >= chunks) { return make_uint4(0, 0, 0, 0); }
  return __ldg(plan.slabs[(uint64_t)chunk / plan.slab_chunks] + (uint64_t)chunk % plan.slab_chunks);
}

__global__ void __launch_bounds__(256)
ship_planes(ShipPlan plan)
{
  const uint64_t first_global = plan.begin >> 5, last_global = (plan.end + 31) >> 5;   // global chunks [first, last)
  const uint64_t g = first_global + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= last_global) { return; }
  const uint64_t chunks = (plan.end - plan.begin + 31) >> 5;
  // global chunk g starts at slice position 32 g - begin: local chunk `local` at bit `offset`
  const int64_t start = (int64_t)(g << 5) - (int64_t)plan.begin;
  const int64_t local = (start >= 0 ? start >> 5 : -1);
  const uint32_t offset = (uint32_t)(start & 31);
  uint4 lo = ship_local_chunk(plan, local, chunks), hi = make_uint4(0, 0, 0, 0);
  if(offset != 0) { hi = ship_local_chunk(plan, local + 1, chunks); }
  uint4 value;
  value.x = __funnelshift_r(lo.x, hi.x, offset); value.y = __funnelshift_r(lo.y, hi.y, offset); value.z = __funnelshift_r(lo.z, hi.z, offset);
  value.w = 0;
  // positions of the chunk outside the slice belong to the neighbours
  uint32_t keep = 0xFFFFFFFFu;
  if((g << 5) < plan.begin) { keep &= ~low_mask((int)(plan.begin - (g << 5))); }
  if(((g + 1) << 5) > plan.end) { keep &= low_mask((int)(plan.end - (g << 5))); }
  value.x &= keep; value.y &= keep; value.z &= keep;
  const bool shared_chunk = (keep != 0xFFFFFFFFu);
  for(int p = 0; p < plan.world; p++)
  {
    uint4* destination = plan.peers[p] + g;
    if(!shared_chunk) { *destination = value; }
    else
    {
      uint32_t* words = reinterpret_cast<uint32_t*>(destination);
      if(value.x != 0) { atomicOr(words, value.x); }
      if(value.y != 0) { atomicOr(words + 1, value.y); }
      if(value.z != 0) { atomicOr(words + 2, value.z); }
    }
  }
}

// Merges the sorted pieces [offsets[k], offsets[k+1]) of `src` pairwise, ping-ponging between two buffers
// (log2 G streaming rounds instead of another radix sort). *result points to the buffer with the answer.
template<class KeyT>
static int merge_sorted_pieces(KeyT* src, KeyT* dst, std::vector<uint64_t> offsets, cudaStream_t stream, KeyT** result)
{
  DeviceBuffer temp;
  while(offsets.size() > 2)
  {
    std::vector<uint64_t> next; next.push_back(0);
    for(size_t k = 0; k + 1 < offsets.size(); k += 2)
    {
      uint64_t begin = offsets[k], middle = offsets[k + 1], end = (k + 2 < offsets.size() ? offsets[k + 2] : offsets[k + 1]);
      if(end == middle || middle == begin)   // a single (or empty) piece: copy
      {
        if(end > begin) { BWTM_CUDA(cudaMemcpyAsync(dst + begin, src + begin, (end - begin) * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream)); }
      }
      else
      {
        size_t bytes = 0;
        BWTM_CUDA(cub::DeviceMerge::MergeKeys(nullptr, bytes, src + begin, (int)(middle - begin), src + middle, (int)(end - middle), dst + begin, ::cuda::std::less<>{}, stream));
        if(bytes > temp.bytes) { BWTM_TRY(temp.allocate(bytes)); }
        BWTM_CUDA(cub::DeviceMerge::MergeKeys(temp.ptr, bytes, src + begin, (int)(middle - begin), src + middle, (int)(end - middle), dst + begin, ::cuda::std::less<>{}, stream));
        count_launch(2);
      }
      next.push_back(end);
    }
    offsets.swap(next);
    std::swap(src, dst);
  }
  BWTM_CUDA(cudaStreamSynchronize(stream));
  *result = src;
  return BWTM_OK;
}

// All ranks reach this point of their streams before any of them goes on: a one-word all-reduce.
static int stream_barrier(bwtm_comm* comm, cudaStream_t stream)
{
  BWTM_NCCL(nccl()->AllReduce(comm->d_flag, comm->d_flag, 1, ncclUint64, ncclMax, comm->comm, stream));
  return BWTM_OK;
}

// The merge is a collective: a rank that fails locally (allocation, capacity, invalid input) must not leave the
// others waiting in the next all-reduce or receive. Every phase that can fail on one rank alone ends here: the
// smallest status code of all ranks becomes everybody's status, so all ranks return together with an error.
static int agree_on_status(bwtm_comm* comm, int rc, cudaStream_t stream, const char* phase)
{
  long long mine = rc, all = rc;
  if(cudaMemcpyAsync(comm->d_status, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream) != cudaSuccess) { return cuda_failed(cudaGetLastError(), "status upload", __FILE__, __LINE__); }
  BWTM_NCCL(nccl()->AllReduce(comm->d_status, comm->d_status, 1, ncclInt64, ncclMin, comm->comm, stream));
  BWTM_CUDA(cudaMemcpyAsync(&all, comm->d_status, sizeof(all), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  if(all != BWTM_OK && rc == BWTM_OK) { set_error("distributed merge stopped after '%s': another rank failed with status %lld", phase, all); }
  return (int)all;
}

// Fault injection for the tests of the above: BWTM_INJECT_FAILURE=<phase>:<rank> makes that rank fail locally in that
// phase ("search", "interleave" or "writer"), as an allocation failure would.
static int injected_failure(const char* phase, int rank)
{
  const char* spec = getenv("BWTM_INJECT_FAILURE");
  if(spec == nullptr) { return BWTM_OK; }
  size_t length = strlen(phase);
  if(strncmp(spec, phase, length) != 0 || spec[length] != ':' || atoi(spec + length + 1) != rank) { return BWTM_OK; }
  set_error("injected failure in phase '%s' on rank %d", phase, rank);
  return BWTM_ERR_MEMORY;
}

static void window_close(bwtm_comm* comm, PeerWindow* window)
{
  for(int p = 0; p < (int)window->mapped.size(); p++)
  {
    if(p != comm->rank && window->mapped[p] != nullptr) { cudaIpcCloseMemHandle(window->mapped[p]); }
  }
  window->mapped.clear();
}

struct WindowTicket { cudaIpcMemHandle_t handle; unsigned long long ok; };

// Collective: makes `window` at least `need` bytes on every rank (all ranks pass the same value) and maps
// it everywhere. *usable is false when the windows cannot be used; nothing is left half-open in that case.
static int window_reserve(bwtm_comm* comm, PeerWindow* window, uint64_t need, cudaStream_t stream, bool* usable)
{
  *usable = false;
  if(comm->peer_state < 0) { return BWTM_OK; }
  if(need <= window->capacity) { *usable = true; return BWTM_OK; }
  NcclApi* api = nccl();
  const int G = comm->world, r = comm->rank;

  // Nobody may free a window that a peer still has mapped.
  window_close(comm, window);
  BWTM_TRY(stream_barrier(comm, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  if(window->local != nullptr) { cudaFree(window->local); window->local = nullptr; }
  window->capacity = 0;

  uint64_t capacity = div_up(need + (need >> 3) + (1ull << 20), 1ull << 21) << 21;
  WindowTicket mine; std::memset(&mine, 0, sizeof(mine));
  void* fresh = nullptr;
  cudaError_t status = cudaMalloc(&fresh, capacity);
  if(status != cudaSuccess)   // the stream-ordered pool may be sitting on the memory
  {
    cudaGetLastError();
    int device = 0; cudaMemPool_t pool;
    if(cudaGetDevice(&device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { cudaMemPoolTrimTo(pool, 0); }
    status = cudaMalloc(&fresh, capacity);
  }
  if(status == cudaSuccess && cudaIpcGetMemHandle(&mine.handle, fresh) == cudaSuccess) { mine.ok = 1; }
  cudaGetLastError();

  std::vector<WindowTicket> tickets(G);
  DeviceBuffer d_mine, d_all;
  BWTM_TRY(d_mine.allocate(sizeof(WindowTicket))); BWTM_TRY(d_all.allocate(G * sizeof(WindowTicket)));
  auto all_agree = [&](bool* everyone) -> int
  {
    BWTM_CUDA(cudaMemcpyAsync(d_mine.ptr, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream));
    BWTM_NCCL(api->AllGather(d_mine.ptr, d_all.ptr, sizeof(WindowTicket), ncclUint8, comm->comm, stream));
    BWTM_CUDA(cudaMemcpyAsync(tickets.data(), d_all.ptr, G * sizeof(WindowTicket), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
    *everyone = true;
    for(int p = 0; p < G; p++) { *everyone = *everyone && (tickets[p].ok != 0); }
    return BWTM_OK;
  };

  bool everyone = false;
  BWTM_TRY(all_agree(&everyone));
  if(everyone)
  {
    std::vector<WindowTicket> handles = tickets;
    window->mapped.assign(G, nullptr);
    window->mapped[r] = static_cast<uint8_t*>(fresh);
    for(int p = 0; p < G; p++)
    {
      if(p == r) { continue; }
      void* remote = nullptr;
      if(cudaIpcOpenMemHandle(&remote, handles[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; break; }
      window->mapped[p] = static_cast<uint8_t*>(remote);
    }
    BWTM_TRY(all_agree(&everyone));
  }
  if(!everyone)
  {
    window_close(comm, window);
    BWTM_TRY(stream_barrier(comm, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
    if(fresh != nullptr) { cudaFree(fresh); }
    comm->peer_state = -1;
    if(getenv("BWTM_DEBUG") != nullptr) { fprintf(stderr, "bwtm[%d] peer windows unavailable, using NCCL for the bulk data\n", r); }
    return BWTM_OK;
  }
  window->local = static_cast<uint8_t*>(fresh); window->capacity = capacity;
  comm->peer_state = 1;
  *usable = true;
  return BWTM_OK;
}

template<class KeyT> struct NcclKey;
template<> struct NcclKey<uint32_t> { static constexpr ncclDataType_t type = ncclUint32; };
template<> struct NcclKey<uint64_t> { static constexpr ncclDataType_t type = ncclUint64; };

struct PhaseClock   // BWTM_DEBUG=1: host wall clock per phase of the exchange, printed by every rank
{
  bool enabled; int rank; std::chrono::steady_clock::time_point last;
  explicit PhaseClock(int r) : enabled(getenv("BWTM_DEBUG") != nullptr), rank(r), last(std::chrono::steady_clock::now()) {}
  void mark(const char* name)
  {
    if(!enabled) { return; }
    cudaDeviceSynchronize();
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "bwtm[%d] %-22s %8.3f ms\n", rank, name, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
};

template<class KeyT>
static int merge_distributed_impl(bwtm_comm* comm, bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
                                  bwtm_index** result, bwtm_timings* timings)
{
  NcclApi* api = nccl();
  cudaStream_t stream = 0;
  const int G = comm->world, r = comm->rank;
  const uint64_t n_a = a->size, n_b = b->size, m_b = b->sequences;
  EventTimer timer(stream);
  PhaseClock phase(r);

  // 1. search + local sort
  uint64_t seq_first = 0, seq_count = 0;
  bwtm_shard_range(m_b, (uint32_t)r, (uint32_t)G, &seq_first, &seq_count);
  uint64_t capacity = std::min<uint64_t>(n_b, (uint64_t)((double)n_b * ((double)seq_count / (double)m_b) * 1.25) + (1ull << 20));
  DeviceBuffer keys, alt;
  uint64_t local_n = 0;
  const int bits = bit_length_host(n_a);
  KeyT* sorted = nullptr;
  auto search_and_sort = [&]() -> int
  {
    BWTM_TRY(injected_failure("search", r));
    // Every rank walks 1/G of b's sequences but would build the pair records of all of a and b: the two-step walk is
    // chosen as on one GPU, with the inserted share in place of |b|.
    BWTM_TRY(prepare_walk(a, b, (uint64_t)((double)n_b * ((double)seq_count / (double)m_b)), stream, timings));
    timer.start();
    for(int attempt = 0; attempt < 2; attempt++)
    {
      BWTM_TRY(keys.allocate(std::max<uint64_t>(capacity, 1) * sizeof(KeyT)));
      if(seq_count == 0) { break; }
      int rc = walk_sequences<KeyT>(a, b, seq_first, seq_first + seq_count - 1, keys.as<KeyT>(), capacity, &local_n, stream);
      if(rc == BWTM_OK) { break; }
      if(rc != BWTM_ERR_CAPACITY || attempt == 1 || capacity == n_b) { return rc; }
      capacity = n_b;   // sequences of very different lengths: take the upper bound
    }
    timings->search_seconds = timer.stop() * 1e-3;
    timings->walk_kernel_launches = (seq_count > 0 ? 1 : 0);

    timer.start();
    sorted = keys.as<KeyT>();
    if(local_n > 0)
    {
      BWTM_TRY(alt.allocate(local_n * sizeof(KeyT)));
      BWTM_TRY(sort_keys<KeyT>(keys.as<KeyT>(), alt.as<KeyT>(), local_n, bits, &sorted, stream, n_a + 1));
    }
    timings->sort_seconds = timer.stop() * 1e-3;
    return BWTM_OK;
  };
  BWTM_TRY(agree_on_status(comm, search_and_sort(), stream, "search and local sort"));
  phase.mark("walk + local sort");

  // Record windows for the planes of the result (see ship_planes): reserved and zeroed now, long before anybody
  // writes into them; the barrier of the key exchange separates the two.
  const uint64_t result_records = ((n_a + n_b) >> RECORD_SHIFT) + 1;
  bool ship_records = false;
  if(G > 1 && G <= MAX_SHIP_PEERS && options->skip_index == 0 && getenv("BWTM_NCCL_EXCHANGE") == nullptr && getenv("BWTM_RLE_INDEX") == nullptr)
  {
    BWTM_TRY(window_reserve(comm, &(comm->record_window), result_records * 64, stream, &ship_records));
    if(ship_records) { BWTM_CUDA(cudaMemsetAsync(comm->record_window.local, 0, result_records * 64, stream)); }
  }

  // 2. splitters: smallest p with p + #{keys < p} >= k (n_a + n_b) / G
  timer.start();
  auto exchange_start = std::chrono::steady_clock::now();
  const int P = G - 1;
  std::vector<unsigned long long> lo(std::max(P, 1), 0), hi(std::max(P, 1), n_a + 1);
  DeviceBuffer d_probes, d_counts;
  std::vector<unsigned long long> target(std::max(P, 1), 0);
  for(int k = 0; k < P; k++) { target[k] = (unsigned long long)(((__uint128_t)(n_a + n_b) * (k + 1)) / G); }
  // PROBES candidates per splitter and round: five bits of the answer per all-reduce instead of one.
  const int PROBES = 31;
  std::vector<unsigned long long> candidates((size_t)std::max(P, 1) * PROBES, 0), counted((size_t)std::max(P, 1) * PROBES, 0);
  BWTM_TRY(d_probes.allocate(candidates.size() * sizeof(unsigned long long)));
  BWTM_TRY(d_counts.allocate(candidates.size() * sizeof(unsigned long long)));
  for(int iteration = 0; P > 0 && iteration < 66; iteration++)
  {
    bool open = false;
    for(int k = 0; k < P; k++)
    {
      open = open || (lo[k] < hi[k]);
      unsigned long long width = hi[k] - lo[k];
      for(int t = 0; t < PROBES; t++)   // increasing, all in [lo, hi); duplicates are harmless
      {
        candidates[(size_t)k * PROBES + t] = lo[k] + (unsigned long long)(((__uint128_t)width * (t + 1)) / (PROBES + 1));
        if(width > 0 && candidates[(size_t)k * PROBES + t] >= hi[k]) { candidates[(size_t)k * PROBES + t] = hi[k] - 1; }
      }
    }
    if(!open) { break; }
    const int n_probes = P * PROBES;
    BWTM_CUDA(cudaMemcpyAsync(d_probes.ptr, candidates.data(), n_probes * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    lower_bounds<KeyT><<<div_up(n_probes, 128), 128, 0, stream>>>(sorted, local_n, d_probes.as<unsigned long long>(), n_probes, d_counts.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
    BWTM_NCCL(api->AllReduce(d_counts.ptr, d_counts.ptr, n_probes, ncclUint64, ncclSum, comm->comm, stream));
    BWTM_CUDA(cudaMemcpyAsync(counted.data(), d_counts.ptr, n_probes * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
    for(int k = 0; k < P; k++)
    {
      if(lo[k] >= hi[k]) { continue; }
      // p + #{keys < p} is non-decreasing in p: the answer lies after the last candidate that fails.
      for(int t = 0; t < PROBES; t++)
      {
        unsigned long long c = candidates[(size_t)k * PROBES + t];
        if(c < lo[k]) { continue; }
        if(c + counted[(size_t)k * PROBES + t] >= target[k]) { hi[k] = c; break; }
        lo[k] = c + 1;
      }
    }
  }
  phase.mark("splitter search");
  std::vector<unsigned long long> splitter(G + 1, 0);
  for(int k = 0; k < P; k++) { splitter[k + 1] = std::max(lo[k], splitter[k]); }
  splitter[G] = n_a + 1;

  // 3. how many of my values go to each rank
  std::vector<unsigned long long> send_offset(G + 1, 0);
  if(P > 0)
  {
    BWTM_CUDA(cudaMemcpyAsync(d_probes.ptr, splitter.data() + 1, P * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    lower_bounds<KeyT><<<(unsigned)div_up(P, 32), 32, 0, stream>>>(sorted, local_n, d_probes.as<unsigned long long>(), P, d_counts.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
    BWTM_CUDA(cudaMemcpyAsync(send_offset.data() + 1, d_counts.ptr, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
  }
  send_offset[G] = local_n;
  std::vector<unsigned long long> send_count(G, 0), matrix((size_t)G * G, 0);
  for(int k = 0; k < G; k++) { send_count[k] = send_offset[k + 1] - send_offset[k]; }
  DeviceBuffer d_send_count, d_matrix;
  BWTM_TRY(d_send_count.allocate(G * sizeof(unsigned long long)));
  BWTM_TRY(d_matrix.allocate((size_t)G * G * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_send_count.ptr, send_count.data(), G * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  BWTM_NCCL(api->AllGather(d_send_count.ptr, d_matrix.ptr, G, ncclUint64, comm->comm, stream));
  BWTM_CUDA(cudaMemcpyAsync(matrix.data(), d_matrix.ptr, (size_t)G * G * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));

  uint64_t recv_total = 0, b_lo = 0, all_values = 0;
  std::vector<uint64_t> recv_offset(G + 1, 0);
  for(int src = 0; src < G; src++)
  {
    recv_offset[src + 1] = recv_offset[src] + matrix[(size_t)src * G + r];
    for(int dst = 0; dst < G; dst++) { all_values += matrix[(size_t)src * G + dst]; if(dst < r) { b_lo += matrix[(size_t)src * G + dst]; } }
  }
  recv_total = recv_offset[G];
  if(all_values != n_b)
  {
    set_error("the rank array has %llu values but the inserted BWT has %llu symbols: not a valid multi-string BWT",
              (unsigned long long)all_values, (unsigned long long)n_b);
    return BWTM_ERR_INTERNAL;
  }
  timings->ra_values = all_values;

  phase.mark("count matrix");
  // 4. all-to-all by A-position range, then 5. local sort of what arrived (G sorted pieces)
  DeviceBuffer received, received_alt;
  const bool force_nccl = (getenv("BWTM_NCCL_EXCHANGE") != nullptr);
  if(force_nccl) { comm->peer_state = -1; }
  bool direct = false;
  if(G > 1)
  {
    uint64_t largest = 0;   // every rank knows the whole matrix, so all ask for the same size
    for(int dst = 0; dst < G; dst++)
    {
      uint64_t sum = 0;
      for(int src = 0; src < G; src++) { sum += matrix[(size_t)src * G + dst]; }
      largest = std::max(largest, sum);
    }
    BWTM_TRY(window_reserve(comm, &(comm->key_window), std::max<uint64_t>(largest, 1) * sizeof(KeyT), stream, &direct));
  }
  KeyT* arrived = nullptr;
  if(direct)
  {
    // Every piece goes straight into its place in the owner's window. Rank r starts with peer r + 1 so
    // that no receiver is the target of two senders at once.
    for(int step = 0; step < G; step++)
    {
      int peer = (r + step) % G;
      if(send_count[peer] == 0) { continue; }
      uint64_t offset = 0;
      for(int src = 0; src < r; src++) { offset += matrix[(size_t)src * G + peer]; }
      BWTM_CUDA(cudaMemcpyAsync(reinterpret_cast<KeyT*>(comm->key_window.mapped[peer]) + offset, sorted + send_offset[peer],
                                send_count[peer] * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream));
    }
    BWTM_TRY(stream_barrier(comm, stream));   // all pieces have landed everywhere
    arrived = reinterpret_cast<KeyT*>(comm->key_window.local);
  }
  else
  {
    BWTM_TRY(agree_on_status(comm, received.allocate(std::max<uint64_t>(recv_total, 1) * sizeof(KeyT)), stream, "receive buffer"));
    BWTM_NCCL(api->GroupStart());
    for(int peer = 0; peer < G; peer++)
    {
      if(send_count[peer] > 0) { BWTM_NCCL(api->Send(sorted + send_offset[peer], send_count[peer], NcclKey<KeyT>::type, peer, comm->comm, stream)); }
      uint64_t incoming = recv_offset[peer + 1] - recv_offset[peer];
      if(incoming > 0) { BWTM_NCCL(api->Recv(received.as<KeyT>() + recv_offset[peer], incoming, NcclKey<KeyT>::type, peer, comm->comm, stream)); }
    }
    BWTM_NCCL(api->GroupEnd());
    arrived = received.as<KeyT>();
  }
  BWTM_CUDA(cudaStreamSynchronize(stream));
  phase.mark("all-to-all");
  keys.release(); alt.release();
  KeyT* slice_keys = arrived;
  auto merge_received = [&]() -> int
  {
    if(recv_total == 0 || G == 1) { return BWTM_OK; }
    BWTM_TRY(received_alt.allocate(recv_total * sizeof(KeyT)));
    if(recv_total < 0x7FFFFFFFull) { return merge_sorted_pieces<KeyT>(arrived, received_alt.as<KeyT>(), recv_offset, stream, &slice_keys); }
    return sort_keys<KeyT>(arrived, received_alt.as<KeyT>(), recv_total, bits, &slice_keys, stream);
  };
  int merge_rc = merge_received();
  timings->exchange_seconds = timer.stop() * 1e-3;
  phase.mark("sort received");

  // 6. my slice of the merged BWT
  uint64_t a_lo = std::min<uint64_t>(splitter[r], n_a), a_hi = std::min<uint64_t>(splitter[r + 1], n_a);
  uint64_t begin = a_lo + b_lo, end = a_hi + b_lo + recv_total;
  // The slice is cut into at most MAX_PARALLEL_SLABS slabs (each below the 2^31 item limit of the run
  // detection). Symbols, maximal runs and the state-free half of the writer (K4, K3, scan, tile maps) of
  // every slab are computed on all ranks at once; only SlabEncoder::advance(), which moves the 88-byte writer
  // state past a slab, is chained through the slabs and ranks; the bytes are written after the state has
  // been passed on. Slices that would need more slabs (or an explicit small slab size) fall back to running
  // the whole interleave inside the chain.
  const uint64_t MAX_PARALLEL_SLABS = 4;
  uint64_t slice = end - begin;
  uint64_t slab = (options->slab_symbols == 0 ? clamp_slab(MAX_SLAB_SYMBOLS, slice, true) : clamp_slab(options->slab_symbols, slice));
  uint64_t n_slabs = div_up(slice, slab);
  bool parallel = (slice > 0 && n_slabs <= MAX_PARALLEL_SLABS);

  DeviceBuffer tile_j, control;
  std::vector<DeviceBuffer> merged(parallel ? n_slabs : 0);   // plane chunks of every slab, read again by emit()
  std::vector<SlabEncoder> encoders(parallel ? n_slabs : 1);
  float interleave_ms = 0.0f, encode_ms = 0.0f;
  auto interleave_slabs = [&]() -> int
  {
    BWTM_TRY(merge_rc);
    BWTM_TRY(injected_failure("interleave", r));
    BWTM_TRY(control.allocate(sizeof(EncodeControl)));
    if(!parallel) { return BWTM_OK; }
    BWTM_TRY(tile_j.allocate((slab / interleave_tile_size() + 2) * sizeof(uint64_t)));
    for(uint64_t k = 0; k < n_slabs; k++)
    {
      uint64_t p0 = begin + k * slab, p1 = std::min(p0 + slab, end);
      BWTM_TRY(encoders[k].init(p1 - p0, stream));
      BWTM_TRY(merged[k].allocate(div_up(p1 - p0, interleave_tile_size()) * (interleave_tile_size() / 2)));
      timer.start();
      BWTM_TRY(interleave_slab<KeyT>(a, b, slice_keys, b_lo, recv_total, p0, p1, merged[k].as<uint4>(), tile_j.as<uint64_t>(), stream));
      interleave_ms += timer.stop();
      timer.start();
      BWTM_TRY(encoders[k].detect(merged[k].as<uint4>(), p1 - p0, stream));
      encode_ms += timer.stop();
    }
    return BWTM_OK;
  };
  BWTM_TRY(agree_on_status(comm, interleave_slabs(), stream, "merge of the received values and interleave"));
  phase.mark("interleave + runs");
  // All ranks take the same route to the result's rank structure: planes shipped (every slice was interleaved in
  // parallel slabs and the key exchange went through the windows, whose barrier ordered the zeroing) or bytes decoded.
  {
    unsigned long long can_ship = (ship_records && direct && (parallel || slice == 0) ? 1 : 0), everybody = 0;
    BWTM_CUDA(cudaMemcpyAsync(comm->d_flag, &can_ship, sizeof(can_ship), cudaMemcpyHostToDevice, stream));
    BWTM_NCCL(api->AllReduce(comm->d_flag, comm->d_flag, 1, ncclUint64, ncclMin, comm->comm, stream));
    BWTM_CUDA(cudaMemcpyAsync(&everybody, comm->d_flag, sizeof(everybody), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaMemsetAsync(comm->d_flag, 0, sizeof(unsigned long long), stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
    ship_records = (everybody != 0);
  }
  bool shipping = false;   // the side stream holds work the main stream has to wait for
  if(ship_records && slice > 0)
  {
    ShipPlan plan; std::memset(&plan, 0, sizeof(plan));
    for(uint64_t k = 0; k < n_slabs; k++) { plan.slabs[k] = merged[k].as<uint4>(); }
    plan.slab_chunks = slab / 32; plan.begin = begin; plan.end = end; plan.world = G;
    for(int p = 0; p < G; p++) { plan.peers[p] = reinterpret_cast<uint4*>(comm->record_window.mapped[p]); }
    uint64_t global_chunks = ((end + 31) >> 5) - (begin >> 5);
    // On a side stream: the stores over NVLink run beside the writer chain and the gather of the slices; the main
    // stream picks the side stream up again before the barrier that ends the gather.
    cudaStream_t ship_on = (comm->ship_stream != nullptr ? comm->ship_stream : stream);
    if(ship_on != stream) { BWTM_CUDA(cudaEventRecord(comm->ship_ready, stream)); BWTM_CUDA(cudaStreamWaitEvent(ship_on, comm->ship_ready, 0)); }
    ship_planes<<<(unsigned)div_up(global_chunks, 256), 256, 0, ship_on>>>(plan);
    BWTM_LAUNCH_CHECK();
    if(ship_on != stream) { BWTM_CUDA(cudaEventRecord(comm->ship_done, ship_on)); }
    shipping = (ship_on != stream);
  }
  phase.mark("ship planes");

  // 7. the writer state comes from the previous slice and goes to the next one
  if(r > 0)
  {
    BWTM_NCCL(api->Recv(control.ptr, sizeof(EncodeControl), ncclUint8, r - 1, comm->comm, stream));
  }
  else { BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream)); }
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  // A rank that failed earlier in the chain passes POISONED_STATE on: the ranks after it skip their part but keep
  // the chain moving, and everybody learns the status in agree_on_status below.
  const unsigned long long POISONED_STATE = ~0ull;
  bool upstream_failed = (ctl.out_size == POISONED_STATE);
  OutputBuffer out = { nullptr, 0, (upstream_failed ? 0 : ctl.out_size), nullptr };
  uint64_t estimate = (a->rle_bytes + b->rle_bytes) / G;
  int rc = BWTM_OK;
  if(upstream_failed) { set_error("an earlier rank of the writer chain failed"); rc = BWTM_ERR_COMM; }
  else { rc = ensure_capacity(&out, out.origin + estimate + (estimate >> 2) + (1 << 20), out.origin, stream); }
  if(rc == BWTM_OK) { rc = injected_failure("writer", r); }
  timer.start();
  if(rc == BWTM_OK && slice > 0)
  {
    if(parallel)
    {
      for(uint64_t k = 0; rc == BWTM_OK && k < n_slabs; k++) { rc = encoders[k].advance(&out, control.as<EncodeControl>(), stream); }
    }
    else
    {
      rc = interleave_range<KeyT>(a, b, slice_keys, b_lo, recv_total, begin, end, options->slab_symbols, &out,
                                  control.as<EncodeControl>(), false, &interleave_ms, &encode_ms, stream);
    }
  }
  if(rc != BWTM_OK && !upstream_failed)
  {
    EncodeControl poisoned; std::memset(&poisoned, 0, sizeof(poisoned)); poisoned.out_size = POISONED_STATE;
    cudaMemcpyAsync(control.ptr, &poisoned, sizeof(poisoned), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
  }
  // The next slice only needs the state after this one: it goes out before the bytes are written.
  if(r < G - 1) { BWTM_NCCL(api->Send(control.ptr, sizeof(EncodeControl), ncclUint8, r + 1, comm->comm, stream)); }
  if(parallel)
  {
    for(uint64_t k = 0; rc == BWTM_OK && k < n_slabs; k++) { rc = encoders[k].emit(&out, stream); }
  }
  if(rc == BWTM_OK && r == G - 1)
  {
    if(!parallel) { rc = encoders[0].init(4096, stream); }
    if(rc == BWTM_OK) { rc = encoders[0].finish(&out, control.as<EncodeControl>(), stream); }
  }
  if(parallel) { encode_ms += timer.stop(); } else { timer.stop(); }
  rc = agree_on_status(comm, rc, stream, "chained writer");
  if(rc != BWTM_OK)
  {
    if(shipping) { cudaStreamSynchronize(comm->ship_stream); }   // the plane chunks are freed on return: nobody may still read them
    device_free(out.ptr);
    return rc;
  }
  BWTM_CUDA(cudaMemcpyAsync(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  timings->interleave_seconds = interleave_ms * 1e-3;
  timings->encode_seconds = encode_ms * 1e-3;
  received.release(); received_alt.release();
  // The plane chunks are freed in the order of the main stream: it must not pass the side stream that still reads them.
  if(shipping) { BWTM_CUDA(cudaStreamWaitEvent(stream, comm->ship_done, 0)); shipping = false; }
  encoders.clear(); merged.clear();

  phase.mark("chained writer");
  // 8. every rank gets the complete run-length BWT
  timer.start();
  unsigned long long mine[3] = { out.origin, ctl.out_size - out.origin, ctl.runs_total };
  std::vector<unsigned long long> slices((size_t)3 * G, 0);
  DeviceBuffer d_mine, d_slices;
  BWTM_TRY(d_mine.allocate(sizeof(mine))); BWTM_TRY(d_slices.allocate(slices.size() * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_mine.ptr, mine, sizeof(mine), cudaMemcpyHostToDevice, stream));
  BWTM_NCCL(api->AllGather(d_mine.ptr, d_slices.ptr, 3, ncclUint64, comm->comm, stream));
  BWTM_CUDA(cudaMemcpyAsync(slices.data(), d_slices.ptr, slices.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t total_bytes = slices[(size_t)3 * (G - 1)] + slices[(size_t)3 * (G - 1) + 1];
  timings->merged_bytes = total_bytes; timings->merged_runs = slices[(size_t)3 * (G - 1) + 2];
  OutputBuffer full = { nullptr, 0, 0, nullptr };
  bool direct_gather = false;
  if(G > 1) { BWTM_TRY(window_reserve(comm, &(comm->rle_window), total_bytes + RLE_PADDING, stream, &direct_gather)); }
  for(int k = 0; k < G; k++)
  {
    uint64_t offset = slices[(size_t)3 * k], bytes = slices[(size_t)3 * k + 1];
    if(offset + bytes > total_bytes) { set_error("inconsistent slice layout"); device_free(out.ptr); return BWTM_ERR_INTERNAL; }
  }
  if(direct_gather)
  {
    // My slice goes to its place in every replica's window, my own included.
    uint64_t offset = slices[(size_t)3 * r], bytes = slices[(size_t)3 * r + 1];
    for(int step = 0; bytes > 0 && step < G; step++)
    {
      int peer = (r + 1 + step) % G;
      BWTM_CUDA(cudaMemcpyAsync(comm->rle_window.mapped[peer] + offset, out.ptr, bytes, cudaMemcpyDeviceToDevice, stream));
    }
    if(shipping) { BWTM_CUDA(cudaStreamWaitEvent(stream, comm->ship_done, 0)); shipping = false; }   // the barrier covers the planes too
    BWTM_TRY(stream_barrier(comm, stream));
    full.ptr = comm->rle_window.local; full.capacity = comm->rle_window.capacity; full.borrowed = true;
  }
  else
  {
    rc = agree_on_status(comm, ensure_capacity(&full, total_bytes + RLE_PADDING, 0, stream), stream, "gather buffer");
    if(rc != BWTM_OK) { device_free(out.ptr); device_free(full.ptr); return rc; }
    for(int k = 0; k < G; k++)
    {
      uint64_t offset = slices[(size_t)3 * k], bytes = slices[(size_t)3 * k + 1];
      if(bytes == 0) { continue; }
      BWTM_NCCL(api->Broadcast(k == r ? (const void*)out.ptr : (const void*)(full.ptr + offset), full.ptr + offset, bytes, ncclUint8, k, comm->comm, stream));
    }
  }
  if(ship_records && !direct_gather)   // the NCCL broadcasts do not order the shipped planes: an explicit barrier does
  {
    if(shipping) { BWTM_CUDA(cudaStreamWaitEvent(stream, comm->ship_done, 0)); shipping = false; }
    BWTM_TRY(stream_barrier(comm, stream));
  }
  BWTM_CUDA(cudaStreamSynchronize(stream));
  device_free(out.ptr);
  timings->exchange_seconds += timer.stop() * 1e-3;
  (void)exchange_start;

  phase.mark("gather slices");
  // 9. replica of the rank structure
  timer.start();
  uint64_t counts[SIGMA];
  for(int c = 0; c < SIGMA; c++) { counts[c] = a->counts[c] + b->counts[c]; }
  if(ship_records)
  {
    // The barrier after the gather of the slices also covered the shipped planes: the window holds the planes of
    // the whole result. They become the records of the new index (a copy: the window stays with the communicator).
    DeviceBuffer records;
    rc = records.allocate(result_records * 64, true);
    if(rc == BWTM_OK && cudaMemcpyAsync(records.ptr, comm->record_window.local, result_records * 64, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    {
      rc = cuda_failed(cudaGetLastError(), "copy of the record window", __FILE__, __LINE__);
    }
    if(rc == BWTM_OK) { rc = finish_index(&full, total_bytes, counts, a->sequences + b->sequences, false, stream, result, &records, n_a + n_b); }
  }
  else { rc = finish_index(&full, total_bytes, counts, a->sequences + b->sequences, options->skip_index != 0, stream, result); }
  if(!full.borrowed) { device_free(full.ptr); }
  timings->index_seconds += timer.stop() * 1e-3;
  phase.mark("index");
  return rc;
}


//------------------------------------------------------------------------------
// The distributed merge with the search in batches (options.sequence_blocks > 1): every rank keeps its rank-array
// values as S sorted runs (bwtm_batches.cuh), the exchange moves the pieces of every run straight into the owners'
// windows (G S sorted pieces per rank), and the owner merges them range by range while it interleaves, inside the
// writer chain. Work memory per rank: its values once (+ one batch-sized scratch), then the window; no second copy.

// counts[p] = #{keys < probes[p]} over all runs; per_run[p * S + k] (optional) = the part of run k.
template<class KeyT>
__global__ void lower_bounds_runs(const KeyT* __restrict__ keys, const unsigned long long* __restrict__ run_offsets, int S,
                                  const unsigned long long* __restrict__ probes, int n_probes,
                                  unsigned long long* __restrict__ counts, unsigned long long* __restrict__ per_run)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= n_probes) { return; }
  unsigned long long probe = probes[p], total = 0;
  for(int k = 0; k < S; k++)
  {
    unsigned long long lo = run_offsets[k], hi = run_offsets[k + 1];
    while(lo < hi)
    {
      unsigned long long mid = lo + (hi - lo) / 2;
      if((unsigned long long)keys[mid] < probe) { lo = mid + 1; } else { hi = mid; }
    }
    if(per_run != nullptr) { per_run[(size_t)p * S + k] = lo - run_offsets[k]; }
    total += lo - run_offsets[k];
  }
  counts[p] = total;
}

// Steps 8 and 9 of the distributed merge for a slice that sits in `out`: every rank gets all slices (peer windows
// or NCCL broadcasts) and builds its replica of the rank structure from the bytes.
static int gather_slices_and_index(bwtm_comm* comm, const bwtm_index* a, const bwtm_index* b, const bwtm_merge_options* options,
                                   OutputBuffer& out, const EncodeControl& ctl, cudaStream_t stream, bwtm_index** result, bwtm_timings* timings)
{
  NcclApi* api = nccl();
  const int G = comm->world, r = comm->rank;
  EventTimer timer(stream);
  timer.start();
  unsigned long long mine[3] = { out.origin, ctl.out_size - out.origin, ctl.runs_total };
  std::vector<unsigned long long> slices((size_t)3 * G, 0);
  DeviceBuffer d_mine, d_slices;
  BWTM_TRY(d_mine.allocate(sizeof(mine))); BWTM_TRY(d_slices.allocate(slices.size() * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_mine.ptr, mine, sizeof(mine), cudaMemcpyHostToDevice, stream));
  BWTM_NCCL(api->AllGather(d_mine.ptr, d_slices.ptr, 3, ncclUint64, comm->comm, stream));
  BWTM_CUDA(cudaMemcpyAsync(slices.data(), d_slices.ptr, slices.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  uint64_t total_bytes = slices[(size_t)3 * (G - 1)] + slices[(size_t)3 * (G - 1) + 1];
  timings->merged_bytes = total_bytes; timings->merged_runs = slices[(size_t)3 * (G - 1) + 2];
  for(int k = 0; k < G; k++)
  {
    if(slices[(size_t)3 * k] + slices[(size_t)3 * k + 1] > total_bytes) { set_error("inconsistent slice layout"); return BWTM_ERR_INTERNAL; }
  }
  OutputBuffer full = { nullptr, 0, 0, nullptr };
  bool direct_gather = false;
  BWTM_TRY(window_reserve(comm, &(comm->rle_window), total_bytes + RLE_PADDING, stream, &direct_gather));
  int rc = BWTM_OK;
  if(direct_gather)
  {
    uint64_t offset = slices[(size_t)3 * r], bytes = slices[(size_t)3 * r + 1];
    for(int step = 0; bytes > 0 && step < G; step++)
    {
      int peer = (r + 1 + step) % G;
      BWTM_CUDA(cudaMemcpyAsync(comm->rle_window.mapped[peer] + offset, out.ptr, bytes, cudaMemcpyDeviceToDevice, stream));
    }
    BWTM_TRY(stream_barrier(comm, stream));
    full.ptr = comm->rle_window.local; full.capacity = comm->rle_window.capacity; full.borrowed = true;
  }
  else
  {
    rc = agree_on_status(comm, ensure_capacity(&full, total_bytes + RLE_PADDING, 0, stream), stream, "gather buffer");
    if(rc != BWTM_OK) { device_free(full.ptr); return rc; }
    for(int k = 0; k < G; k++)
    {
      uint64_t offset = slices[(size_t)3 * k], bytes = slices[(size_t)3 * k + 1];
      if(bytes == 0) { continue; }
      BWTM_NCCL(api->Broadcast(k == r ? (const void*)out.ptr : (const void*)(full.ptr + offset), full.ptr + offset, bytes, ncclUint8, k, comm->comm, stream));
    }
  }
  BWTM_CUDA(cudaStreamSynchronize(stream));
  timings->exchange_seconds += timer.stop() * 1e-3;
  timer.start();
  uint64_t counts[SIGMA];
  for(int c = 0; c < SIGMA; c++) { counts[c] = a->counts[c] + b->counts[c]; }
  rc = finish_index(&full, total_bytes, counts, a->sequences + b->sequences, options->skip_index != 0, stream, result);
  if(!full.borrowed) { device_free(full.ptr); }
  timings->index_seconds += timer.stop() * 1e-3;
  return rc;
}

template<class KeyT>
static int merge_distributed_batched(bwtm_comm* comm, bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options, int S,
                                     bwtm_index** result, bwtm_timings* timings)
{
  NcclApi* api = nccl();
  cudaStream_t stream = 0;
  const int G = comm->world, r = comm->rank;
  const uint64_t n_a = a->size, n_b = b->size, m_b = b->sequences;
  const int bits = bit_length_host(n_a);
  EventTimer timer(stream);
  PhaseClock phase(r);

  // 1. search + sort in batches: S sorted runs
  uint64_t seq_first = 0, seq_count = 0;
  bwtm_shard_range(m_b, (uint32_t)r, (uint32_t)G, &seq_first, &seq_count);
  uint64_t capacity = std::min<uint64_t>(n_b, (uint64_t)((double)n_b * ((double)seq_count / (double)m_b) * 1.05) + (1ull << 20));
  DeviceBuffer runs, scratch;
  std::vector<unsigned long long> run_offsets(S + 1, 0);
  auto search_in_batches = [&]() -> int
  {
    BWTM_TRY(injected_failure("search", r));
    BWTM_TRY(prepare_walk(a, b, (uint64_t)((double)n_b * ((double)seq_count / (double)m_b)), stream, timings));
    float search_ms = 0.0f, sort_ms = 0.0f;
    for(int attempt = 0; attempt < 2; attempt++)
    {
      BWTM_TRY(runs.allocate(std::max<uint64_t>(capacity, 1) * sizeof(KeyT)));
      bool fits = true;
      std::fill(run_offsets.begin(), run_offsets.end(), 0);
      for(int k = 0; k < S && fits; k++)
      {
        uint64_t first = seq_first + (uint64_t)(((__uint128_t)seq_count * k) / S), last = seq_first + (uint64_t)(((__uint128_t)seq_count * (k + 1)) / S);
        run_offsets[k + 1] = run_offsets[k];
        if(last == first) { continue; }
        uint64_t emitted = 0;
        timer.start();
        int rc = walk_sequences<KeyT>(a, b, first, last - 1, runs.as<KeyT>() + run_offsets[k], capacity - run_offsets[k], &emitted, stream);
        search_ms += timer.stop();
        if(rc == BWTM_ERR_CAPACITY) { fits = false; break; }
        BWTM_TRY(rc);
        run_offsets[k + 1] = run_offsets[k] + emitted;
        if(emitted == 0) { continue; }
        timer.start();
        if(emitted * sizeof(KeyT) > scratch.bytes) { BWTM_TRY(scratch.allocate(emitted * sizeof(KeyT) + (emitted * sizeof(KeyT) >> 3))); }
        KeyT* sorted = nullptr;
        BWTM_TRY(sort_keys<KeyT>(runs.as<KeyT>() + run_offsets[k], scratch.as<KeyT>(), emitted, bits, &sorted, stream, n_a + 1));
        if(sorted != runs.as<KeyT>() + run_offsets[k])
        {
          BWTM_CUDA(cudaMemcpyAsync(runs.as<KeyT>() + run_offsets[k], sorted, emitted * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream));
        }
        sort_ms += timer.stop();
      }
      if(fits) { break; }
      for(int k = S; k > 0; k--) { run_offsets[k] = run_offsets[0]; }
      if(attempt == 1 || capacity == n_b) { set_error("rank array buffer too small"); return BWTM_ERR_CAPACITY; }
      capacity = n_b;   // sequences of very different lengths: take the upper bound
    }
    scratch.release();
    timings->search_seconds = search_ms * 1e-3; timings->sort_seconds = sort_ms * 1e-3;
    timings->walk_kernel_launches = S; timings->search_batches = S;
    return BWTM_OK;
  };
  BWTM_TRY(agree_on_status(comm, search_in_batches(), stream, "search and local sort"));
  const uint64_t local_n = run_offsets[S];
  DeviceBuffer d_run_offsets;
  BWTM_TRY(d_run_offsets.allocate((S + 1) * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_run_offsets.ptr, run_offsets.data(), (S + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  phase.mark("walk + local sort");

  // 2. splitters: smallest p with p + #{keys < p} >= k (n_a + n_b) / G
  timer.start();
  const int P = G - 1;
  std::vector<unsigned long long> lo(std::max(P, 1), 0), hi(std::max(P, 1), n_a + 1), target(std::max(P, 1), 0);
  for(int k = 0; k < P; k++) { target[k] = (unsigned long long)(((__uint128_t)(n_a + n_b) * (k + 1)) / G); }
  const int PROBES = 31;
  std::vector<unsigned long long> candidates((size_t)std::max(P, 1) * PROBES, 0), counted((size_t)std::max(P, 1) * PROBES, 0);
  DeviceBuffer d_probes, d_counts, d_per_run;
  BWTM_TRY(d_probes.allocate(candidates.size() * sizeof(unsigned long long)));
  BWTM_TRY(d_counts.allocate(candidates.size() * sizeof(unsigned long long)));
  BWTM_TRY(d_per_run.allocate((size_t)std::max(P, 1) * S * sizeof(unsigned long long)));
  for(int iteration = 0; P > 0 && iteration < 66; iteration++)
  {
    bool open = false;
    for(int k = 0; k < P; k++)
    {
      open = open || (lo[k] < hi[k]);
      unsigned long long width = hi[k] - lo[k];
      for(int t = 0; t < PROBES; t++)
      {
        candidates[(size_t)k * PROBES + t] = lo[k] + (unsigned long long)(((__uint128_t)width * (t + 1)) / (PROBES + 1));
        if(width > 0 && candidates[(size_t)k * PROBES + t] >= hi[k]) { candidates[(size_t)k * PROBES + t] = hi[k] - 1; }
      }
    }
    if(!open) { break; }
    const int n_probes = P * PROBES;
    BWTM_CUDA(cudaMemcpyAsync(d_probes.ptr, candidates.data(), n_probes * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    lower_bounds_runs<KeyT><<<div_up(n_probes, 128), 128, 0, stream>>>(runs.as<KeyT>(), d_run_offsets.as<unsigned long long>(), S, d_probes.as<unsigned long long>(),
                                                                        n_probes, d_counts.as<unsigned long long>(), nullptr);
    BWTM_LAUNCH_CHECK();
    BWTM_NCCL(api->AllReduce(d_counts.ptr, d_counts.ptr, n_probes, ncclUint64, ncclSum, comm->comm, stream));
    BWTM_CUDA(cudaMemcpyAsync(counted.data(), d_counts.ptr, n_probes * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
    for(int k = 0; k < P; k++)
    {
      if(lo[k] >= hi[k]) { continue; }
      for(int t = 0; t < PROBES; t++)
      {
        unsigned long long c = candidates[(size_t)k * PROBES + t];
        if(c < lo[k]) { continue; }
        if(c + counted[(size_t)k * PROBES + t] >= target[k]) { hi[k] = c; break; }
        lo[k] = c + 1;
      }
    }
  }
  std::vector<unsigned long long> splitter(G + 1, 0);
  for(int k = 0; k < P; k++) { splitter[k + 1] = std::max(lo[k], splitter[k]); }
  splitter[G] = n_a + 1;
  phase.mark("splitter search");

  // 3. the pieces: bound[d][k] = #{keys of run k below splitter[d]}
  std::vector<unsigned long long> bound((size_t)(G + 1) * S, 0);
  if(P > 0)
  {
    BWTM_CUDA(cudaMemcpyAsync(d_probes.ptr, splitter.data() + 1, P * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    lower_bounds_runs<KeyT><<<(unsigned)div_up(P, 32), 32, 0, stream>>>(runs.as<KeyT>(), d_run_offsets.as<unsigned long long>(), S, d_probes.as<unsigned long long>(),
                                                                        P, d_counts.as<unsigned long long>(), d_per_run.as<unsigned long long>());
    BWTM_LAUNCH_CHECK();
    BWTM_CUDA(cudaMemcpyAsync(bound.data() + S, d_per_run.ptr, (size_t)P * S * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    BWTM_CUDA(cudaStreamSynchronize(stream));
  }
  for(int k = 0; k < S; k++) { bound[(size_t)G * S + k] = run_offsets[k + 1] - run_offsets[k]; }
  // piece[src][dst][k], all-gathered
  std::vector<unsigned long long> my_pieces((size_t)G * S, 0), pieces((size_t)G * G * S, 0);
  for(int d = 0; d < G; d++) { for(int k = 0; k < S; k++) { my_pieces[(size_t)d * S + k] = bound[(size_t)(d + 1) * S + k] - bound[(size_t)d * S + k]; } }
  DeviceBuffer d_my_pieces, d_pieces;
  BWTM_TRY(d_my_pieces.allocate(my_pieces.size() * sizeof(unsigned long long)));
  BWTM_TRY(d_pieces.allocate(pieces.size() * sizeof(unsigned long long)));
  BWTM_CUDA(cudaMemcpyAsync(d_my_pieces.ptr, my_pieces.data(), my_pieces.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
  BWTM_NCCL(api->AllGather(d_my_pieces.ptr, d_pieces.ptr, my_pieces.size(), ncclUint64, comm->comm, stream));
  BWTM_CUDA(cudaMemcpyAsync(pieces.data(), d_pieces.ptr, pieces.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  auto piece = [&](int src, int dst, int k) -> unsigned long long { return pieces[((size_t)src * G + dst) * S + k]; };
  uint64_t all_values = 0, b_lo = 0, largest = 0;
  std::vector<uint64_t> into(G, 0);   // values every rank receives
  for(int src = 0; src < G; src++) { for(int dst = 0; dst < G; dst++) { for(int k = 0; k < S; k++)
  {
    all_values += piece(src, dst, k); into[dst] += piece(src, dst, k);
    if(dst < r) { b_lo += piece(src, dst, k); }
  } } }
  for(int dst = 0; dst < G; dst++) { largest = std::max(largest, into[dst]); }
  if(all_values != n_b)
  {
    set_error("the rank array has %llu values but the inserted BWT has %llu symbols: not a valid multi-string BWT",
              (unsigned long long)all_values, (unsigned long long)n_b);
    return BWTM_ERR_INTERNAL;
  }
  timings->ra_values = all_values;
  const uint64_t recv_total = into[r];
  phase.mark("count matrix");

  // 4. every piece goes straight to its place in the owner's window: ordered by (source rank, run)
  bool direct = false;
  BWTM_TRY(window_reserve(comm, &(comm->key_window), std::max<uint64_t>(largest, 1) * sizeof(KeyT), stream, &direct));
  {
    int rc = (direct ? BWTM_OK : BWTM_ERR_COMM);
    if(!direct) { set_error("the batched distributed merge needs peer windows (CUDA IPC between the GPUs)"); }
    BWTM_TRY(agree_on_status(comm, rc, stream, "peer windows"));
  }
  for(int step = 0; step < G; step++)
  {
    int peer = (r + step) % G;
    uint64_t offset = 0;
    for(int src = 0; src < r; src++) { for(int k = 0; k < S; k++) { offset += piece(src, peer, k); } }
    for(int k = 0; k < S; k++)
    {
      uint64_t count = piece(r, peer, k);
      if(count > 0)
      {
        BWTM_CUDA(cudaMemcpyAsync(reinterpret_cast<KeyT*>(comm->key_window.mapped[peer]) + offset, runs.as<KeyT>() + run_offsets[k] + bound[(size_t)peer * S + k],
                                  count * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream));
      }
      offset += count;
    }
  }
  BWTM_TRY(stream_barrier(comm, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  runs.release();
  const KeyT* arrived = reinterpret_cast<const KeyT*>(comm->key_window.local);
  std::vector<unsigned long long> arrived_offsets(1, 0);   // G S sorted runs
  for(int src = 0; src < G; src++) { for(int k = 0; k < S; k++) { arrived_offsets.push_back(arrived_offsets.back() + piece(src, r, k)); } }
  timings->exchange_seconds = timer.stop() * 1e-3;
  phase.mark("all-to-all");

  // 5. + 6. my slice of the merged BWT, merged range by range inside the writer chain
  const uint64_t a_lo = std::min<uint64_t>(splitter[r], n_a), a_hi = std::min<uint64_t>(splitter[r + 1], n_a);
  const uint64_t begin = a_lo + b_lo, end = a_hi + b_lo + recv_total;
  DeviceBuffer control;
  BWTM_TRY(agree_on_status(comm, control.allocate(sizeof(EncodeControl)), stream, "writer state"));
  if(r > 0) { BWTM_NCCL(api->Recv(control.ptr, sizeof(EncodeControl), ncclUint8, r - 1, comm->comm, stream)); }
  else { BWTM_CUDA(cudaMemsetAsync(control.ptr, 0, sizeof(EncodeControl), stream)); }
  EncodeControl ctl;
  BWTM_CUDA(cudaMemcpyAsync(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  const unsigned long long POISONED_STATE = ~0ull;
  const bool upstream_failed = (ctl.out_size == POISONED_STATE);
  OutputBuffer out = { nullptr, 0, (upstream_failed ? 0 : ctl.out_size), nullptr };
  const uint64_t estimate = (a->rle_bytes + b->rle_bytes) / G;
  int rc = BWTM_OK;
  float merge_ms = 0.0f, interleave_ms = 0.0f, encode_ms = 0.0f;
  if(upstream_failed) { set_error("an earlier rank of the writer chain failed"); rc = BWTM_ERR_COMM; }
  else { rc = ensure_capacity(&out, out.origin + estimate + (estimate >> 2) + (1 << 20), out.origin, stream); }
  if(rc == BWTM_OK) { rc = injected_failure("writer", r); }
  if(rc == BWTM_OK && end > begin)
  {
    rc = merge_ranges<KeyT>(a, b, arrived, arrived_offsets, splitter[r], splitter[r + 1], b_lo, begin, end, options, &out, control.as<EncodeControl>(), false,
                            &merge_ms, &interleave_ms, &encode_ms, nullptr, stream);
  }
  if(rc != BWTM_OK && !upstream_failed)
  {
    EncodeControl poisoned; std::memset(&poisoned, 0, sizeof(poisoned)); poisoned.out_size = POISONED_STATE;
    cudaMemcpyAsync(control.ptr, &poisoned, sizeof(poisoned), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);
  }
  if(r < G - 1) { BWTM_NCCL(api->Send(control.ptr, sizeof(EncodeControl), ncclUint8, r + 1, comm->comm, stream)); }
  if(rc == BWTM_OK && r == G - 1)
  {
    SlabEncoder last; rc = last.init(4096, stream);
    if(rc == BWTM_OK) { rc = last.finish(&out, control.as<EncodeControl>(), stream); }
  }
  rc = agree_on_status(comm, rc, stream, "chained writer");
  if(rc != BWTM_OK) { device_free(out.ptr); return rc; }
  BWTM_CUDA(cudaMemcpyAsync(&ctl, control.ptr, sizeof(EncodeControl), cudaMemcpyDeviceToHost, stream));
  BWTM_CUDA(cudaStreamSynchronize(stream));
  timings->exchange_seconds += merge_ms * 1e-3;
  timings->interleave_seconds = interleave_ms * 1e-3; timings->encode_seconds = encode_ms * 1e-3;
  phase.mark("chained merge + writer");

  // 7. + 8. the complete result on every rank
  rc = gather_slices_and_index(comm, a, b, options, out, ctl, stream, result, timings);
  device_free(out.ptr);
  phase.mark("gather + index");
  return rc;
}

} // namespace bwtm

extern "C"
{

// Contiguous block of `total` items owned by `rank` of `world`: [first, first + count). The split of B's
// sequence ids across GPUs (the reference's ParallelLoop blocks, utils.cpp:169-197, with one block per rank).
int bwtm_shard_range(uint64_t total, uint32_t rank, uint32_t world, uint64_t* first, uint64_t* count)
{
  if(first == nullptr || count == nullptr || world == 0 || rank >= world) { set_error("invalid argument"); return BWTM_ERR_ARGUMENT; }
  uint64_t begin = (uint64_t)(((__uint128_t)total * rank) / world);
  uint64_t end = (uint64_t)(((__uint128_t)total * (rank + 1)) / world);
  *first = begin; *count = end - begin;
  return BWTM_OK;
}

int bwtm_comm_unique_id(uint8_t* id)
{
  if(id == nullptr) { set_error("null argument"); return BWTM_ERR_ARGUMENT; }
  NcclApi* api = nccl();
  if(!api->ok) { set_error("libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "symbol missing"); return BWTM_ERR_COMM; }
  static_assert(sizeof(ncclUniqueId) == BWTM_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId unique;
  BWTM_NCCL(api->GetUniqueId(&unique));
  std::memcpy(id, &unique, sizeof(unique));
  return BWTM_OK;
}

int bwtm_comm_create(const uint8_t* id, int rank, int world, bwtm_comm** out)
{
  if(id == nullptr || out == nullptr || world < 1 || rank < 0 || rank >= world) { set_error("invalid argument"); return BWTM_ERR_ARGUMENT; }
  *out = nullptr;
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    cudaGetLastError(); set_error("no CUDA device available; this library has no CPU fallback"); return BWTM_ERR_CUDA;
  }
  NcclApi* api = nccl();
  if(!api->ok) { set_error("libnccl.so.2 could not be loaded"); return BWTM_ERR_COMM; }
  ncclUniqueId unique;
  std::memcpy(&unique, id, sizeof(unique));
  ncclComm_t comm;
  BWTM_NCCL(api->CommInitRank(&comm, world, unique, rank));
  bwtm_comm* c = new bwtm_comm();
  c->comm = comm; c->rank = rank; c->world = world;
  c->peer_state = 0; c->d_flag = nullptr; c->d_status = nullptr;
  if(cudaMalloc(reinterpret_cast<void**>(&(c->d_flag)), 2 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(c->d_flag, 0, 2 * sizeof(unsigned long long)) != cudaSuccess)
  {
    cudaGetLastError(); api->CommDestroy(comm); delete c;
    set_error("cannot allocate the barrier word"); return BWTM_ERR_MEMORY;
  }
  c->d_status = reinterpret_cast<long long*>(c->d_flag + 1);
  c->ship_stream = nullptr; c->ship_ready = nullptr; c->ship_done = nullptr;
  if(cudaStreamCreateWithFlags(&(c->ship_stream), cudaStreamNonBlocking) != cudaSuccess ||
     cudaEventCreateWithFlags(&(c->ship_ready), cudaEventDisableTiming) != cudaSuccess ||
     cudaEventCreateWithFlags(&(c->ship_done), cudaEventDisableTiming) != cudaSuccess)
  {
    cudaGetLastError(); c->ship_stream = nullptr;   // planes are shipped on the main stream instead
  }
  *out = c;
  return BWTM_OK;
}

int bwtm_comm_destroy(bwtm_comm* comm)
{
  if(comm == nullptr) { return BWTM_OK; }
  // Call it on all ranks once the last merge has returned everywhere (like ncclCommDestroy): the windows
  // of this rank go away here.
  cudaDeviceSynchronize();
  window_close(comm, &(comm->key_window)); window_close(comm, &(comm->rle_window)); window_close(comm, &(comm->record_window));
  if(comm->key_window.local != nullptr) { cudaFree(comm->key_window.local); }
  if(comm->rle_window.local != nullptr) { cudaFree(comm->rle_window.local); }
  if(comm->record_window.local != nullptr) { cudaFree(comm->record_window.local); }
  if(comm->d_flag != nullptr) { cudaFree(comm->d_flag); }
  if(comm->ship_stream != nullptr) { cudaStreamDestroy(comm->ship_stream); }
  if(comm->ship_ready != nullptr) { cudaEventDestroy(comm->ship_ready); }
  if(comm->ship_done != nullptr) { cudaEventDestroy(comm->ship_done); }
  cudaGetLastError();
  if(nccl()->ok) { nccl()->CommDestroy(comm->comm); }
  delete comm;
  return BWTM_OK;
}

int bwtm_merge_distributed(bwtm_comm* comm, bwtm_index* a, bwtm_index* b, const bwtm_merge_options* options,
                           bwtm_index** out, bwtm_timings* timings)
{
  bwtm_merge_options defaults; std::memset(&defaults, 0, sizeof(defaults));
  if(options == nullptr) { options = &defaults; }
  bwtm_timings local; std::memset(&local, 0, sizeof(local));
  bool keep = (options->keep_inputs != 0);

  int rc = BWTM_OK;
  if(comm == nullptr || a == nullptr || b == nullptr || out == nullptr) { set_error("null argument"); rc = BWTM_ERR_ARGUMENT; }
  else if(a->d_records == nullptr || b->d_records == nullptr) { set_error("an input has no rank structure (it was built with skip_index)"); rc = BWTM_ERR_ARGUMENT; }
  else if(b->sequences == 0) { set_error("the inserted BWT has no sequences"); rc = BWTM_ERR_ARGUMENT; }
  if(rc == BWTM_OK)
  {
    *out = nullptr;
    uint64_t launches_before = bwtm_kernel_launches();
    auto start = std::chrono::steady_clock::now();
    // options.sequence_blocks > 1 (the same on every rank): the search in batches and the range-wise merge.
    uint64_t batches = options->sequence_blocks;
    if(const char* env = getenv("BWTM_SEQUENCE_BLOCKS")) { batches = strtoull(env, nullptr, 10); }
    batches = std::min<uint64_t>(batches, 64);
    const bool narrow = (a->size < 0xFFFFFFFFull && getenv("BWTM_FORCE_WIDE") == nullptr);
    if(batches > 1 && comm->world > 1)
    {
      rc = (narrow ? merge_distributed_batched<uint32_t>(comm, a, b, options, (int)batches, out, &local)
                   : merge_distributed_batched<uint64_t>(comm, a, b, options, (int)batches, out, &local));
    }
    else if(narrow) { rc = merge_distributed_impl<uint32_t>(comm, a, b, options, out, &local); }
    else { rc = merge_distributed_impl<uint64_t>(comm, a, b, options, out, &local); }
    cudaDeviceSynchronize();
    local.total_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    local.kernel_launches = bwtm_kernel_launches() - launches_before;
  }
  if(!keep) { index_free(a); index_free(b); }
  if(timings != nullptr) { *timings = local; }
  return rc;
}

} // extern "C"
