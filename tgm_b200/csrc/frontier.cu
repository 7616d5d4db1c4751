// Frontier compaction: indices of the non-padded neighbour slots, in order.
// Hop h+1 seeds are flatten(hop h) (reference tgm/hooks/neighbors/recency.py:141-143) and the
// non-padded subset is what DeduplicationHook keeps (tgm/hooks/dedup.py:44-48).  ONE pass over the
// ids: every CTA takes a tile ticket, ranks its slots from warp ballots, publishes its tile total
// and obtains its global offset by decoupled look-back over the totals of earlier tiles (chained
// scan), then scatters -- 4 B read + 8 B written per kept slot, no second read of the ids.
#include <mutex>

#include "common.cuh"

using namespace tgm;

// per-device scratch for the tile counts, grown on demand and kept (cudaMallocAsync's pool is
// trimmed at every synchronisation, which made each call pay a fresh allocation)
namespace {
struct Scratch {
  int64_t *p = nullptr;
  size_t cap = 0;
};
Scratch g_scratch[64];
std::mutex g_scratch_mu;
}  // namespace

namespace {

constexpr int kTileThreads = 256;
constexpr int kItemsPerThread = 16;
constexpr int kTile = kTileThreads * kItemsPerThread;  // 4096 slots per CTA
constexpr int kWarps = kTileThreads / 32;

// tile status word: bits 63..62 = 0 not ready | 1 tile total | 2 inclusive prefix; low 62 = value
constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagPrefix = 2ull << 62,
                             kValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Element i of a tile is handled by (iteration it, thread tid): i = it*256 + tid, so a warp's 32
// lanes cover 32 consecutive slots and the ballot bit order is the output order.
// ws[0] = ticket counter, ws[1 + t] = status of tile t; all zero at launch.  Tiles are claimed in
// ticket order, so every tile a CTA waits on is already running: the look-back cannot deadlock.
__global__ void __launch_bounds__(kTileThreads)
frontier_compact_kernel(const int32_t *__restrict__ nid, int64_t n, int64_t tiles,
                        unsigned long long *__restrict__ ws, int64_t *__restrict__ out_idx,
                        int64_t *__restrict__ out_count) {
  __shared__ int s_cnt[kItemsPerThread][kWarps];  // kept slots per (iteration, warp) chunk
  __shared__ int s_off[kItemsPerThread][kWarps];  // exclusive offset of the chunk inside the tile
  __shared__ long long s_tile;
  __shared__ long long s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (long long)atomicAdd(ws, 1ull);
  __syncthreads();
  const int64_t tile = s_tile;
  unsigned long long *status = ws + 1;
  const int64_t base = tile * kTile;
  unsigned masks[kItemsPerThread];
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int64_t i = base + it * kTileThreads + tid;
    const bool ok = i < n && __ldg(nid + i) != TGM_PADDED_NODE_ID;
    masks[it] = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_cnt[it][warp] = __popc(masks[it]);
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive scan of the 16 x 8 chunk counts in (iteration, warp) order: 4 per lane
    const int *cnt = &s_cnt[0][0];
    int *off = &s_off[0][0];
    int v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = cnt[lane * 4 + q];
      sum += v[q];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int run = incl - sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      off[lane * 4 + q] = run;
      run += v[q];
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // publish, then look back over earlier tiles, 32 at a time
    unsigned long long prefix = 0;
    if (tile == 0) {
      if (lane == 0) st_status(status, kFlagPrefix | (unsigned long long)total);
    } else {
      if (lane == 0) st_status(status + tile, kFlagAggregate | (unsigned long long)total);
      int64_t look = tile - 1;
      while (true) {
        const int64_t j = look - lane;
        unsigned long long sv = kFlagPrefix;  // tiles before 0: an empty inclusive prefix
        if (j >= 0) {
          do { sv = ld_status(status + j); } while ((sv >> 62) == 0);
        }
        const unsigned done = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
        // lanes up to and including the closest tile that already knows its inclusive prefix
        const int stop = done ? __ffs(done) - 1 : 31;
        unsigned long long part = lane <= stop ? (sv & kValueMask) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        prefix += part;
        if (done) break;
        look -= 32;
      }
      if (lane == 0)
        st_status(status + tile, kFlagPrefix | (prefix + (unsigned long long)total));
    }
    if (lane == 0) {
      s_prefix = (long long)prefix;
      if (tile == tiles - 1) out_count[0] = (long long)prefix + total;
    }
  }
  __syncthreads();
  const int64_t run = s_prefix;
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const unsigned m = masks[it];
    if (m & (1u << lane))
      out_idx[run + s_off[it][warp] + __popc(m & ((1u << lane) - 1u))] =
          base + it * kTileThreads + tid;
  }
}

}  // namespace

extern "C" int tgm_frontier_compact(const int32_t *nid, int64_t n, int64_t *out_idx,
                                    int64_t *out_count, tgm_stream stream) {
  TGM_REQUIRE(n >= 0, "tgm_frontier_compact: n must be >= 0");
  TGM_REQUIRE(out_count != nullptr, "tgm_frontier_compact: out_count is NULL");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    TGM_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st));
    return TGM_OK;
  }
  TGM_REQUIRE(nid && out_idx, "tgm_frontier_compact: NULL array argument");
  const int64_t tiles = (n + kTile - 1) / kTile;
  int dev = 0;
  TGM_CUDA(cudaGetDevice(&dev));
  TGM_REQUIRE(dev >= 0 && dev < 64, "tgm_frontier_compact: unsupported device ordinal");
  int64_t *tile_cnt = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    Scratch &sc = g_scratch[dev];
    if (size_t(tiles) + 1 > sc.cap) {
      TGM_CUDA(cudaStreamSynchronize(st));
      cudaFree(sc.p);
      sc.p = nullptr, sc.cap = 0;
      const size_t cap = size_t(tiles) * 2 + 1024;
      TGM_CUDA(cudaMalloc(&sc.p, cap * sizeof(int64_t)));
      sc.cap = cap;
    }
    tile_cnt = sc.p;
  }
  unsigned long long *ws = reinterpret_cast<unsigned long long *>(tile_cnt);
  TGM_CUDA(cudaMemsetAsync(ws, 0, size_t(tiles + 1) * sizeof(unsigned long long), st));
  frontier_compact_kernel<<<int(tiles), kTileThreads, 0, st>>>(nid, n, tiles, ws, out_idx, out_count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "frontier launch", __FILE__, __LINE__);
  return TGM_OK;
}
