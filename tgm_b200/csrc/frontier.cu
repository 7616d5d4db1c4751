// Frontier compaction: indices of the non-padded neighbour slots, in order.
// Hop h+1 seeds are flatten(hop h) (reference tgm/hooks/neighbors/recency.py:141-143) and the
// non-padded subset is what DeduplicationHook keeps (tgm/hooks/dedup.py:44-48).  Three launches:
// per-tile popcount of warp ballots, a single-CTA scan of the tile counts, and a stable scatter
// that re-derives each slot's rank from its warp ballot + tile prefix.
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

// Element i of a tile is handled by (iteration it, thread tid): i = it*256 + tid, so a warp's 32
// lanes cover 32 consecutive slots and the ballot bit order is the output order.
__global__ void __launch_bounds__(kTileThreads)
frontier_count_kernel(const int32_t *__restrict__ nid, int64_t n, int64_t *__restrict__ tile_cnt) {
  __shared__ int s_warp[kTileThreads / 32];
  const int64_t base = int64_t(blockIdx.x) * kTile;
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int64_t i = base + it * kTileThreads + threadIdx.x;
    const bool ok = i < n && nid[i] != TGM_PADDED_NODE_ID;
    cnt += __popc(__ballot_sync(0xffffffffu, ok));  // every lane ends with the warp total
  }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) tot += s_warp[w];
    tile_cnt[blockIdx.x] = tot;
  }
}

// exclusive scan of the tile counts in place (one CTA; tiles <= n/4096)
__global__ void __launch_bounds__(1024)
frontier_scan_kernel(int64_t *__restrict__ tile_cnt, int64_t tiles, int64_t *__restrict__ total) {
  __shared__ int64_t s_part[1024];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < tiles; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < tiles ? tile_cnt[i] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      int64_t add = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < tiles) tile_cnt[i] = s_carry + s_part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kTileThreads)
frontier_scatter_kernel(const int32_t *__restrict__ nid, int64_t n,
                        const int64_t *__restrict__ tile_off, int64_t *__restrict__ out_idx) {
  __shared__ int s_cnt[kItemsPerThread][kTileThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = int64_t(blockIdx.x) * kTile;
  unsigned masks[kItemsPerThread];
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    const int64_t i = base + it * kTileThreads + threadIdx.x;
    const bool ok = i < n && nid[i] != TGM_PADDED_NODE_ID;
    masks[it] = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_cnt[it][warp] = __popc(masks[it]);
  }
  __syncthreads();
  // rank of (it, warp) chunk inside the tile = all chunks of earlier iterations + earlier warps
  int64_t run = tile_off[blockIdx.x];
#pragma unroll
  for (int it = 0; it < kItemsPerThread; ++it) {
    int before = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) {
      const int c = s_cnt[it][w];
      if (w < warp) before += c;
    }
    int row_total = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) row_total += s_cnt[it][w];
    const unsigned m = masks[it];
    if (m & (1u << lane)) {
      const int64_t i = base + it * kTileThreads + threadIdx.x;
      out_idx[run + before + __popc(m & ((1u << lane) - 1u))] = i;
    }
    run += row_total;
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
    if (size_t(tiles) > sc.cap) {
      TGM_CUDA(cudaStreamSynchronize(st));
      cudaFree(sc.p);
      sc.p = nullptr, sc.cap = 0;
      const size_t cap = size_t(tiles) * 2 + 1024;
      TGM_CUDA(cudaMalloc(&sc.p, cap * sizeof(int64_t)));
      sc.cap = cap;
    }
    tile_cnt = sc.p;
  }
  frontier_count_kernel<<<int(tiles), kTileThreads, 0, st>>>(nid, n, tile_cnt);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) {
    frontier_scan_kernel<<<1, 1024, 0, st>>>(tile_cnt, tiles, out_count);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    frontier_scatter_kernel<<<int(tiles), kTileThreads, 0, st>>>(nid, n, tile_cnt, out_idx);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cuda_fail(e, "frontier launch", __FILE__, __LINE__);
  return TGM_OK;
}
