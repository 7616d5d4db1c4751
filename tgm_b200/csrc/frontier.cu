// Frontier compaction: indices of the non-padded neighbour slots, in order.
// Hop h+1 seeds are flatten(hop h) (reference tgm/hooks/neighbors/recency.py:141-143) and the
// non-padded subset is what DeduplicationHook keeps (tgm/hooks/dedup.py:44-48).  ONE pass over the
// ids: a persistent grid (2 CTAs per SM) of contiguous segments; a CTA streams its segment into
// ballot words kept in shared memory, the segment totals are exchanged once, and the indices are
// produced from the mask words -- 4 B read per slot + 8 B written per kept slot, no second read of
// the ids and no per-tile look-back chain (the round-2 tile kernel stood at 0.48 of the HBM peak:
// 32 tiles of 48 KB per L2 round trip is all a chained scan moves).
#include <mutex>

#include "common.cuh"

using namespace tgm;

// per-device scratch for the segment status words, allocated once and kept
namespace {
struct Scratch {
  unsigned long long *p = nullptr;
  unsigned tag = 0;                // tag of the last launch
  cudaStream_t last_stream = nullptr;
  cudaEvent_t last_done = nullptr;  // recorded after every launch: a call on ANOTHER stream waits for it
};
Scratch g_scratch[64];
std::mutex g_scratch_mu;
}  // namespace

namespace {

constexpr int kSegThreads = 512;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegCtasPerSm = 2;
constexpr int kQuadsInFlight = 4;                  // 512-byte warp loads per batch, two batches in flight
constexpr int kMaxWordsPerWarp = 1600;             // 16 warps x 1600 x 4 B = 100 KB of masks per CTA
constexpr int kMinWordsPerWarp = 8;

// segment status word: bits 63..42 = tag of the launch that wrote it (never 0), low 42 bits = kept
// slots of the segment (segment 0 adds the carry of earlier launches).  The words are not cleared
// between launches: a word counts as published when it carries this launch's tag.
constexpr int kStatusStride = 16;  // 64-bit words between two status words: one 128-byte line each,
                                   // so that the polling of a launch spreads over the L2 slices
constexpr int kTagShift = 42;
constexpr unsigned kTagMod = 1u << 21;
constexpr unsigned long long kValueMask = (1ull << kTagShift) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int4 ld_stream_i32x4(const int32_t *p) {
  int4 v;
  asm("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// ids of the 4 slots [i, i+4), padding beyond n (the ragged end of the input, unaligned inputs)
__device__ __forceinline__ int4 load_quad_checked(const int32_t *__restrict__ nid, int64_t i,
                                                  int64_t n) {
  int4 v;
  v.x = i < n ? __ldg(nid + i) : TGM_PADDED_NODE_ID;
  v.y = i + 1 < n ? __ldg(nid + i + 1) : TGM_PADDED_NODE_ID;
  v.z = i + 2 < n ? __ldg(nid + i + 2) : TGM_PADDED_NODE_ID;
  v.w = i + 3 < n ? __ldg(nid + i + 3) : TGM_PADDED_NODE_ID;
  return v;
}

// The ids are read ONCE and nothing but their validity bits is kept: a CTA owns one contiguous
// segment of the input (16 warps x `words` 32-slot words, `words` a multiple of 4).  Pass 1 streams
// it front to back as ONE stream per CTA (a lane loads 4 consecutive ids with one 128-bit load, the
// 16 warps take consecutive 512-byte groups, 4 such loads per warp and batch, the next batch issued
// before the current one is consumed: 64 KB in flight per CTA; one stream per WARP -- 4736 open
// streams -- measured 3.2 TB/s) and leaves one validity word per 32 slots in shared memory; the
// segment totals are exchanged through one status word per segment (ONE wait per CTA instead of
// one look-back per 4096-slot tile); in pass 2 every warp turns its run of mask words into the
// kept indices -- the output value is the slot's own position, so the ids are not needed again --
// with lanes along the 32 slots of a word, so a warp store covers consecutive output elements.
// status[s * kStatusStride] = status of segment s = CTA s: a CTA only waits on CTAs with smaller
// block indices, which the hardware dispatches first (the same reliance as CUB's look-back scan).
// `first` = first slot of this launch, `carry` = read *out_count as the output offset
// of the launch (inputs longer than one launch's mask capacity take consecutive launches).
template <bool VEC>
__global__ void __launch_bounds__(kSegThreads, kSegCtasPerSm)
frontier_compact_kernel(const int32_t *__restrict__ nid, int64_t first, int64_t n, int words,
                        int carry, unsigned tag, unsigned long long *__restrict__ status,
                        int64_t *__restrict__ out_idx,
                        int64_t *__restrict__ out_count) {
  extern __shared__ __align__(16) unsigned s_mask[];  // [kSegWarps * words]: one validity word per 32 slots
  __shared__ int s_wcnt[kSegWarps];
  __shared__ long long s_red[kSegWarps];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t seg = blockIdx.x;
  const int G = kSegWarps * (words >> 2);              // 128-slot groups of the segment
  const int64_t a0 = first + seg * int64_t(G) * 128;   // first slot of the segment
  const int64_t a = a0 + int64_t(warp) * words * 32;   // first slot of this warp's run of words
  unsigned *my = s_mask + warp * words;

  // ---- pass 1: ids -> mask words.  Lane l holds slots 4l..4l+3 of its group, so word q of the
  // group is the OR over lanes 8q..8q+7 of nibble << 4(l & 7) ----
  // groups [0, full) lie inside n and (VEC) start 16-byte aligned: plain 128-bit loads
  int full = 0;
  if (VEC && a0 < n) full = int(min(int64_t(G), (n - a0) >> 7));
  auto load = [&](int g) -> int4 {  // g is warp-uniform
    const int64_t i = a0 + int64_t(g) * 128 + lane * 4;
    if (g < full) return ld_stream_i32x4(nid + i);
    if (g < G) return load_quad_checked(nid, i, n);
    return make_int4(TGM_PADDED_NODE_ID, TGM_PADDED_NODE_ID, TGM_PADDED_NODE_ID, TGM_PADDED_NODE_ID);
  };
  {
    constexpr int kStep = kSegWarps * kQuadsInFlight;  // groups per CTA batch
    int4 cur[kQuadsInFlight], nxt[kQuadsInFlight];
#pragma unroll
    for (int u = 0; u < kQuadsInFlight; ++u) cur[u] = load(u * kSegWarps + warp);
    for (int gb = 0; gb < G; gb += kStep) {
#pragma unroll
      for (int u = 0; u < kQuadsInFlight; ++u) nxt[u] = load(gb + kStep + u * kSegWarps + warp);
#pragma unroll
      for (int u = 0; u < kQuadsInFlight; ++u) {
        const int g = gb + u * kSegWarps + warp;
        if (g < G) {
          const int4 v = cur[u];
          unsigned x = unsigned(v.x != TGM_PADDED_NODE_ID) | (unsigned(v.y != TGM_PADDED_NODE_ID) << 1) |
                       (unsigned(v.z != TGM_PADDED_NODE_ID) << 2) | (unsigned(v.w != TGM_PADDED_NODE_ID) << 3);
          x <<= (lane & 7) * 4;
          x |= __shfl_xor_sync(0xffffffffu, x, 1);
          x |= __shfl_xor_sync(0xffffffffu, x, 2);
          x |= __shfl_xor_sync(0xffffffffu, x, 4);
          if ((lane & 7) == 0) s_mask[g * 4 + (lane >> 3)] = x;
        }
        cur[u] = nxt[u];
      }
    }
  }
  __syncthreads();
  int cnt = 0;  // kept slots of this warp's run
  for (int i = lane; i < words; i += 32) cnt += __popc(my[i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_wcnt[warp] = cnt;
  __syncthreads();

  // ---- exchange: publish the segment total, add up the totals of every earlier segment ----
  int total = 0, woff = 0;
#pragma unroll
  for (int w = 0; w < kSegWarps; ++w) {
    const int c = s_wcnt[w];
    if (w < warp) woff += c;
    total += c;
  }
  // the count earlier launches left; read before this segment publishes, i.e. before the last
  // segment of this launch can overwrite it
  long long carried = 0;
  if (tid == 0) {
    if (seg == 0 && carry) carried = out_count[0];
    st_status(status + seg * kStatusStride,
              ((unsigned long long)tag << kTagShift) | (unsigned long long)(carried + total));
  }
  long long part = 0;
  for (int64_t j = tid; j < seg; j += kSegThreads) {
    unsigned long long sv;
    while (unsigned((sv = ld_status(status + j * kStatusStride)) >> kTagShift) != tag) __nanosleep(40);
    part += (long long)(sv & kValueMask);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    long long base = 0;
#pragma unroll
    for (int w = 0; w < kSegWarps; ++w) base += s_red[w];
    base += carried;
    s_base = base;
    if (seg == int64_t(gridDim.x) - 1) out_count[0] = base + total;
  }
  __syncthreads();

  // ---- pass 2: mask words -> indices.  Four words per (broadcast) 128-bit shared-memory read;
  // lane l owns slot l of every word; the running output offset is a warp-uniform popcount sum,
  // so no shuffle sits between a word and its store ----
  int64_t *o = out_idx + s_base + woff;
  const uint4 *my4 = reinterpret_cast<const uint4 *>(my);  // words % 4 == 0: 16-byte aligned
  const unsigned below = (1u << lane) - 1u;
  int64_t val = a + lane;
  int r = 0;
  auto emit = [&](unsigned m, int64_t v) {
    if ((m >> lane) & 1u) o[r + __popc(m & below)] = v;
    r += __popc(m);
  };
#pragma unroll 2
  for (int g = 0; g < (words >> 2); ++g, val += 128) {
    const uint4 m = my4[g];
    emit(m.x, val);
    emit(m.y, val + 32);
    emit(m.z, val + 64);
    emit(m.w, val + 96);
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
  int dev = 0;
  TGM_CUDA(cudaGetDevice(&dev));
  TGM_REQUIRE(dev >= 0 && dev < 64, "tgm_frontier_compact: unsupported device ordinal");
  const int max_ctas = kSmCount * kSegCtasPerSm;
  const bool vec = aligned16(nid);
  // one launch covers up to max_ctas x 16 warps x kMaxWordsPerWarp x 32 slots (2.4e8); longer
  // inputs take consecutive launches, each starting at the count the previous one left
  const int64_t per_launch = int64_t(max_ctas) * kSegWarps * kMaxWordsPerWarp * 32;
  // the scratch state (status words, launch tag) is per device and advanced under the lock, in
  // the order the launches are issued
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  Scratch &sc = g_scratch[dev];
  if (!sc.p) {
    const int smem = kSegWarps * kMaxWordsPerWarp * int(sizeof(unsigned));
    TGM_CUDA(cudaFuncSetAttribute(frontier_compact_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TGM_CUDA(cudaFuncSetAttribute(frontier_compact_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TGM_CUDA(cudaMalloc(&sc.p, size_t(max_ctas) * kStatusStride * sizeof(unsigned long long)));
    TGM_CUDA(cudaMemset(sc.p, 0, size_t(max_ctas) * kStatusStride * sizeof(unsigned long long)));
    TGM_CUDA(cudaEventCreateWithFlags(&sc.last_done, cudaEventDisableTiming));
    sc.tag = 0, sc.last_stream = st;
  }
  // launches share the status words, so they must not overlap: same-stream
  // calls are ordered already, a call on another stream first waits for the previous launch
  if (sc.last_stream != st) {
    TGM_CUDA(cudaStreamWaitEvent(st, sc.last_done, 0));
    sc.last_stream = st;
  }
  for (int64_t first = 0; first < n; first += per_launch) {
    const int64_t m = (n - first < per_launch) ? n - first : per_launch;
    int64_t ctas = (m + int64_t(kSegWarps) * kMinWordsPerWarp * 32 - 1) /
                   (int64_t(kSegWarps) * kMinWordsPerWarp * 32);
    if (ctas > max_ctas) ctas = max_ctas;
    const int64_t per_warp = (m + ctas * kSegWarps - 1) / (ctas * kSegWarps);
    const int words = int((per_warp + 127) / 128) * 4;
    if (++sc.tag >= kTagMod) {
      // the tag wraps (once per two million launches): clear the status words so that none
      // carries the new tag already
      TGM_CUDA(cudaMemsetAsync(sc.p, 0, size_t(max_ctas) * kStatusStride * sizeof(unsigned long long), st));
      sc.tag = 1;
    }
    const size_t smem = size_t(kSegWarps) * words * sizeof(unsigned);
    if (vec)
      frontier_compact_kernel<true><<<int(ctas), kSegThreads, smem, st>>>(
          nid, first, first + m, words, first > 0 ? 1 : 0, sc.tag, sc.p, out_idx,
          out_count);
    else
      frontier_compact_kernel<false><<<int(ctas), kSegThreads, smem, st>>>(
          nid, first, first + m, words, first > 0 ? 1 : 0, sc.tag, sc.p, out_idx,
          out_count);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "frontier launch", __FILE__, __LINE__);
  }
  TGM_CUDA(cudaEventRecord(sc.last_done, st));
  return TGM_OK;
}
