// Stateless recency sampler over a per-node chronological adjacency.
//
// Same answers as RecencyNeighborHook driven batch by batch (reference tgm-team/tgm @ 5183dc9,
// tgm/hooks/neighbors/recency.py:119-171, :239-321, :323-399), restated without mutable state
// (SURVEY.md Appendix A.2): the ring of node v at the time batch b queries it holds the last B
// entries of v's history restricted to batches < b, and that history, ordered the way the
// reference's stable sort appends it -- (batch, time, side, edge) -- is immutable.  So it is
// built once per loader geometry, and one launch then serves the seeds of thousands of loader
// batches: per seed an edge-index cut selects the visible prefix, the window is its last B
// entries, the time test and the right-aligned k-gather follow.
//
// HBM layout (E edges, n = 2E entries when undirected):
//   entries  Entry[n]      16 B each {nbr, eid, t}, grouped by node, chronological per node
//   rowptr   int64[N+1]
//   anchors  uint2[2][E]   per stream edge and endpoint: {first entry of the endpoint's list that
//                          belongs to the edge's own batch or later, #entries before it}
//                          -> hop-0 seeds that are edge endpoints need no search at all
//   xrows    float[n, D]   optional: feature rows in adjacency order (window = contiguous read)
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <new>

#include "store.cuh"
#include "tma.cuh"

using namespace tgm;

struct tgm_csr {
  int device = -1;
  const tgm_store *store = nullptr;  // borrowed: must outlive the csr
  int64_t e_start = 0, Ew = 0, bs = 1, n = 0;
  int32_t N = 0, D = 0;
  int directed = 0, colocate = 0;
  Entry *entries = nullptr;
  int64_t *rowptr = nullptr;
  uint2 *anchors = nullptr;  // [2][Ew]
  float *xrows = nullptr;    // [n, D] when colocate
  // dynamic chunk counters of the TMA kernel: launch i uses ticket[i % kTickets], so launches of
  // one handle that overlap on different streams never share a counter
  static constexpr int kTickets = 64;
  unsigned long long *ticket = nullptr;
  mutable unsigned launches = 0;
  // device staging of the host-buffer entry points: per slot one block per output kind
  // (nid, t, x, eid, mean), grown on demand
  struct Stage {
    void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0};
  } stage[TGM_HOST_SLOTS];
  ~tgm_csr() {
    if (device >= 0) {
      DeviceGuard g(device);
      cudaFree(entries);
      cudaFree(rowptr);
      cudaFree(anchors);
      cudaFree(xrows);
      cudaFree(ticket);
      for (auto &st : stage)
        for (void *q : st.p) cudaFree(q);
    }
  }
};

// 0: feature rows copied by the warp (LSU), 1: by the TMA unit (cp.async.bulk staging)
static int g_csr_feature_copy = 1;
static int g_csr_tma_ctas_per_sm = 0;
static int g_trace = 0;  // tgm_set_option("trace", 1): phase timings of tgm_csr_build on stderr  // 0 = as many as shared memory allows (<= TGM_FAST_MIN_BLOCKS)

namespace {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;

// ---- build -----------------------------------------------------------------------------------
// Position of every entry in the global (batch, time, side, edge) order.  Inside one batch the
// edges sharing a timestamp form a run [rs, re); its side-0 entries precede its side-1 entries.
__global__ void csr_keys_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                const int64_t *__restrict__ t, int64_t Ew, int64_t bs, int directed,
                                uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (int64_t l = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; l < Ew;
       l += int64_t(gridDim.x) * blockDim.x) {
    if (directed) {
      keys[l] = uint32_t(src[l]);
      vals[l] = uint32_t(l) << 1;
      continue;
    }
    const int64_t bl = (l / bs) * bs;
    const int64_t bh = bl + bs < Ew ? bl + bs : Ew;
    const int64_t te = t[l];
    int64_t lo = bl, hi = l;  // first index in [bl, l] with t == te
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (t[mid] < te) lo = mid + 1; else hi = mid;
    }
    const int64_t rs = lo;
    lo = l + 1, hi = bh;  // first index in (l, bh] with t > te
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (t[mid] <= te) lo = mid + 1; else hi = mid;
    }
    const int64_t re = lo;
    const int64_t p0 = 2 * rs + (l - rs);
    const int64_t p1 = 2 * rs + (re - rs) + (l - rs);
    keys[p0] = uint32_t(src[l]);
    vals[p0] = uint32_t(l) << 1;
    keys[p1] = uint32_t(dst[l]);
    vals[p1] = (uint32_t(l) << 1) | 1u;
  }
}

__global__ void csr_rowptr_kernel(const uint32_t *__restrict__ keys, int64_t n, int32_t N,
                                  int64_t *__restrict__ rowptr) {
  for (int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; v <= N;
       v += int64_t(gridDim.x) * blockDim.x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < uint32_t(v)) lo = mid + 1; else hi = mid;
    }
    rowptr[v] = lo;
  }
}

// Run heads of the sorted list: position j starts a new (node, batch) run.  hp[j] = j for heads,
// 0 otherwise; an inclusive max-scan then gives every position the start of its run -- which is
// the anchor of every edge endpoint in that run (first entry of the node that belongs to the
// edge's own batch), with no search.
__global__ void csr_heads_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                 int64_t n, uint32_t bs, uint32_t *__restrict__ hp) {
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += int64_t(gridDim.x) * blockDim.x) {
    bool head = j == 0;
    if (!head)
      head = keys[j] != keys[j - 1] || (vals[j] >> 1) / bs != (vals[j - 1] >> 1) / bs;
    hp[j] = head ? uint32_t(j) : 0u;
  }
}

struct MaxU32 {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// entries in adjacency order + (optionally) the anchor table, scattered back to stream order
__global__ void csr_entries_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                   const uint32_t *__restrict__ run_start, int64_t n,
                                   const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                   const int64_t *__restrict__ t, const int64_t *__restrict__ rowptr,
                                   int64_t e_start, int64_t Ew, Entry *__restrict__ entries,
                                   uint2 *__restrict__ anchors) {
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t val = vals[j];
    const int64_t l = val >> 1;
    Entry en;
    en.nbr = (val & 1u) ? src[l] : dst[l];
    en.eid = int32_t(e_start + l);
    en.t = t[l];
    entries[j] = en;
    if (anchors) {
      const uint32_t rs = run_start[j];
      anchors[(val & 1u) ? Ew + l : l] = make_uint2(rs, rs - uint32_t(__ldg(rowptr + keys[j])));
    }
  }
}

// first index in [lo, hi) whose entry belongs to an edge >= cut (monotone: batch-major order)
__device__ __forceinline__ int64_t lower_bound_eid(const Entry *__restrict__ entries, int64_t lo,
                                                   int64_t hi, int64_t cut) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(entries[mid].eid) < cut) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// search-based anchors (directed adjacencies: a dst endpoint owns no entry in its batch's run)
__global__ void csr_anchor_kernel(const Entry *__restrict__ entries,
                                  const int64_t *__restrict__ rowptr,
                                  const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                  int64_t Ew, int64_t bs, int64_t e_start, int32_t N,
                                  uint2 *__restrict__ anchors) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * Ew;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int side = i >= Ew;
    const int64_t l = side ? i - Ew : i;
    const int32_t v = side ? dst[l] : src[l];
    uint2 a = make_uint2(0u, 0u);
    if (v >= 0 && v < N) {
      const int64_t lo = rowptr[v], hi = rowptr[v + 1];
      const int64_t cut = e_start + (l / bs) * bs;
      const int64_t pos = lower_bound_eid(entries, lo, hi, cut);
      a = make_uint2(uint32_t(pos), uint32_t(pos - lo));
    }
    anchors[i] = a;
  }
}

template <bool VEC4>
__global__ void csr_gather_x_kernel(const Entry *__restrict__ entries, int64_t n,
                                    const float *__restrict__ x, int D, float *__restrict__ xrows) {
  if (VEC4) {
    const int D4 = D >> 2;
    const int64_t total = n * D4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *o4 = reinterpret_cast<float4 *>(xrows);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t j = i / D4;
      const int d = int(i - j * D4);
      o4[i] = __ldg(x4 + int64_t(entries[j].eid) * D4 + d);
    }
  } else {
    const int64_t total = n * D;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t j = i / D;
      const int d = int(i - j * D);
      xrows[i] = __ldg(x + int64_t(entries[j].eid) * D + d);
    }
  }
}

// ---- sample ----------------------------------------------------------------------------------
constexpr int kSampleThreads = 256;

// The part shared by both entry points: given the visible window [wstart, wstart + nwin) of a
// seed's adjacency, apply the time test and emit the right-aligned k-gather (recency.py:267-319).
template <bool VEC4, bool COLOC>
__device__ __forceinline__ void emit_window(const Entry *__restrict__ entries,
                                            const float *__restrict__ xsrc, int D, int64_t wstart,
                                            int nwin, int64_t q, int k, int64_t s,
                                            int32_t *__restrict__ out_nid,
                                            int64_t *__restrict__ out_t, float *__restrict__ out_x,
                                            uint32_t *my, int lane) {
  int last = -1;  // right-most window position with time < tq
  for (int base = 0; base < nwin; base += 32) {
    const int j = base + lane;
    const bool ok = (j < nwin) && (entries[wstart + j].t < q);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) last = base + 31 - __clz(m);
  }
  for (int c = lane; c < k; c += 32) {
    const int p = last - (k - 1 - c);
    int32_t id = TGM_PADDED_NODE_ID;
    int64_t tt = 0;
    uint32_t row = kNoRow;
    if (p >= 0) {
      const Entry en = entries[wstart + p];
      id = en.nbr;
      tt = en.t;
      row = COLOC ? uint32_t(wstart + p) : uint32_t(en.eid);
    }
    my[c] = row;
    out_nid[s * k + c] = id;
    out_t[s * k + c] = tt;
  }
  __syncwarp();
  if (D > 0) {
    if (VEC4) {
      const int D4 = D >> 2;
      const float4 *x4 = reinterpret_cast<const float4 *>(xsrc);
      float4 *o4 = reinterpret_cast<float4 *>(out_x) + s * int64_t(k) * D4;
      const int total = k * D4;
      for (int i = lane; i < total; i += 32) {
        const int c = i / D4, d = i - c * D4;
        const uint32_t row = my[c];
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row != kNoRow) val = ldg_stream_f4(x4 + int64_t(row) * D4 + d);
        stg_stream_f4(o4 + i, val);
      }
    } else {
      float *o = out_x + s * int64_t(k) * D;
      const int total = k * D;
      for (int i = lane; i < total; i += 32) {
        const int c = i / D, d = i - c * D;
        const uint32_t row = my[c];
        o[i] = row != kNoRow ? __ldg(xsrc + int64_t(row) * D + d) : 0.f;
      }
    }
  }
  __syncwarp();
}

template <bool VEC4, bool COLOC>
__global__ void __launch_bounds__(kSampleThreads)
csr_sample_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                  const float *__restrict__ xsrc, int32_t N, int D,
                  const int32_t *__restrict__ seeds, const int64_t *__restrict__ tq,
                  const int64_t *__restrict__ cut, int64_t cut_group, int64_t S, int B, int k,
                  int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                  float *__restrict__ out_x) {
  extern __shared__ uint32_t s_rows[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  uint32_t *my = s_rows + warp * k;
  for (int64_t s = int64_t(blockIdx.x) * wpb + warp; s < S; s += int64_t(gridDim.x) * wpb) {
    const int32_t v = seeds[s];
    int64_t wstart = 0;
    int nwin = 0;
    if (v >= 0 && v < N) {  // a padded seed (-1) always yields an all-padding row
      const int64_t lo = rowptr[v], hi = rowptr[v + 1];
      const int64_t pos = lower_bound_eid(entries, lo, hi, cut[s / cut_group]);
      wstart = pos - B > lo ? pos - B : lo;
      nwin = int(pos - wstart);
    }
    emit_window<VEC4, COLOC>(entries, xsrc, D, wstart, nwin, tq[s], k, s, out_nid, out_t, out_x,
                             my, lane);
  }
}

template <bool VEC4, bool COLOC>
__global__ void __launch_bounds__(kSampleThreads)
csr_sample_edges_kernel(const Entry *__restrict__ entries, const uint2 *__restrict__ anchors,
                        const float *__restrict__ xsrc, const int64_t *__restrict__ t, int D,
                        int64_t Ew, int64_t bs, int64_t l_lo, int64_t l_hi, int B, int k,
                        int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                        float *__restrict__ out_x) {
  extern __shared__ uint32_t s_rows[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  uint32_t *my = s_rows + warp * k;
  const int64_t S = 2 * (l_hi - l_lo);
  for (int64_t s = int64_t(blockIdx.x) * wpb + warp; s < S; s += int64_t(gridDim.x) * wpb) {
    // row s -> (edge, endpoint): batch jb owns rows [2*bs*jb, ...), src seeds then dst seeds
    const int64_t jb = s / (2 * bs);
    const int64_t bstart = l_lo + jb * bs;
    const int64_t nb = l_hi - bstart < bs ? l_hi - bstart : bs;
    const int64_t rr = s - jb * 2 * bs;
    const int side = rr >= nb;
    const int64_t l = bstart + (side ? rr - nb : rr);
    const uint2 a = anchors[side ? Ew + l : l];
    const int nwin = a.y < uint32_t(B) ? int(a.y) : B;
    const int64_t wstart = int64_t(a.x) - nwin;
    emit_window<VEC4, COLOC>(entries, xsrc, D, wstart, nwin, t[l], k, s, out_nid, out_t, out_x, my,
                             lane);
  }
}

// ---- fast path: B <= 32, features colocated (or absent) ---------------------------------------
// A warp owns 32 consecutive seeds.  Prologue, lane-parallel: lane i resolves seed i's window
// (one coalesced anchor/time load, or a private binary search for general seeds).  Then the warp
// walks its 32 seeds: lane j holds window entry j (ONE 128-bit load per lane, issued one seed
// ahead), a ballot finds the right-most entry with t < tq, shuffles route entries to their
// output columns, and because the features are colocated the k-gather is a single contiguous
// copy of nvalid*D floats -- no per-row indexing, no shared memory.
constexpr int kFastThreads = 256;
#ifndef TGM_FAST_MIN_BLOCKS
#define TGM_FAST_MIN_BLOCKS 6
#endif

struct SeedWin {
  int64_t wstart;  // first visible-window entry
  int nwin;        // window length (<= B <= 32)
  int64_t q;       // query time
};

__device__ __forceinline__ Entry load_window_entry(const Entry *__restrict__ entries,
                                                   int64_t wstart, int nwin, int lane) {
  Entry e;
  e.nbr = TGM_PADDED_NODE_ID;
  e.eid = 0;
  e.t = 0;
  if (lane < nwin) e = ldg_stream_entry(entries + wstart + lane);
  return e;
}

__device__ __forceinline__ void emit_fast(const float4 *__restrict__ x4, int D4, int64_t wstart,
                                          int nwin, int64_t q, int k, int64_t s, const Entry &cur,
                                          int32_t *__restrict__ out_nid,
                                          int64_t *__restrict__ out_t, float4 *__restrict__ out_x4,
                                          int lane, int32_t *__restrict__ out_eid = nullptr) {
  const unsigned m = __ballot_sync(0xffffffffu, lane < nwin && cur.t < q);
  const int last = m ? 31 - __clz(m) : -1;  // recency.py:267-281
  const int nvalid = last + 1 < k ? last + 1 : k;
  const int first = last + 1 - nvalid, pad = k - nvalid;
  const int srcl = (first + lane - pad) & 31;
  const int32_t nbr = __shfl_sync(0xffffffffu, cur.nbr, srcl);
  const int64_t tt = shfl_i64(cur.t, srcl);
  const int32_t eid = __shfl_sync(0xffffffffu, cur.eid, srcl);
  if (lane < k) {  // right-aligned, left-padded with (-1, 0) (:287-319)
    const bool v = lane >= pad;
    if (out_nid) out_nid[s * k + lane] = v ? nbr : TGM_PADDED_NODE_ID;
    if (out_t) out_t[s * k + lane] = v ? tt : 0;
    if (out_eid) out_eid[s * k + lane] = v ? eid : -1;
  }
  if (D4 > 0) {
    float4 *o4 = out_x4 + s * int64_t(k) * D4;
    const int npad4 = pad * D4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < npad4; i += 32) stg_stream_f4(o4 + i, zero);
    o4 += npad4;
    const float4 *src = x4 + (wstart + first) * D4;
    const int n4 = nvalid * D4;
    int i = lane;
    for (; i + 96 < n4; i += 128) {
      const float4 a = ldg_stream_f4(src + i), b = ldg_stream_f4(src + i + 32),
                   c = ldg_stream_f4(src + i + 64), d = ldg_stream_f4(src + i + 96);
      stg_stream_f4(o4 + i, a);
      stg_stream_f4(o4 + i + 32, b);
      stg_stream_f4(o4 + i + 64, c);
      stg_stream_f4(o4 + i + 96, d);
    }
    if (i + 32 < n4) {
      const float4 a = ldg_stream_f4(src + i), b = ldg_stream_f4(src + i + 32);
      stg_stream_f4(o4 + i, a);
      stg_stream_f4(o4 + i + 32, b);
      i += 64;
    }
    for (; i < n4; i += 32) stg_stream_f4(o4 + i, ldg_stream_f4(src + i));
  }
}

// Chunks of 32 seeds are handed out dynamically: the first by warp id, the rest from a ticket
// counter (windows differ in length; a static split leaves a ~1/iterations tail).
__device__ __forceinline__ int64_t next_chunk(unsigned long long *ticket, int64_t wstride,
                                              int lane) {
  unsigned long long nx = 0;
  if (lane == 0) nx = atomicAdd(ticket, 1ull);
  return wstride + int64_t(__shfl_sync(0xffffffffu, (unsigned int)(nx & 0xffffffffu), 0)) +
         (int64_t(__shfl_sync(0xffffffffu, (unsigned int)(nx >> 32), 0)) << 32);
}

// walk the (up to) 32 seeds whose windows the lanes resolved
__device__ __forceinline__ void walk_chunk(const Entry *__restrict__ entries,
                                           const float4 *__restrict__ x4, int D4,
                                           const SeedWin &mine, int64_t s_base, int nseeds, int k,
                                           int32_t *__restrict__ out_nid,
                                           int64_t *__restrict__ out_t,
                                           float4 *__restrict__ out_x4, int lane,
                                           int32_t *__restrict__ out_eid = nullptr) {
  int64_t wstart = shfl_i64(mine.wstart, 0);
  int nwin = __shfl_sync(0xffffffffu, mine.nwin, 0);
  Entry cur = load_window_entry(entries, wstart, nwin, lane);
  for (int i = 0; i < nseeds; ++i) {
    const int64_t q = shfl_i64(mine.q, i);
    const int nxt = i + 1 < nseeds ? i + 1 : i;
    const int64_t wstart_n = shfl_i64(mine.wstart, nxt);
    const int nwin_n = __shfl_sync(0xffffffffu, mine.nwin, nxt);
    const Entry ahead = load_window_entry(entries, wstart_n, nwin_n, lane);  // one seed ahead
    emit_fast(x4, D4, wstart, nwin, q, k, s_base + i, cur, out_nid, out_t, out_x4, lane, out_eid);
    cur = ahead;
    wstart = wstart_n;
    nwin = nwin_n;
  }
}

__global__ void __launch_bounds__(kFastThreads, TGM_FAST_MIN_BLOCKS)
csr_sample_edges_fast_kernel(const Entry *__restrict__ entries, const uint2 *__restrict__ anchors,
                             const float4 *__restrict__ x4, const int64_t *__restrict__ t, int D4,
                             int64_t Ew, uint32_t bs, int64_t l_lo, int64_t l_hi, int B, int k,
                             int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                             float4 *__restrict__ out_x4, unsigned long long *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  const int64_t S = 2 * (l_hi - l_lo);
  const int64_t nchunks = (S + 31) >> 5;
  const int64_t wstride = int64_t(gridDim.x) * (kFastThreads >> 5);
  for (int64_t ch = int64_t(blockIdx.x) * (kFastThreads >> 5) + (threadIdx.x >> 5); ch < nchunks;
       ch = next_chunk(ticket, wstride, lane)) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    SeedWin mine{0, 0, 0};
    if (s < S) {
      // row s -> (edge, endpoint): batch jb owns rows [2*bs*jb, ...), src seeds then dst seeds
      const int64_t jb = s / (2 * int64_t(bs));
      const int64_t bstart = l_lo + jb * bs;
      const int64_t nb = l_hi - bstart < int64_t(bs) ? l_hi - bstart : int64_t(bs);
      const int64_t rr = s - jb * 2 * int64_t(bs);
      const bool side = rr >= nb;
      const int64_t l = bstart + (side ? rr - nb : rr);
      const uint2 a = __ldg(anchors + (side ? Ew + l : l));
      mine.nwin = a.y < uint32_t(B) ? int(a.y) : B;
      mine.wstart = int64_t(a.x) - mine.nwin;
      mine.q = __ldg(t + l);
    }
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    walk_chunk(entries, x4, D4, mine, s_base, nseeds, k, out_nid, out_t, out_x4, lane);
  }
}

__global__ void __launch_bounds__(kFastThreads, TGM_FAST_MIN_BLOCKS)
csr_sample_fast_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                       const float4 *__restrict__ x4, int32_t N, int D4,
                       const int32_t *__restrict__ seeds, const int64_t *__restrict__ tq,
                       const int64_t *__restrict__ cut, int64_t cut_group, int64_t S, int B, int k,
                       int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                       float4 *__restrict__ out_x4, unsigned long long *__restrict__ ticket,
                       int32_t *__restrict__ out_eid) {
  const int lane = threadIdx.x & 31;
  const int64_t nchunks = (S + 31) >> 5;
  const int64_t wstride = int64_t(gridDim.x) * (kFastThreads >> 5);
  for (int64_t ch = int64_t(blockIdx.x) * (kFastThreads >> 5) + (threadIdx.x >> 5); ch < nchunks;
       ch = next_chunk(ticket, wstride, lane)) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    SeedWin mine{0, 0, 0};
    if (s < S) {
      const int32_t v = __ldg(seeds + s);
      mine.q = __ldg(tq + s);
      if (v >= 0 && v < N) {  // a padded seed (-1) always yields an all-padding row
        const int64_t lo = __ldg(rowptr + v), hi = __ldg(rowptr + v + 1);
        const int64_t pos = lower_bound_eid(entries, lo, hi, __ldg(cut + s / cut_group));
        mine.wstart = pos - B > lo ? pos - B : lo;
        mine.nwin = int(pos - mine.wstart);
      }
    }
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    walk_chunk(entries, x4, D4, mine, s_base, nseeds, k, out_nid, out_t, out_x4, lane, out_eid);
  }
}

// ---- hop-0 forms that never materialise the (S, k, D) feature block ---------------------------
// Where a seed's window comes from: the prebuilt anchor table, or (SEARCH) the seed is read from
// the store's src/dst slab and its history cut is found by a private binary search -- the form the
// host-buffer entry points use, so that the slab they upload is what the kernel consumes.
struct EdgeSeedArgs {
  const uint2 *anchors;
  const int64_t *rowptr;
  const int32_t *src, *dst;  // store slabs, offset to the adjacency's e_start
  const int64_t *t;
  int64_t Ew, l_lo, l_hi, e_start;
  uint32_t bs;
  int32_t N;
};

template <bool SEARCH>
__device__ __forceinline__ SeedWin resolve_edge_seed(const Entry *__restrict__ entries,
                                                     const EdgeSeedArgs &a, int64_t s, int B) {
  // row s -> (edge, endpoint): batch jb owns rows [2*bs*jb, ...), src seeds then dst seeds
  const int64_t bs = a.bs;
  const int64_t jb = s / (2 * bs);
  const int64_t bstart = a.l_lo + jb * bs;
  const int64_t nb = a.l_hi - bstart < bs ? a.l_hi - bstart : bs;
  const int64_t rr = s - jb * 2 * bs;
  const bool side = rr >= nb;
  const int64_t l = bstart + (side ? rr - nb : rr);
  SeedWin w{0, 0, __ldg(a.t + l)};
  if (SEARCH) {
    const int32_t v = side ? __ldg(a.dst + l) : __ldg(a.src + l);
    if (v >= 0 && v < a.N) {
      const int64_t lo = __ldg(a.rowptr + v), hi = __ldg(a.rowptr + v + 1);
      const int64_t pos = lower_bound_eid(entries, lo, hi, a.e_start + bstart);
      w.wstart = pos - B > lo ? pos - B : lo;
      w.nwin = int(pos - w.wstart);
    }
  } else {
    const uint2 an = __ldg(a.anchors + (side ? a.Ew + l : l));
    w.nwin = an.y < uint32_t(B) ? int(an.y) : B;
    w.wstart = int64_t(an.x) - w.nwin;
  }
  return w;
}

// ids only: (nid, t, eid) per slot, 16 bytes -- what a caller that owns the feature table lacks
constexpr int kIdsMinBlocks = 4;  // 64 registers: no spills in the walk / mean loops

template <bool SEARCH>
__global__ void __launch_bounds__(kFastThreads, kIdsMinBlocks)
csr_sample_edges_ids_kernel(const Entry *__restrict__ entries, const EdgeSeedArgs a, int B, int k,
                            int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                            int32_t *__restrict__ out_eid, unsigned long long *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  const int64_t S = 2 * (a.l_hi - a.l_lo);
  const int64_t nchunks = (S + 31) >> 5;
  const int64_t wstride = int64_t(gridDim.x) * (kFastThreads >> 5);
  for (int64_t ch = int64_t(blockIdx.x) * (kFastThreads >> 5) + (threadIdx.x >> 5); ch < nchunks;
       ch = next_chunk(ticket, wstride, lane)) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    SeedWin mine{0, 0, 0};
    if (s < S) mine = resolve_edge_seed<SEARCH>(entries, a, s, B);
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    int64_t wstart = shfl_i64(mine.wstart, 0);
    int nwin = __shfl_sync(0xffffffffu, mine.nwin, 0);
    Entry cur = load_window_entry(entries, wstart, nwin, lane);
    for (int i = 0; i < nseeds; ++i) {
      const int64_t q = shfl_i64(mine.q, i);
      const int nxt = i + 1 < nseeds ? i + 1 : i;
      const int64_t wstart_n = shfl_i64(mine.wstart, nxt);
      const int nwin_n = __shfl_sync(0xffffffffu, mine.nwin, nxt);
      const Entry ahead = load_window_entry(entries, wstart_n, nwin_n, lane);  // one seed ahead
      emit_fast(nullptr, 0, wstart, nwin, q, k, s_base + i, cur, out_nid, out_t, nullptr, lane,
                out_eid);
      cur = ahead;
      wstart = wstart_n;
      nwin = nwin_n;
    }
  }
}

// Fused sample + masked mean (examples/linkproppred/graphmixer.py:131-135 over the rows
// _get_recency_neighbors returns, recency.py:239-321): out_mean[s, :] = sum of the seed's valid
// feature rows / max(1, #valid), accumulated oldest to newest in fp32 -- bit-identical to
// tgm_masked_mean over tgm_csr_sample_edges' output, without the (S, k, D) block ever existing.
// Phase 1 walks the 32 seeds of a chunk like the kernels above (ids/times/eids are optional
// outputs); lane i keeps seed i's first valid row and count.  Phase 2: groups of G lanes (G = the
// power of two >= D/4, <= 32) take one seed each, lane c of a group owns float4 column c and adds
// the rows in order; the row loads are independent, so several are in flight per lane.
template <bool SEARCH>
__global__ void __launch_bounds__(kFastThreads, kIdsMinBlocks)
csr_sample_edges_mean_kernel(const Entry *__restrict__ entries, const EdgeSeedArgs a,
                             const float4 *__restrict__ x4, int D4, int G, int B, int k,
                             int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                             int32_t *__restrict__ out_eid, float4 *__restrict__ out_mean4,
                             unsigned long long *__restrict__ ticket) {
  const int lane = threadIdx.x & 31;
  const int64_t S = 2 * (a.l_hi - a.l_lo);
  const int64_t nchunks = (S + 31) >> 5;
  const int64_t wstride = int64_t(gridDim.x) * (kFastThreads >> 5);
  const int per = 32 / G, sub = lane / G, col = lane - sub * G;
  const bool want_ids = out_nid || out_t || out_eid;
  for (int64_t ch = int64_t(blockIdx.x) * (kFastThreads >> 5) + (threadIdx.x >> 5); ch < nchunks;
       ch = next_chunk(ticket, wstride, lane)) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    SeedWin mine{0, 0, 0};
    if (s < S) mine = resolve_edge_seed<SEARCH>(entries, a, s, B);
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    int64_t wstart = shfl_i64(mine.wstart, 0);
    int nwin = __shfl_sync(0xffffffffu, mine.nwin, 0);
    Entry cur = load_window_entry(entries, wstart, nwin, lane);
    int64_t row0 = 0;  // lane i: first valid feature row of seed i
    int nv = 0;        //         and how many follow
    for (int i = 0; i < nseeds; ++i) {
      const int64_t q = shfl_i64(mine.q, i);
      const int nxt = i + 1 < nseeds ? i + 1 : i;
      const int64_t wstart_n = shfl_i64(mine.wstart, nxt);
      const int nwin_n = __shfl_sync(0xffffffffu, mine.nwin, nxt);
      const Entry ahead = load_window_entry(entries, wstart_n, nwin_n, lane);
      const unsigned m = __ballot_sync(0xffffffffu, lane < nwin && cur.t < q);
      const int last = m ? 31 - __clz(m) : -1;
      const int nvalid = last + 1 < k ? last + 1 : k;
      const int first = last + 1 - nvalid;
      if (lane == i) {
        row0 = wstart + first;
        nv = nvalid;
      }
      if (want_ids)
        emit_fast(nullptr, 0, wstart, nwin, q, k, s_base + i, cur, out_nid, out_t, nullptr, lane,
                  out_eid);
      cur = ahead;
      wstart = wstart_n;
      nwin = nwin_n;
    }
    for (int g0 = 0; g0 < nseeds; g0 += per) {
      const int si = g0 + sub;
      const int64_t r0 = shfl_i64(row0, si & 31);
      const int n = __shfl_sync(0xffffffffu, nv, si & 31);
      if (si >= nseeds) continue;
      const float den = float(n > 1 ? n : 1);
      for (int c = col; c < D4; c += G) {
        const float4 *p = x4 + r0 * D4 + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
          const float4 v = ldg_stream_f4(p + int64_t(r) * D4);
          acc.x = __fadd_rn(acc.x, v.x);
          acc.y = __fadd_rn(acc.y, v.y);
          acc.z = __fadd_rn(acc.z, v.z);
          acc.w = __fadd_rn(acc.w, v.w);
        }
        out_mean4[(s_base + si) * D4 + c] =
            make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
      }
    }
  }
}

// ---- ring export: the stateful sampler's state after the edges [e_start, e_cut) were pushed ------
// One warp per node: the last min(count, B) visible entries land in the slots sequential pushes
// would have used (write_pos = number of pushes), so a stateful RecencyNeighborHook can take over
// from a windowed run (train epoch pre-sampled, validation stream driven batch by batch).
__global__ void __launch_bounds__(256)
csr_export_ring_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                       const float *__restrict__ x, int32_t N, int D, int64_t e_cut, int B,
                       int32_t *__restrict__ ids, int64_t *__restrict__ times,
                       float *__restrict__ feats, int32_t *__restrict__ wpos) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t v = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); v < N;
       v += int64_t(gridDim.x) * wpb) {
    const int64_t lo = rowptr[v], hi = rowptr[v + 1];
    const int64_t pos = lower_bound_eid(entries, lo, hi, e_cut);
    const int64_t cnt = pos - lo;
    const int nkeep = cnt < B ? int(cnt) : B;
    for (int j = lane; j < B; j += 32) {  // unrolled position j (oldest .. newest) of the ring
      const int64_t slot = (cnt + j) % B;  // == (write_pos - B + j) mod B with write_pos = cnt
      const int64_t src = pos - B + j;     // entry feeding it
      const bool has = j >= B - nkeep;
      Entry en;
      if (has) en = entries[src];
      ids[v * B + slot] = has ? en.nbr : TGM_PADDED_NODE_ID;
      times[v * B + slot] = has ? en.t : 0;
    }
    if (D > 0) {
      for (int i = lane; i < B * D; i += 32) {
        const int j = i / D, d = i - j * D;
        const int64_t slot = (cnt + j) % B;
        const bool has = j >= B - nkeep;
        float val = 0.f;
        if (has) val = __ldg(x + int64_t(entries[pos - B + j].eid) * D + d);
        feats[(v * B + slot) * D + d] = val;
      }
    }
    if (lane == 0) wpos[v] = int32_t(cnt);
  }
}

// ---- uniform full-history sampler (array_backend.py:108-171 via uniform.py:87-142) -------------
// Needs the adjacency built with batch_size 1: per node the entries are then ordered (edge, side),
// which is the order the reference appends candidates in (:132-137).  Candidates of a seed are its
// entries with e_lo <= edge < e_hi; up to k of them are returned left-aligned / right-padded
// (:155-169).  When there are more than k the reference draws k with CPython's random.sample
// (:152-153), which no device path can reproduce: here a k-subset is drawn uniformly with Floyd's
// algorithm from a counter-based generator keyed by (rng_seed, node id) -- like the reference, all
// occurrences of a node in one call share the same draw (:119,:166).
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr int kUniformThreads = 256;

// first index in [lo, hi) whose entry has time >= t (a node's list is in edge order = time order)
__device__ __forceinline__ int64_t lower_bound_time(const Entry *__restrict__ entries, int64_t lo,
                                                    int64_t hi, int64_t t) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (entries[mid].t < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// BY_TIME: the candidate range is given as a closed time interval [e_lo, e_hi] instead of the edge
// index range [e_lo, e_hi) -- the same candidates for a time-sorted stream (e < upper_bound(t, T)
// <=> t[e] <= T), without resolving the slice to edge indices first
template <bool BY_TIME>
__global__ void __launch_bounds__(kUniformThreads)
csr_uniform_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                   const float *__restrict__ x, int32_t N, int D, const int32_t *__restrict__ seeds,
                   int64_t S, int64_t e_lo, int64_t e_hi, int k, uint64_t rng_seed,
                   int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                   float *__restrict__ out_x) {
  extern __shared__ int64_t s_pick[];  // [warps][k] chosen entry index, -1 = padding
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int64_t *pick = s_pick + warp * k;
  for (int64_t s = int64_t(blockIdx.x) * wpb + warp; s < S; s += int64_t(gridDim.x) * wpb) {
    const int32_t v = seeds[s];
    int64_t lo = 0, cnt = 0;
    if (v >= 0 && v < N) {
      const int64_t r0 = rowptr[v], r1 = rowptr[v + 1];
      if (BY_TIME) {
        lo = lower_bound_time(entries, r0, r1, e_lo);
        cnt = (e_hi == INT64_MAX ? r1 : lower_bound_time(entries, lo, r1, e_hi + 1)) - lo;
      } else {
        lo = lower_bound_eid(entries, r0, r1, e_lo);
        cnt = lower_bound_eid(entries, lo, r1, e_hi) - lo;
      }
    }
    if (cnt <= k) {
      for (int c = lane; c < k; c += 32) pick[c] = c < cnt ? lo + c : -1;
    } else if (lane == 0) {
      // Floyd: for j = cnt-k .. cnt-1 pick t ~ U[0, j]; if t was already chosen take j instead
      uint64_t state = splitmix64(rng_seed ^ (uint64_t(uint32_t(v)) * 0xD1B54A32D192ED03ull));
      for (int i = 0; i < k; ++i) {
        const int64_t j = cnt - k + i;
        state = splitmix64(state);
        int64_t t = int64_t(state % uint64_t(j + 1));
        for (int u = 0; u < i; ++u)
          if (pick[u] == lo + t) {
            t = j;
            break;
          }
        pick[i] = lo + t;
      }
    }
    __syncwarp();
    for (int c = lane; c < k; c += 32) {
      const int64_t p = pick[c];
      int32_t id = TGM_PADDED_NODE_ID;
      int64_t tt = 0;
      if (p >= 0) {
        const Entry en = entries[p];
        id = en.nbr;
        tt = en.t;
        pick[c] = en.eid;  // feature row
      }
      out_nid[s * k + c] = id;
      out_t[s * k + c] = tt;
    }
    __syncwarp();
    if (D > 0) {
      float *o = out_x + s * int64_t(k) * D;
      const int total = k * D;
      for (int i = lane; i < total; i += 32) {
        const int c = i / D, d = i - c * D;
        const int64_t row = pick[c];
        o[i] = row >= 0 ? __ldg(x + row * D + d) : 0.f;
      }
    }
    __syncwarp();
  }
}

// ---- TMA variant of the fast hop-0 kernel -------------------------------------------------------
// Same per-seed logic, but the feature rows never pass through registers: lane 0 of each warp
// drives a ring of shared-memory stages with 1-D bulk async copies (cp.async.bulk, the TMA unit):
//   global xrows --(bulk load, mbarrier complete_tx)--> smem stage --(bulk store)--> global out_x
// Loads run kTmaLag seeds ahead of the stores, so every warp keeps several 1-KB row blocks in
// flight without spending registers or issue slots on them; the left padding of a row is a bulk
// store from a zeroed smem block.  Used when a seed's feature block (k*D*4 bytes) fits a stage.
#ifndef TGM_TMA_STAGES
#define TGM_TMA_STAGES 4
#define TGM_TMA_LAG 2
#endif
constexpr int kTmaStages = TGM_TMA_STAGES, kTmaLag = TGM_TMA_LAG, kTmaMaxStageBytes = 4096;

struct TmaMeta {  // what the retiring step needs to know about an in-flight seed
  float *dst;        // first float of the seed's out_x row block
  uint32_t nbytes;   // valid feature bytes (bulk-loaded); 0 = nothing was loaded
  uint32_t padbytes; // zero bytes in front of them; bit 31 carries the mbarrier phase parity
};

// EDGE_SEEDS: hop-0 seeds are the endpoints of stream edges [l_lo, l_hi) (anchor table, no seed
// arrays); otherwise arbitrary seeds (seeds/tq/cut arrays, private binary search per lane).
struct TmaSeedArgs {
  // EDGE_SEEDS
  const uint2 *anchors;
  const int64_t *t;
  int64_t Ew, l_lo, l_hi;
  uint32_t bs;
  // general seeds
  const int64_t *rowptr;
  const int32_t *seeds;
  const int64_t *tq, *cut;
  int64_t cut_group, S;
  int32_t N;
};

template <bool EDGE_SEEDS>
__global__ void __launch_bounds__(kFastThreads)
csr_sample_tma_kernel(const Entry *__restrict__ entries, const float *__restrict__ xrows,
                      const TmaSeedArgs a, int D, int B, int k, int32_t *__restrict__ out_nid,
                      int64_t *__restrict__ out_t, float *__restrict__ out_x, int stage_bytes,
                      unsigned long long *__restrict__ ticket) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int W = kFastThreads >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // layout: zero block | W * kTmaStages stages | W * kTmaStages mbarriers | W * kTmaStages metas
  unsigned char *zero = smem_raw;
  unsigned char *stages = zero + stage_bytes + size_t(warp) * kTmaStages * stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + stage_bytes +
                                                size_t(W) * kTmaStages * stage_bytes) +
                   warp * kTmaStages;
  TmaMeta *meta = reinterpret_cast<TmaMeta *>(smem_raw + stage_bytes +
                                              size_t(W) * kTmaStages * (stage_bytes + 8)) +
                  warp * kTmaStages;
  for (int i = threadIdx.x * 4; i < stage_bytes; i += kFastThreads * 4)
    *reinterpret_cast<uint32_t *>(zero + i) = 0u;
  if (lane == 0)
    for (int s = 0; s < kTmaStages; ++s) mbar_init(smem_u32(bars + s), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zero_s = smem_u32(zero), stage_s = smem_u32(stages), bar_s = smem_u32(bars);

  auto retire = [&](uint32_t q) {  // lane 0 only: seed number q of this warp has landed -> store it
    const int st = int(q % kTmaStages);
    const TmaMeta m = meta[st];
    const uint32_t padbytes = m.padbytes & 0x7fffffffu;
    if (m.nbytes) {
      mbar_wait(bar_s + st * 8, m.padbytes >> 31);
      bulk_s2g(reinterpret_cast<unsigned char *>(m.dst) + padbytes, stage_s + st * stage_bytes,
               m.nbytes);
    }
    if (padbytes) bulk_s2g(m.dst, zero_s, padbytes);
    bulk_commit();
  };

  const int64_t S = EDGE_SEEDS ? 2 * (a.l_hi - a.l_lo) : a.S;
  const int64_t nchunks = (S + 31) >> 5;
  const int64_t wstride = int64_t(gridDim.x) * W;
  uint32_t g = 0;       // seeds this warp has started
  uint32_t phases = 0;  // lane 0: bit st = parity of the next phase of stage st's mbarrier (a
                        // stage's barrier only advances when a load is actually issued on it)
  // chunks of 32 seeds are handed out dynamically (first one by warp id, the rest from a ticket
  // counter): windows differ in length, and a static split leaves a ~1/iterations tail
  for (int64_t ch = int64_t(blockIdx.x) * W + warp; ch < nchunks;) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    SeedWin mine{0, 0, 0};
    if (s < S) {
      if (EDGE_SEEDS) {
        // row s -> (edge, endpoint): batch jb owns rows [2*bs*jb, ...), src seeds then dst seeds
        const int64_t bs = a.bs;
        const int64_t jb = s / (2 * bs);
        const int64_t bstart = a.l_lo + jb * bs;
        const int64_t nb = a.l_hi - bstart < bs ? a.l_hi - bstart : bs;
        const int64_t rr = s - jb * 2 * bs;
        const bool side = rr >= nb;
        const int64_t l = bstart + (side ? rr - nb : rr);
        const uint2 an = __ldg(a.anchors + (side ? a.Ew + l : l));
        mine.nwin = an.y < uint32_t(B) ? int(an.y) : B;
        mine.wstart = int64_t(an.x) - mine.nwin;
        mine.q = __ldg(a.t + l);
      } else {
        const int32_t v = __ldg(a.seeds + s);
        mine.q = __ldg(a.tq + s);
        if (v >= 0 && v < a.N) {  // a padded seed (-1) always yields an all-padding row
          const int64_t lo = __ldg(a.rowptr + v), hi = __ldg(a.rowptr + v + 1);
          const int64_t pos = lower_bound_eid(entries, lo, hi, __ldg(a.cut + s / a.cut_group));
          mine.wstart = pos - B > lo ? pos - B : lo;
          mine.nwin = int(pos - mine.wstart);
        }
      }
    }
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    int64_t wstart = shfl_i64(mine.wstart, 0);
    int nwin = __shfl_sync(0xffffffffu, mine.nwin, 0);
    Entry cur = load_window_entry(entries, wstart, nwin, lane);
    for (int i = 0; i < nseeds; ++i, ++g) {
      const int64_t q = shfl_i64(mine.q, i);
      const int nxt = i + 1 < nseeds ? i + 1 : i;
      const int64_t wstart_n = shfl_i64(mine.wstart, nxt);
      const int nwin_n = __shfl_sync(0xffffffffu, mine.nwin, nxt);
      const Entry ahead = load_window_entry(entries, wstart_n, nwin_n, lane);
      const unsigned m = __ballot_sync(0xffffffffu, lane < nwin && cur.t < q);
      const int last = m ? 31 - __clz(m) : -1;
      const int nvalid = last + 1 < k ? last + 1 : k;
      const int first = last + 1 - nvalid, pad = k - nvalid;
      const int srcl = (first + lane - pad) & 31;
      const int32_t nbr = __shfl_sync(0xffffffffu, cur.nbr, srcl);
      const int64_t tt = shfl_i64(cur.t, srcl);
      const int64_t sg = s_base + i;
      if (lane < k) {
        const bool v = lane >= pad;
        out_nid[sg * k + lane] = v ? nbr : TGM_PADDED_NODE_ID;
        out_t[sg * k + lane] = v ? tt : 0;
      }
      if (lane == 0) {
        const int st = int(g % kTmaStages);
        bulk_wait_read<kTmaStages - kTmaLag - 1>();  // the store that last read this stage is done
        TmaMeta mt;
        mt.dst = out_x + sg * int64_t(k) * D;
        mt.nbytes = uint32_t(nvalid) * uint32_t(D) * 4u;
        mt.padbytes = uint32_t(pad) * uint32_t(D) * 4u;
        if (mt.nbytes) {
          mt.padbytes |= ((phases >> st) & 1u) << 31;
          phases ^= 1u << st;
        }
        meta[st] = mt;
        if (mt.nbytes) {
          mbar_expect_tx(bar_s + st * 8, mt.nbytes);
          bulk_g2s(stage_s + st * stage_bytes, xrows + (wstart + first) * int64_t(D), mt.nbytes,
                   bar_s + st * 8);
        }
        if (g >= uint32_t(kTmaLag)) retire(g - kTmaLag);
      }
      cur = ahead;
      wstart = wstart_n;
      nwin = nwin_n;
    }
    ch = next_chunk(ticket, wstride, lane);
  }
  if (lane == 0) {  // drain
    for (uint32_t q = g >= uint32_t(kTmaLag) ? g - kTmaLag : 0; q < g; ++q) retire(q);
    bulk_wait_read<0>();
  }
}

// wall-clock phase marks of a build (each mark synchronises the stream; only when tracing)
struct Trace {
  cudaStream_t st;
  std::chrono::steady_clock::time_point last;
  explicit Trace(cudaStream_t s) : st(s), last(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!g_trace) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[tgm_csr_build] %-16s %8.2f ms\n", what,
            std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
};

int bits_for(uint32_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b) != 0) ++b;
  return b;
}

}  // namespace

extern "C" int tgm_csr_build(tgm_csr **out, const tgm_store *store, int64_t e_start,
                             int64_t batch_size, int directed, int colocate_x, tgm_stream stream) {
  TGM_REQUIRE(out != nullptr, "tgm_csr_build: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(store != nullptr, "tgm_csr_build: store is NULL");
  if (store->device < 0) return fail(TGM_ERR_NO_DEVICE, "tgm_csr_build: metadata-only store");
  TGM_REQUIRE(batch_size > 0, "tgm_csr_build: batch_size must be > 0");
  TGM_REQUIRE(e_start >= 0 && e_start <= store->E, "tgm_csr_build: e_start outside [0, E]");
  TGM_REQUIRE(store->num_nodes > 0, "tgm_csr_build: store has no nodes");
  DeviceGuard g(store->device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_csr_build: cannot select device");
  cudaStream_t st = as_stream(stream);

  tgm_csr *c = new (std::nothrow) tgm_csr();
  if (!c) return fail(TGM_ERR_OOM, "tgm_csr_build: host allocation failed");
  c->device = store->device;
  c->store = store;
  c->e_start = e_start;
  c->Ew = store->E - e_start;
  c->bs = batch_size;
  c->directed = directed ? 1 : 0;
  c->N = store->num_nodes;
  c->D = store->D;
  c->n = directed ? c->Ew : 2 * c->Ew;
  c->colocate = (colocate_x && c->D > 0) ? 1 : 0;
  const int64_t n = c->n, Ew = c->Ew;

  uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
  unsigned char *work = nullptr;
  auto cleanup_tmp = [&]() { cudaFree(work); };
  auto bail = [&](int code) {
    cleanup_tmp();
    delete c;
    return code;
  };
#define CSR_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) return bail(cuda_fail(_e, #expr, __FILE__, __LINE__));   \
  } while (0)

  Trace tr(st);
  CSR_CUDA(cudaMalloc(&c->rowptr, size_t(c->N + 1) * 8));
  CSR_CUDA(cudaMalloc(&c->ticket, tgm_csr::kTickets * sizeof(unsigned long long)));
  if (n == 0) {
    CSR_CUDA(cudaMemsetAsync(c->rowptr, 0, size_t(c->N + 1) * 8, st));
    CSR_CUDA(cudaStreamSynchronize(st));
    *out = c;
    return TGM_OK;
  }
  const size_t nn = size_t(n);
  CSR_CUDA(cudaMalloc(&c->entries, nn * sizeof(Entry)));
  CSR_CUDA(cudaMalloc(&c->anchors, size_t(2 * Ew) * sizeof(uint2)));
  if (c->colocate) CSR_CUDA(cudaMalloc(&c->xrows, nn * size_t(c->D) * 4));
  // one workspace for every temporary: 4 key/value arrays + the CUB scratch
  const int end_bit = bits_for(uint32_t(c->N - 1));
  size_t sort_bytes = 0, scan_bytes = 0;
  CSR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_a, keys_b, vals_a, vals_b, n, 0,
                                           end_bit, st));
  CSR_CUDA(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, keys_a, vals_a, MaxU32(), n, st));
  const size_t arr = (nn * 4 + 255) & ~size_t(255);
  const size_t tmp_bytes = std::max(sort_bytes, scan_bytes);
  CSR_CUDA(cudaMalloc(&work, 4 * arr + tmp_bytes + 256));
  keys_a = reinterpret_cast<uint32_t *>(work);
  keys_b = reinterpret_cast<uint32_t *>(work + arr);
  vals_a = reinterpret_cast<uint32_t *>(work + 2 * arr);
  vals_b = reinterpret_cast<uint32_t *>(work + 3 * arr);
  void *cub_tmp = work + 4 * arr;
  tr.mark("alloc");

  const int32_t *src = store->src + e_start, *dst = store->dst + e_start;
  const int64_t *t = store->t + e_start;
  const int threads = 256;
  csr_keys_kernel<<<grid_for(Ew, threads, 8), threads, 0, st>>>(src, dst, t, Ew, batch_size,
                                                                c->directed, keys_a, vals_a);
  CSR_CUDA(cudaGetLastError());
  tr.mark("keys");

  // stable LSD radix sort by node id; the input order already is the chronological order
  size_t tb = tmp_bytes;
  CSR_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys_a, keys_b, vals_a, vals_b, n, 0,
                                           end_bit, st));
  tr.mark("sort");

  csr_rowptr_kernel<<<grid_for(int64_t(c->N) + 1, threads, 8), threads, 0, st>>>(keys_b, n, c->N,
                                                                                 c->rowptr);
  CSR_CUDA(cudaGetLastError());
  const bool scan_anchors = !c->directed && batch_size < (int64_t(1) << 32);
  if (scan_anchors) {
    csr_heads_kernel<<<grid_for(n, threads, 8), threads, 0, st>>>(keys_b, vals_b, n,
                                                                  uint32_t(batch_size), keys_a);
    CSR_CUDA(cudaGetLastError());
    tb = tmp_bytes;
    CSR_CUDA(cub::DeviceScan::InclusiveScan(cub_tmp, tb, keys_a, vals_a, MaxU32(), n, st));
  }
  tr.mark("rowptr+runs");
  csr_entries_kernel<<<grid_for(n, threads, 8), threads, 0, st>>>(
      keys_b, vals_b, vals_a, n, src, dst, t, c->rowptr, e_start, Ew, c->entries,
      scan_anchors ? c->anchors : nullptr);
  CSR_CUDA(cudaGetLastError());
  if (!scan_anchors) {
    csr_anchor_kernel<<<grid_for(2 * Ew, threads, 8), threads, 0, st>>>(
        c->entries, c->rowptr, src, dst, Ew, batch_size, e_start, c->N, c->anchors);
    CSR_CUDA(cudaGetLastError());
  }
  tr.mark("entries+anchors");
  if (c->colocate) {
    const bool vec4 = (c->D % 4 == 0) && aligned16(store->x) && aligned16(c->xrows);
    if (vec4)
      csr_gather_x_kernel<true><<<grid_for(n * (c->D / 4), threads, 16), threads, 0, st>>>(
          c->entries, n, store->x, c->D, c->xrows);
    else
      csr_gather_x_kernel<false><<<grid_for(n * c->D, threads, 16), threads, 0, st>>>(
          c->entries, n, store->x, c->D, c->xrows);
    CSR_CUDA(cudaGetLastError());
  }
  CSR_CUDA(cudaStreamSynchronize(st));
  tr.mark("gather_x");
#undef CSR_CUDA
  cleanup_tmp();
  tr.mark("free");
  *out = c;
  return TGM_OK;
}

extern "C" void tgm_csr_destroy(tgm_csr *c) { delete c; }

extern "C" int tgm_csr_info(const tgm_csr *c, int64_t *num_entries, int64_t *e_start,
                            int64_t *batch_size, int *directed, int *colocate_x) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_info: csr is NULL");
  if (num_entries) *num_entries = c->n;
  if (e_start) *e_start = c->e_start;
  if (batch_size) *batch_size = c->bs;
  if (directed) *directed = c->directed;
  if (colocate_x) *colocate_x = c->colocate;
  return TGM_OK;
}

namespace {
struct SampleCfg {
  bool vec4, coloc, fast;
  int fast_grid;
  const float *xsrc;
  int grid;
  size_t smem;
};
int sample_cfg(const tgm_csr *c, const char *who, int64_t S, int32_t B, int32_t k, float *out_x,
               SampleCfg *cfg) {
  if (!(B >= 1)) return fail(TGM_ERR_INVALID, std::string(who) + ": B must be >= 1");
  if (!(k >= 1 && k <= B)) return fail(TGM_ERR_INVALID, std::string(who) + ": k must be in [1, B]");
  if (c->D > 0 && out_x == nullptr)
    return fail(TGM_ERR_INVALID, std::string(who) + ": out_x is NULL but D > 0");
  cfg->coloc = c->colocate != 0;
  cfg->xsrc = cfg->coloc ? c->xrows : c->store->x;
  cfg->vec4 = c->D > 0 && (c->D % 4 == 0) && aligned16(cfg->xsrc) && aligned16(out_x) &&
              ((int64_t(k) * c->D * 4) % 16 == 0);
  const int wpb = kSampleThreads / 32;
  cfg->smem = size_t(wpb) * size_t(k) * sizeof(uint32_t);
  if (cfg->smem > 48 * 1024) return fail(TGM_ERR_INVALID, std::string(who) + ": k too large");
  cfg->grid = grid_for(S, wpb, 8);
  // warp-per-32-seeds kernels: ring width fits a warp and feature rows (if any) are colocated
  cfg->fast = B <= 32 && (c->D == 0 || (cfg->vec4 && cfg->coloc));
  cfg->fast_grid = grid_for((S + 31) / 32, kFastThreads / 32, TGM_FAST_MIN_BLOCKS);
  return TGM_OK;
}
}  // namespace

namespace {
// a zeroed chunk counter for one launch (launches of a handle rotate through kTickets counters)
int new_ticket(const tgm_csr *c, cudaStream_t st, unsigned long long **out) {
  *out = c->ticket + (c->launches++ % tgm_csr::kTickets);
  TGM_CUDA(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), st));
  return TGM_OK;
}
bool tma_applies(const tgm_csr *c, int k) {
  return c->D > 0 && g_csr_feature_copy == 1 && k * c->D * 4 <= kTmaMaxStageBytes;
}
template <bool EDGE_SEEDS>
int launch_tma(const tgm_csr *c, const TmaSeedArgs &a, int64_t S, int B, int k, const float *xrows,
               int32_t *out_nid, int64_t *out_t, float *out_x, cudaStream_t st) {
  const int stage_bytes = k * c->D * 4;
  const size_t smem = size_t(stage_bytes) +
                      size_t(kFastThreads / 32) * kTmaStages * (size_t(stage_bytes) + 8 + 16);
  if (smem > 48 * 1024)
    TGM_CUDA(cudaFuncSetAttribute(csr_sample_tma_kernel<EDGE_SEEDS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = int(std::max<size_t>(1, std::min<size_t>(TGM_FAST_MIN_BLOCKS, (220 * 1024) / (smem + 1024))));
  if (g_csr_tma_ctas_per_sm > 0 && g_csr_tma_ctas_per_sm < per_sm) per_sm = g_csr_tma_ctas_per_sm;
  const int grid = grid_for((S + 31) / 32, kFastThreads / 32, per_sm);
  unsigned long long *ticket = nullptr;
  int trc = new_ticket(c, st, &ticket);
  if (trc != TGM_OK) return trc;
  csr_sample_tma_kernel<EDGE_SEEDS><<<grid, kFastThreads, smem, st>>>(
      c->entries, xrows, a, c->D, B, k, out_nid, out_t, out_x, stage_bytes, ticket);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
}  // namespace

#define DISPATCH_SAMPLE(KERNEL, ...)                                                           \
  do {                                                                                         \
    if (cfg.vec4 && cfg.coloc)                                                                 \
      KERNEL<true, true><<<cfg.grid, kSampleThreads, cfg.smem, st>>>(__VA_ARGS__);             \
    else if (cfg.vec4)                                                                         \
      KERNEL<true, false><<<cfg.grid, kSampleThreads, cfg.smem, st>>>(__VA_ARGS__);            \
    else if (cfg.coloc)                                                                        \
      KERNEL<false, true><<<cfg.grid, kSampleThreads, cfg.smem, st>>>(__VA_ARGS__);            \
    else                                                                                       \
      KERNEL<false, false><<<cfg.grid, kSampleThreads, cfg.smem, st>>>(__VA_ARGS__);           \
  } while (0)

extern "C" int tgm_csr_sample(const tgm_csr *c, const int32_t *seeds, const int64_t *tq,
                              const int64_t *cut, int64_t cut_group, int64_t S, int32_t B,
                              int32_t k, int32_t *out_nid, int64_t *out_t, float *out_x,
                              tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample: csr is NULL");
  TGM_REQUIRE(S >= 0, "tgm_csr_sample: S must be >= 0");
  TGM_REQUIRE(cut_group >= 1, "tgm_csr_sample: cut_group must be >= 1");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && tq && cut && out_nid && out_t, "tgm_csr_sample: NULL array argument");
  SampleCfg cfg;
  int rc = sample_cfg(c, "tgm_csr_sample", S, B, k, out_x, &cfg);
  if (rc != TGM_OK) return rc;
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  if (cfg.fast && tma_applies(c, k)) {
    TmaSeedArgs a{};
    a.rowptr = c->rowptr, a.seeds = seeds, a.tq = tq, a.cut = cut, a.cut_group = cut_group, a.S = S;
    a.N = c->N;
    rc = launch_tma<false>(c, a, S, B, k, cfg.xsrc, out_nid, out_t, out_x, st);
    if (rc != TGM_OK) return rc;
  } else if (cfg.fast) {
    unsigned long long *ticket = nullptr;
    rc = new_ticket(c, st, &ticket);
    if (rc != TGM_OK) return rc;
    csr_sample_fast_kernel<<<cfg.fast_grid, kFastThreads, 0, st>>>(
        c->entries, c->rowptr, reinterpret_cast<const float4 *>(cfg.xsrc), c->N, c->D / 4, seeds,
        tq, cut, cut_group, S, B, k, out_nid, out_t, reinterpret_cast<float4 *>(out_x), ticket,
        nullptr);
  } else {
    DISPATCH_SAMPLE(csr_sample_kernel, c->entries, c->rowptr, cfg.xsrc, c->N, c->D, seeds, tq,
                    cut, cut_group, S, B, k, out_nid, out_t, out_x);
  }
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_csr_sample_edges(const tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                    int32_t k, int32_t *out_nid, int64_t *out_t, float *out_x,
                                    tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_edges: csr is NULL");
  const int64_t l_lo = e_lo - c->e_start, l_hi = e_hi - c->e_start;
  TGM_REQUIRE(l_lo >= 0 && l_lo <= l_hi && l_hi <= c->Ew,
              "tgm_csr_sample_edges: [e_lo, e_hi) outside the indexed stream");
  TGM_REQUIRE(l_lo % c->bs == 0, "tgm_csr_sample_edges: e_lo must sit on a batch boundary");
  const int64_t S = 2 * (l_hi - l_lo);
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(out_nid && out_t, "tgm_csr_sample_edges: NULL array argument");
  SampleCfg cfg;
  int rc = sample_cfg(c, "tgm_csr_sample_edges", S, B, k, out_x, &cfg);
  if (rc != TGM_OK) return rc;
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  if (cfg.fast && c->bs < (int64_t(1) << 31) && tma_applies(c, k)) {
    TmaSeedArgs a{};
    a.anchors = c->anchors, a.t = c->store->t + c->e_start, a.Ew = c->Ew, a.l_lo = l_lo, a.l_hi = l_hi;
    a.bs = uint32_t(c->bs);
    rc = launch_tma<true>(c, a, S, B, k, cfg.xsrc, out_nid, out_t, out_x, st);
    if (rc != TGM_OK) return rc;
  } else if (cfg.fast && c->bs < (int64_t(1) << 31)) {
    unsigned long long *ticket = nullptr;
    rc = new_ticket(c, st, &ticket);
    if (rc != TGM_OK) return rc;
    csr_sample_edges_fast_kernel<<<cfg.fast_grid, kFastThreads, 0, st>>>(
        c->entries, c->anchors, reinterpret_cast<const float4 *>(cfg.xsrc),
        c->store->t + c->e_start, c->D / 4, c->Ew, uint32_t(c->bs), l_lo, l_hi, B, k, out_nid,
        out_t, reinterpret_cast<float4 *>(out_x), ticket);
  } else {
    DISPATCH_SAMPLE(csr_sample_edges_kernel, c->entries, c->anchors, cfg.xsrc,
                    c->store->t + c->e_start, c->D, c->Ew, c->bs, l_lo, l_hi, B, k, out_nid, out_t,
                    out_x);
  }
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}


namespace {
// common checks + argument block of the hop-0 forms without a feature block
int edge_seed_args(const tgm_csr *c, const char *who, int64_t e_lo, int64_t e_hi, int32_t B,
                   int32_t k, EdgeSeedArgs *a) {
  if (c == nullptr) return fail(TGM_ERR_INVALID, std::string(who) + ": csr is NULL");
  const int64_t l_lo = e_lo - c->e_start, l_hi = e_hi - c->e_start;
  if (!(l_lo >= 0 && l_lo <= l_hi && l_hi <= c->Ew))
    return fail(TGM_ERR_INVALID, std::string(who) + ": [e_lo, e_hi) outside the indexed stream");
  if (l_lo % c->bs != 0)
    return fail(TGM_ERR_INVALID, std::string(who) + ": e_lo must sit on a batch boundary");
  if (!(B >= 1 && B <= 32))
    return fail(TGM_ERR_INVALID, std::string(who) + ": B must be in [1, 32]");
  if (!(k >= 1 && k <= B)) return fail(TGM_ERR_INVALID, std::string(who) + ": k must be in [1, B]");
  if (c->bs >= (int64_t(1) << 31))
    return fail(TGM_ERR_INVALID, std::string(who) + ": batch_size must be < 2^31");
  a->anchors = c->anchors;
  a->rowptr = c->rowptr;
  a->src = c->store->src + c->e_start;
  a->dst = c->store->dst + c->e_start;
  a->t = c->store->t + c->e_start;
  a->Ew = c->Ew, a->l_lo = l_lo, a->l_hi = l_hi, a->e_start = c->e_start;
  a->bs = uint32_t(c->bs);
  a->N = c->N;
  return TGM_OK;
}

int stage_reserve(tgm_csr::Stage &sg, int which, size_t bytes, cudaStream_t st) {
  if (bytes <= sg.cap[which]) return TGM_OK;
  TGM_CUDA(cudaStreamSynchronize(st));  // growing is the only synchronising path
  cudaFree(sg.p[which]);
  sg.p[which] = nullptr, sg.cap[which] = 0;
  TGM_CUDA(cudaMalloc(&sg.p[which], bytes));
  sg.cap[which] = bytes;
  return TGM_OK;
}

// H2D of the caller's slab of stream edges into the store (any pointer may be NULL = resident)
int upload_slab(const tgm_csr *c, int64_t e_lo, int64_t nE, const int32_t *h_src,
                const int32_t *h_dst, const int64_t *h_t, const float *h_x, cudaStream_t st) {
  const tgm_store *s = c->store;
  const size_t D = size_t(c->D);
  if (h_src) TGM_CUDA(cudaMemcpyAsync(const_cast<int32_t *>(s->src) + e_lo, h_src, size_t(nE) * 4,
                                      cudaMemcpyHostToDevice, st));
  if (h_dst) TGM_CUDA(cudaMemcpyAsync(const_cast<int32_t *>(s->dst) + e_lo, h_dst, size_t(nE) * 4,
                                      cudaMemcpyHostToDevice, st));
  if (h_t) TGM_CUDA(cudaMemcpyAsync(const_cast<int64_t *>(s->t) + e_lo, h_t, size_t(nE) * 8,
                                    cudaMemcpyHostToDevice, st));
  if (h_x && D) TGM_CUDA(cudaMemcpyAsync(const_cast<float *>(s->x) + size_t(e_lo) * D, h_x,
                                         size_t(nE) * D * 4, cudaMemcpyHostToDevice, st));
  return TGM_OK;
}
}  // namespace

extern "C" int tgm_csr_sample_ids(const tgm_csr *c, const int32_t *seeds, const int64_t *tq,
                                  const int64_t *cut, int64_t cut_group, int64_t S, int32_t B,
                                  int32_t k, int32_t *out_nid, int64_t *out_t, int32_t *out_eid,
                                  tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_ids: csr is NULL");
  TGM_REQUIRE(S >= 0, "tgm_csr_sample_ids: S must be >= 0");
  TGM_REQUIRE(cut_group >= 1, "tgm_csr_sample_ids: cut_group must be >= 1");
  TGM_REQUIRE(B >= 1 && B <= 32, "tgm_csr_sample_ids: B must be in [1, 32]");
  TGM_REQUIRE(k >= 1 && k <= B, "tgm_csr_sample_ids: k must be in [1, B]");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && tq && cut && (out_nid || out_t || out_eid),
              "tgm_csr_sample_ids: NULL array argument");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  unsigned long long *ticket = nullptr;
  int rc = new_ticket(c, st, &ticket);
  if (rc != TGM_OK) return rc;
  const int grid = grid_for((S + 31) / 32, kFastThreads / 32, TGM_FAST_MIN_BLOCKS);
  csr_sample_fast_kernel<<<grid, kFastThreads, 0, st>>>(c->entries, c->rowptr, nullptr, c->N, 0,
                                                        seeds, tq, cut, cut_group, S, B, k, out_nid,
                                                        out_t, nullptr, ticket, out_eid);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_csr_sample_edges_ids(const tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                        int32_t k, int search, int32_t *out_nid, int64_t *out_t,
                                        int32_t *out_eid, tgm_stream stream) {
  EdgeSeedArgs a{};
  int rc = edge_seed_args(c, "tgm_csr_sample_edges_ids", e_lo, e_hi, B, k, &a);
  if (rc != TGM_OK) return rc;
  const int64_t S = 2 * (a.l_hi - a.l_lo);
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(out_nid || out_t || out_eid, "tgm_csr_sample_edges_ids: every output is NULL");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  unsigned long long *ticket = nullptr;
  rc = new_ticket(c, st, &ticket);
  if (rc != TGM_OK) return rc;
  const int grid = grid_for((S + 31) / 32, kFastThreads / 32, kIdsMinBlocks);
  if (search)
    csr_sample_edges_ids_kernel<true><<<grid, kFastThreads, 0, st>>>(c->entries, a, B, k, out_nid,
                                                                     out_t, out_eid, ticket);
  else
    csr_sample_edges_ids_kernel<false><<<grid, kFastThreads, 0, st>>>(c->entries, a, B, k, out_nid,
                                                                      out_t, out_eid, ticket);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_csr_sample_edges_mean(const tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                         int32_t k, int search, int32_t *out_nid, int64_t *out_t,
                                         int32_t *out_eid, float *out_mean, tgm_stream stream) {
  EdgeSeedArgs a{};
  int rc = edge_seed_args(c, "tgm_csr_sample_edges_mean", e_lo, e_hi, B, k, &a);
  if (rc != TGM_OK) return rc;
  TGM_REQUIRE(c->D > 0 && c->D % 4 == 0 && c->colocate && aligned16(c->xrows),
              "tgm_csr_sample_edges_mean: needs colocated feature rows with D % 4 == 0 (otherwise "
              "tgm_csr_sample_edges + tgm_masked_mean)");
  const int64_t S = 2 * (a.l_hi - a.l_lo);
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(out_mean != nullptr && aligned16(out_mean),
              "tgm_csr_sample_edges_mean: out_mean must be a 16-byte aligned array");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  unsigned long long *ticket = nullptr;
  rc = new_ticket(c, st, &ticket);
  if (rc != TGM_OK) return rc;
  const int D4 = c->D / 4;
  int G = 1;
  while (G < D4 && G < 32) G <<= 1;
  const int grid = grid_for((S + 31) / 32, kFastThreads / 32, kIdsMinBlocks);
  const float4 *x4 = reinterpret_cast<const float4 *>(c->xrows);
  float4 *o4 = reinterpret_cast<float4 *>(out_mean);
  if (search)
    csr_sample_edges_mean_kernel<true><<<grid, kFastThreads, 0, st>>>(
        c->entries, a, x4, D4, G, B, k, out_nid, out_t, out_eid, o4, ticket);
  else
    csr_sample_edges_mean_kernel<false><<<grid, kFastThreads, 0, st>>>(
        c->entries, a, x4, D4, G, B, k, out_nid, out_t, out_eid, o4, ticket);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

// Host-buffer form: H2D of the slab, the same kernel, D2H of the outputs, all on `stream`.
extern "C" int tgm_csr_sample_edges_host(tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                         int32_t k, const int32_t *h_src, const int32_t *h_dst,
                                         const int64_t *h_t, const float *h_x, int32_t *h_out_nid,
                                         int64_t *h_out_t, float *h_out_x, int slot,
                                         tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_edges_host: csr is NULL");
  TGM_REQUIRE(slot >= 0 && slot < TGM_HOST_SLOTS, "tgm_csr_sample_edges_host: bad slot");
  const int64_t l_lo = e_lo - c->e_start, l_hi = e_hi - c->e_start;
  TGM_REQUIRE(l_lo >= 0 && l_lo <= l_hi && l_hi <= c->Ew,
              "tgm_csr_sample_edges_host: [e_lo, e_hi) outside the indexed stream");
  const int64_t nE = e_hi - e_lo, cells = 2 * nE * k;
  if (nE == 0) return TGM_OK;
  TGM_REQUIRE(h_out_nid && h_out_t, "tgm_csr_sample_edges_host: NULL output argument");
  TGM_REQUIRE(c->D == 0 || h_out_x, "tgm_csr_sample_edges_host: h_out_x is NULL but D > 0");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  const size_t D = size_t(c->D);
  // the slab of the window as the caller holds it on the host (the reference keeps its arrays on
  // the CPU and copies each batch's properties to the device, tgm/core/graph.py:232-263)
  int rc = upload_slab(c, e_lo, nE, h_src, h_dst, h_t, h_x, st);
  if (rc != TGM_OK) return rc;
  tgm_csr::Stage &sg = c->stage[slot];
  if ((rc = stage_reserve(sg, 0, size_t(cells) * 4, st)) != TGM_OK) return rc;
  if ((rc = stage_reserve(sg, 1, size_t(cells) * 8, st)) != TGM_OK) return rc;
  if (D && (rc = stage_reserve(sg, 2, size_t(cells) * D * 4, st)) != TGM_OK) return rc;
  int32_t *d_nid = static_cast<int32_t *>(sg.p[0]);
  int64_t *d_t = static_cast<int64_t *>(sg.p[1]);
  float *d_x = static_cast<float *>(sg.p[2]);
  rc = tgm_csr_sample_edges(c, e_lo, e_hi, B, k, d_nid, d_t, d_x, stream);
  if (rc != TGM_OK) return rc;
  TGM_CUDA(cudaMemcpyAsync(h_out_nid, d_nid, size_t(cells) * 4, cudaMemcpyDeviceToHost, st));
  TGM_CUDA(cudaMemcpyAsync(h_out_t, d_t, size_t(cells) * 8, cudaMemcpyDeviceToHost, st));
  if (D) TGM_CUDA(cudaMemcpyAsync(h_out_x, d_x, size_t(cells) * D * 4, cudaMemcpyDeviceToHost, st));
  return TGM_OK;
}

// Host-buffer forms that move only what the host lacks.  Both consume the uploaded slab: the
// kernel reads its seeds (src/dst), query times and history cuts from it (search form).
extern "C" int tgm_csr_sample_edges_host_ids(tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                             int32_t k, const int32_t *h_src, const int32_t *h_dst,
                                             const int64_t *h_t, int32_t *h_out_nid,
                                             int64_t *h_out_t, int32_t *h_out_eid, int slot,
                                             tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_edges_host_ids: csr is NULL");
  TGM_REQUIRE(slot >= 0 && slot < TGM_HOST_SLOTS, "tgm_csr_sample_edges_host_ids: bad slot");
  const int64_t nE = e_hi - e_lo, cells = 2 * nE * k;
  TGM_REQUIRE(nE >= 0 && e_lo >= 0 && e_hi <= c->store->E,
              "tgm_csr_sample_edges_host_ids: [e_lo, e_hi) outside the store");
  if (nE == 0) return TGM_OK;
  TGM_REQUIRE(h_out_nid && h_out_t && h_out_eid,
              "tgm_csr_sample_edges_host_ids: NULL output argument");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  int rc = upload_slab(c, e_lo, nE, h_src, h_dst, h_t, nullptr, st);
  if (rc != TGM_OK) return rc;
  tgm_csr::Stage &sg = c->stage[slot];
  if ((rc = stage_reserve(sg, 0, size_t(cells) * 4, st)) != TGM_OK) return rc;
  if ((rc = stage_reserve(sg, 1, size_t(cells) * 8, st)) != TGM_OK) return rc;
  if ((rc = stage_reserve(sg, 3, size_t(cells) * 4, st)) != TGM_OK) return rc;
  int32_t *d_nid = static_cast<int32_t *>(sg.p[0]), *d_eid = static_cast<int32_t *>(sg.p[3]);
  int64_t *d_t = static_cast<int64_t *>(sg.p[1]);
  rc = tgm_csr_sample_edges_ids(c, e_lo, e_hi, B, k, 1, d_nid, d_t, d_eid, stream);
  if (rc != TGM_OK) return rc;
  TGM_CUDA(cudaMemcpyAsync(h_out_nid, d_nid, size_t(cells) * 4, cudaMemcpyDeviceToHost, st));
  TGM_CUDA(cudaMemcpyAsync(h_out_t, d_t, size_t(cells) * 8, cudaMemcpyDeviceToHost, st));
  TGM_CUDA(cudaMemcpyAsync(h_out_eid, d_eid, size_t(cells) * 4, cudaMemcpyDeviceToHost, st));
  return TGM_OK;
}

extern "C" int tgm_csr_sample_edges_host_mean(tgm_csr *c, int64_t e_lo, int64_t e_hi, int32_t B,
                                              int32_t k, const int32_t *h_src,
                                              const int32_t *h_dst, const int64_t *h_t,
                                              int32_t *h_out_nid, int64_t *h_out_t,
                                              float *h_out_mean, int slot, tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_edges_host_mean: csr is NULL");
  TGM_REQUIRE(slot >= 0 && slot < TGM_HOST_SLOTS, "tgm_csr_sample_edges_host_mean: bad slot");
  const int64_t nE = e_hi - e_lo, S = 2 * nE, cells = S * k;
  TGM_REQUIRE(nE >= 0 && e_lo >= 0 && e_hi <= c->store->E,
              "tgm_csr_sample_edges_host_mean: [e_lo, e_hi) outside the store");
  if (nE == 0) return TGM_OK;
  TGM_REQUIRE(h_out_mean != nullptr, "tgm_csr_sample_edges_host_mean: h_out_mean is NULL");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  const size_t D = size_t(c->D);
  int rc = upload_slab(c, e_lo, nE, h_src, h_dst, h_t, nullptr, st);
  if (rc != TGM_OK) return rc;
  tgm_csr::Stage &sg = c->stage[slot];
  if (h_out_nid && (rc = stage_reserve(sg, 0, size_t(cells) * 4, st)) != TGM_OK) return rc;
  if (h_out_t && (rc = stage_reserve(sg, 1, size_t(cells) * 8, st)) != TGM_OK) return rc;
  if ((rc = stage_reserve(sg, 4, size_t(S) * D * 4 + 16, st)) != TGM_OK) return rc;
  int32_t *d_nid = h_out_nid ? static_cast<int32_t *>(sg.p[0]) : nullptr;
  int64_t *d_t = h_out_t ? static_cast<int64_t *>(sg.p[1]) : nullptr;
  float *d_mean = static_cast<float *>(sg.p[4]);
  rc = tgm_csr_sample_edges_mean(c, e_lo, e_hi, B, k, 1, d_nid, d_t, nullptr, d_mean, stream);
  if (rc != TGM_OK) return rc;
  if (d_nid) TGM_CUDA(cudaMemcpyAsync(h_out_nid, d_nid, size_t(cells) * 4, cudaMemcpyDeviceToHost, st));
  if (d_t) TGM_CUDA(cudaMemcpyAsync(h_out_t, d_t, size_t(cells) * 8, cudaMemcpyDeviceToHost, st));
  TGM_CUDA(cudaMemcpyAsync(h_out_mean, d_mean, size_t(S) * D * 4, cudaMemcpyDeviceToHost, st));
  return TGM_OK;
}


extern "C" int tgm_csr_sample_uniform(const tgm_csr *c, const int32_t *seeds, int64_t S,
                                      int64_t e_lo, int64_t e_hi, int32_t k, uint64_t rng_seed,
                                      int32_t *out_nid, int64_t *out_t, float *out_x,
                                      tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_uniform: csr is NULL");
  TGM_REQUIRE(c->bs == 1 && c->e_start == 0,
              "tgm_csr_sample_uniform: the adjacency must be built with batch_size 1, e_start 0");
  TGM_REQUIRE(S >= 0 && k >= 1, "tgm_csr_sample_uniform: bad sizes");
  TGM_REQUIRE(0 <= e_lo && e_lo <= e_hi && e_hi <= c->Ew,
              "tgm_csr_sample_uniform: [e_lo, e_hi) outside the store");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && out_nid && out_t, "tgm_csr_sample_uniform: NULL array argument");
  TGM_REQUIRE(c->D == 0 || out_x != nullptr, "tgm_csr_sample_uniform: out_x is NULL but D > 0");
  const int wpb = kUniformThreads / 32;
  const size_t smem = size_t(wpb) * size_t(k) * sizeof(int64_t);
  TGM_REQUIRE(smem <= 48 * 1024, "tgm_csr_sample_uniform: k too large");
  DeviceGuard g(c->device);
  csr_uniform_kernel<false><<<grid_for(S, wpb, 8), kUniformThreads, smem, as_stream(stream)>>>(
      c->entries, c->rowptr, c->store->x, c->N, c->D, seeds, S, e_lo, e_hi, k, rng_seed, out_nid,
      out_t, out_x);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_csr_sample_uniform_time(const tgm_csr *c, const int32_t *seeds, int64_t S,
                                           int64_t t_lo, int has_lo, int64_t t_hi, int has_hi,
                                           int32_t k, uint64_t rng_seed, int32_t *out_nid,
                                           int64_t *out_t, float *out_x, tgm_stream stream) {
  TGM_REQUIRE(c != nullptr, "tgm_csr_sample_uniform_time: csr is NULL");
  TGM_REQUIRE(c->bs == 1 && c->e_start == 0,
              "tgm_csr_sample_uniform_time: the adjacency must be built with batch_size 1, e_start 0");
  TGM_REQUIRE(S >= 0 && k >= 1, "tgm_csr_sample_uniform_time: bad sizes");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && out_nid && out_t, "tgm_csr_sample_uniform_time: NULL array argument");
  TGM_REQUIRE(c->D == 0 || out_x != nullptr, "tgm_csr_sample_uniform_time: out_x is NULL but D > 0");
  const int wpb = kUniformThreads / 32;
  const size_t smem = size_t(wpb) * size_t(k) * sizeof(int64_t);
  TGM_REQUIRE(smem <= 48 * 1024, "tgm_csr_sample_uniform_time: k too large");
  DeviceGuard g(c->device);
  csr_uniform_kernel<true><<<grid_for(S, wpb, 8), kUniformThreads, smem, as_stream(stream)>>>(
      c->entries, c->rowptr, c->store->x, c->N, c->D, seeds, S, has_lo ? t_lo : INT64_MIN,
      has_hi ? t_hi : INT64_MAX, k, rng_seed, out_nid, out_t, out_x);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}


extern "C" int tgm_set_option(const char *name, int value) {
  TGM_REQUIRE(name != nullptr, "tgm_set_option: name is NULL");
  if (std::strcmp(name, "csr_feature_copy") == 0) {
    TGM_REQUIRE(value == 0 || value == 1, "tgm_set_option: csr_feature_copy must be 0 (LSU) or 1 (TMA)");
    g_csr_feature_copy = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "csr_tma_ctas_per_sm") == 0) {
    TGM_REQUIRE(value >= 0 && value <= 32, "tgm_set_option: csr_tma_ctas_per_sm must be in [0, 32]");
    g_csr_tma_ctas_per_sm = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "attn_folded") == 0) {
    TGM_REQUIRE(value == 0 || value == 1, "tgm_set_option: attn_folded must be 0 or 1");
    tgm::g_attn_folded = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "dyg_fused_attn") == 0) {
    TGM_REQUIRE(value == 0 || value == 1, "tgm_set_option: dyg_fused_attn must be 0 or 1");
    tgm::g_dyg_fused_attn = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "tc_bn") == 0) {
    TGM_REQUIRE(value >= 0 && value <= 200, "tgm_set_option: tc_bn must be in [0, 200]");
    tgm::g_tc_bn = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "tc_linear") == 0) {
    TGM_REQUIRE(value >= 0 && value <= 2, "tgm_set_option: tc_linear must be 0, 1 or 2");
    tgm::g_tc_linear = value;
    return TGM_OK;
  }
  if (std::strcmp(name, "trace") == 0) {
    g_trace = value != 0;
    return TGM_OK;
  }
  if (std::strcmp(name, "gemm_fastf32") == 0) {
    TGM_REQUIRE(value == 0 || value == 1, "tgm_set_option: gemm_fastf32 must be 0 (cuBLAS) or 1 (tensor cores)");
    tgm::g_gemm_fastf32 = value;
    return TGM_OK;
  }
  return fail(TGM_ERR_INVALID, std::string("tgm_set_option: unknown option ") + name);
}


extern "C" int tgm_csr_export_ring(const tgm_csr *c, int64_t e_cut, tgm_recency *ring,
                                   tgm_stream stream) {
  TGM_REQUIRE(c != nullptr && ring != nullptr, "tgm_csr_export_ring: NULL handle");
  TGM_REQUIRE(e_cut >= c->e_start && e_cut <= c->e_start + c->Ew,
              "tgm_csr_export_ring: e_cut outside the indexed stream");
  int32_t *ids = nullptr, *wpos = nullptr;
  int64_t *times = nullptr;
  float *feats = nullptr;
  int32_t N = 0, B = 0, D = 0;
  int rc = tgm_recency_dims(ring, &N, &B, &D);
  if (rc != TGM_OK) return rc;
  rc = tgm_recency_state(ring, &ids, &times, &feats, &wpos);
  if (rc != TGM_OK) return rc;
  TGM_REQUIRE(D == c->D, "tgm_csr_export_ring: feature width of ring and store differ");
  TGM_REQUIRE(N >= c->N, "tgm_csr_export_ring: ring has fewer nodes than the store");
  DeviceGuard g(c->device);
  cudaStream_t st = as_stream(stream);
  rc = tgm_recency_reset(ring, stream);  // nodes beyond the store's id range stay empty
  if (rc != TGM_OK) return rc;
  if (c->n == 0) return TGM_OK;
  csr_export_ring_kernel<<<grid_for(c->N, 8, 8), 256, 0, st>>>(
      c->entries, c->rowptr, c->store->x, c->N, c->D, e_cut, B, ids, times, feats, wpos);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}


// accessor for other translation units (uniform_exact.cu)
tgm::CsrView tgm::csr_view(const tgm_csr *c) {
  return tgm::CsrView{c->device, c->entries, c->rowptr, c->store ? c->store->x : nullptr,
                      c->N,      c->D,       c->bs,     c->e_start, c->Ew};
}
