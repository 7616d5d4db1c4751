// Stateful recency sampler: per-node ring buffers on the device.
// Replaces RecencyNeighborHook's state machine (reference tgm-team/tgm @ 5183dc9,
// tgm/hooks/neighbors/recency.py): state :93-97/:410-416, reset_state :111-117,
// _get_recency_neighbors :239-321 (~25 eager ops + an O(N*B) min() scan per call), _update
// :323-399 (argsort + 15 eager ops).  Here a query is one launch (a warp per seed, only the k
// needed feature rows are read) and an update is two launches.
#include <cub/cub.cuh>

#include <algorithm>
#include <new>

#include "common.cuh"
#include "tma.cuh"

using namespace tgm;

struct tgm_recency {
  int32_t N = 0, B = 0, D = 0;
  int device = -1;
  int32_t *ids = nullptr;    // [N, B]
  int64_t *times = nullptr;  // [N, B]
  float *feats = nullptr;    // [N, B, D]
  int32_t *wpos = nullptr;   // [N]
  // update scratch (grown on demand)
  int64_t *dest = nullptr;  // [cap] ring row (node*B+slot) an entry lands in, -1 = dropped
  int32_t *inc = nullptr;   // [cap] write_pos increment carried by the last entry of each node
  int64_t cap = 0;
  ~tgm_recency() {
    if (device >= 0) {
      DeviceGuard g(device);
      cudaFree(ids);
      cudaFree(times);
      cudaFree(feats);
      cudaFree(wpos);
      cudaFree(dest);
      cudaFree(inc);
    }
  }
};

namespace {

constexpr int kQueryThreads = 256;

// ---- query (recency.py:239-321) --------------------------------------------------------------
// One warp per seed.  Unrolled ring position j (oldest .. newest) lives in slot (wp + j) % B.
template <bool VEC4>
__global__ void __launch_bounds__(kQueryThreads)
ring_query_kernel(const int32_t *__restrict__ ids, const int64_t *__restrict__ times,
                  const float *__restrict__ feats, const int32_t *__restrict__ wpos, int N, int B,
                  int D, const int32_t *__restrict__ seeds, const int64_t *__restrict__ tq,
                  int64_t S, int k, int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                  float *__restrict__ out_x) {
  extern __shared__ int32_t s_slot[];  // [warps][k]: ring slot feeding each output column, -1 = pad
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int32_t *my = s_slot + warp * k;
  for (int64_t s = int64_t(blockIdx.x) * wpb + warp; s < S; s += int64_t(gridDim.x) * wpb) {
    int v = seeds[s];
    if (v < 0) v += N;  // torch negative indexing: the padded id -1 reads row N-1 (recency.py:256)
    const bool in_range = (v >= 0 && v < N);
    const int64_t q = tq[s];
    const int64_t row = int64_t(in_range ? v : 0) * B;
    const int wp = in_range ? int(uint32_t(wpos[v]) % uint32_t(B)) : 0;
    int last = -1;  // right-most unrolled position with id != -1 && time < tq (:267-281)
    if (in_range) {
      for (int base = 0; base < B; base += 32) {
        const int j = base + lane;
        bool ok = false;
        if (j < B) {
          int slot = wp + j;
          if (slot >= B) slot -= B;
          ok = (ids[row + slot] != TGM_PADDED_NODE_ID) && (times[row + slot] < q);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) last = base + 31 - __clz(m);
      }
    }
    // k-window ending at `last`, right-aligned (:287-319)
    for (int c = lane; c < k; c += 32) {
      const int p = last - (k - 1 - c);
      int slot = -1;
      int32_t id = TGM_PADDED_NODE_ID;
      int64_t tt = 0;
      if (p >= 0) {
        slot = wp + p;
        if (slot >= B) slot -= B;
        id = ids[row + slot];
        tt = times[row + slot];
      }
      my[c] = slot;
      out_nid[s * k + c] = id;
      out_t[s * k + c] = tt;
    }
    __syncwarp();
    if (D > 0) {
      if (VEC4) {
        const int D4 = D >> 2;
        const float4 *f4 = reinterpret_cast<const float4 *>(feats);
        float4 *o4 = reinterpret_cast<float4 *>(out_x) + s * int64_t(k) * D4;
        const int total = k * D4;
        for (int i = lane; i < total; i += 32) {
          const int c = i / D4, d = i - c * D4;
          const int slot = my[c];
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (slot >= 0) val = __ldg(f4 + (row + slot) * D4 + d);
          o4[i] = val;
        }
      } else {
        float *o = out_x + s * int64_t(k) * D;
        const int total = k * D;
        for (int i = lane; i < total; i += 32) {
          const int c = i / D, d = i - c * D;
          const int slot = my[c];
          o[i] = slot >= 0 ? __ldg(feats + (row + slot) * D + d) : 0.f;
        }
      }
    }
    __syncwarp();
  }
}

// The same query with the feature block moved by bulk copies (the structure of
// csr_sample_tma_kernel): a warp takes 32 consecutive seeds -- seed ids, query times and write
// positions are loaded lane-parallel -- then walks them; lane j holds ring position j (oldest ..
// newest) of the current seed, loaded a seed ahead; one ballot finds the right-most admissible
// entry, shuffles route ids / times to the output columns, and lane 0 moves the k-window's
// feature rows global -> shared -> global through a 4-stage ring (`cp.async.bulk` + mbarrier; the
// window is one or, where it wraps around the ring, two contiguous runs of rows; padding comes
// from a zeroed block), loads running two seeds ahead of stores.  Needs B <= 32, D % 4 == 0 and
// k * D * 4 <= 4096.
constexpr int kRqStages = 4, kRqLag = 2, kRqMaxStageBytes = 4096;
struct RqMeta {
  float *dst;
  uint32_t nbytes;    // bulk-loaded feature bytes; 0 = nothing was loaded
  uint32_t padbytes;  // zero bytes in front of them; bit 31 = mbarrier phase parity
};

__global__ void __launch_bounds__(kQueryThreads)
ring_query_tma_kernel(const int32_t *__restrict__ ids, const int64_t *__restrict__ times,
                      const float *__restrict__ feats, const int32_t *__restrict__ wpos, int N,
                      int B, int D, const int32_t *__restrict__ seeds,
                      const int64_t *__restrict__ tq, int64_t S, int k,
                      int32_t *__restrict__ out_nid, int64_t *__restrict__ out_t,
                      float *__restrict__ out_x, int stage_bytes) {
  extern __shared__ __align__(128) unsigned char rq_smem[];
  constexpr int W = kQueryThreads >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // layout: zero block | W * stages | W * stages mbarriers | W * stages metas
  unsigned char *zero = rq_smem;
  unsigned char *stages = zero + stage_bytes + size_t(warp) * kRqStages * stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(rq_smem + stage_bytes +
                                                size_t(W) * kRqStages * stage_bytes) +
                   warp * kRqStages;
  RqMeta *meta = reinterpret_cast<RqMeta *>(rq_smem + stage_bytes +
                                            size_t(W) * kRqStages * (stage_bytes + 8)) +
                 warp * kRqStages;
  for (int i = threadIdx.x * 4; i < stage_bytes; i += kQueryThreads * 4)
    *reinterpret_cast<uint32_t *>(zero + i) = 0u;
  if (lane == 0)
    for (int s = 0; s < kRqStages; ++s) mbar_init(smem_u32(bars + s), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zero_s = smem_u32(zero), stage_s = smem_u32(stages), bar_s = smem_u32(bars);
  const uint32_t row_bytes = uint32_t(D) * 4u;

  auto retire = [&](uint32_t q) {  // lane 0: seed number q of this warp has landed -> store it
    const int st = int(q % kRqStages);
    const RqMeta m = meta[st];
    const uint32_t padbytes = m.padbytes & 0x7fffffffu;
    if (m.nbytes) {
      mbar_wait(bar_s + st * 8, m.padbytes >> 31);
      bulk_s2g(reinterpret_cast<unsigned char *>(m.dst) + padbytes, stage_s + st * stage_bytes,
               m.nbytes);
    }
    if (padbytes) bulk_s2g(m.dst, zero_s, padbytes);
    bulk_commit();
  };
  // ring position `lane` of node v: {id, time}; id = padded for lanes >= B and out-of-range seeds
  auto load_pos = [&](int v, int wp, int32_t &id, int64_t &tt) {
    id = TGM_PADDED_NODE_ID;
    tt = 0;
    if (v >= 0 && lane < B) {
      int slot = wp + lane;
      if (slot >= B) slot -= B;
      id = __ldg(ids + int64_t(v) * B + slot);
      tt = __ldg(times + int64_t(v) * B + slot);
    }
  };

  const int64_t nchunks = (S + 31) >> 5;
  uint32_t g = 0, phases = 0;
  for (int64_t ch = int64_t(blockIdx.x) * W + warp; ch < nchunks; ch += int64_t(gridDim.x) * W) {
    const int64_t s_base = ch << 5, s = s_base + lane;
    int my_v = -1, my_wp = 0;  // my_v < 0: all-padding row
    int64_t my_q = 0;
    if (s < S) {
      int v = __ldg(seeds + s);
      if (v < 0) v += N;  // torch negative indexing: the padded id -1 reads row N-1 (recency.py:256)
      my_q = __ldg(tq + s);
      if (v >= 0 && v < N) {
        my_v = v;
        my_wp = int(uint32_t(__ldg(wpos + v)) % uint32_t(B));
      }
    }
    const int nseeds = S - s_base < 32 ? int(S - s_base) : 32;
    int v = __shfl_sync(0xffffffffu, my_v, 0), wp = __shfl_sync(0xffffffffu, my_wp, 0);
    int32_t cur_id;
    int64_t cur_t;
    load_pos(v, wp, cur_id, cur_t);
    for (int i = 0; i < nseeds; ++i, ++g) {
      const int64_t q = shfl_i64(my_q, i);
      const int nxt = i + 1 < nseeds ? i + 1 : i;
      const int v_n = __shfl_sync(0xffffffffu, my_v, nxt), wp_n = __shfl_sync(0xffffffffu, my_wp, nxt);
      int32_t ahead_id;
      int64_t ahead_t;
      load_pos(v_n, wp_n, ahead_id, ahead_t);
      const unsigned m = __ballot_sync(0xffffffffu, cur_id != TGM_PADDED_NODE_ID && cur_t < q);
      const int last = m ? 31 - __clz(m) : -1;  // (:267-281)
      const int nvalid = last + 1 < k ? last + 1 : k;
      const int first = last + 1 - nvalid, pad = k - nvalid;
      const int srcl = (first + lane - pad) & 31;
      const int32_t id = __shfl_sync(0xffffffffu, cur_id, srcl);
      const int64_t tt = shfl_i64(cur_t, srcl);
      const int64_t sg = s_base + i;
      if (lane < k) {
        const bool ok = lane >= pad;
        out_nid[sg * k + lane] = ok ? id : TGM_PADDED_NODE_ID;
        out_t[sg * k + lane] = ok ? tt : 0;
      }
      if (lane == 0) {
        const int st = int(g % kRqStages);
        bulk_wait_read<kRqStages - kRqLag - 1>();  // the store that last read this stage is done
        RqMeta mt;
        mt.dst = out_x + sg * int64_t(k) * D;
        mt.nbytes = uint32_t(nvalid) * row_bytes;
        mt.padbytes = uint32_t(pad) * row_bytes;
        if (mt.nbytes) {
          mt.padbytes |= ((phases >> st) & 1u) << 31;
          phases ^= 1u << st;
        }
        meta[st] = mt;
        if (mt.nbytes) {
          // unrolled positions first .. last live in slots (wp + p) % B: one run, or two when
          // the window wraps around the end of the ring
          int s0 = wp + first;
          if (s0 >= B) s0 -= B;
          const int len1 = nvalid < B - s0 ? nvalid : B - s0;
          const float *rowp = feats + int64_t(v) * B * D;
          mbar_expect_tx(bar_s + st * 8, mt.nbytes);
          bulk_g2s(stage_s + st * stage_bytes, rowp + int64_t(s0) * D, uint32_t(len1) * row_bytes,
                   bar_s + st * 8);
          if (len1 < nvalid)
            bulk_g2s(stage_s + st * stage_bytes + uint32_t(len1) * row_bytes, rowp,
                     uint32_t(nvalid - len1) * row_bytes, bar_s + st * 8);
        }
        if (g >= uint32_t(kRqLag)) retire(g - kRqLag);
      }
      cur_id = ahead_id;
      cur_t = ahead_t;
      v = v_n;
      wp = wp_n;
    }
  }
  if (lane == 0) {  // drain
    for (uint32_t q = g >= uint32_t(kRqLag) ? g - kRqLag : 0; q < g; ++q) retire(q);
    bulk_wait_read<0>();
  }
}

// ---- update (recency.py:323-399) -------------------------------------------------------------
// Entry i of the concatenation [src->dst entries | dst->src entries] (:339-342).
struct UpdEntry {
  int32_t node, nbr;
  int64_t t;
  int64_t e;
};
__device__ __forceinline__ UpdEntry load_entry(const int32_t *src, const int32_t *dst,
                                               const int64_t *t, int64_t Eb, int64_t i) {
  UpdEntry u;
  if (i < Eb) {
    u.e = i;
    u.node = src[i];
    u.nbr = dst[i];
  } else {
    u.e = i - Eb;
    u.node = dst[u.e];
    u.nbr = src[u.e];
  }
  u.t = t[u.e];
  return u;
}

constexpr int kRankThreads = 256;

// Phase A: every entry finds its rank among the entries of the same node in the order of the
// reference's stable sort -- (time, position in the concatenation) (:347-349) -- and the node's
// entry count.  The last B are written to (write_pos + j) % B (:373,:389-393).  O(n^2 / tile)
// over a batch of a few hundred edges; the tiles of (node, time) are staged in shared memory.
__global__ void __launch_bounds__(kRankThreads)
ring_update_rank_kernel(int32_t *__restrict__ ids, int64_t *__restrict__ times,
                        const int32_t *__restrict__ wpos, int N, int B,
                        const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                        const int64_t *__restrict__ t, int64_t Eb, int64_t n,
                        int64_t *__restrict__ dest, int32_t *__restrict__ inc) {
  __shared__ int32_t s_node[kRankThreads];
  __shared__ int64_t s_t[kRankThreads];
  const int64_t i = int64_t(blockIdx.x) * kRankThreads + threadIdx.x;
  UpdEntry me{};
  me.node = -2;
  if (i < n) me = load_entry(src, dst, t, Eb, i);
  int rank = 0, cnt = 0;
  for (int64_t base = 0; base < n; base += kRankThreads) {
    const int64_t j = base + threadIdx.x;
    if (j < n) {
      const UpdEntry o = load_entry(src, dst, t, Eb, j);
      s_node[threadIdx.x] = o.node;
      s_t[threadIdx.x] = o.t;
    } else {
      s_node[threadIdx.x] = -3;
    }
    __syncthreads();
    const int lim = int(n - base < kRankThreads ? n - base : kRankThreads);
    for (int u = 0; u < lim; ++u) {
      if (s_node[u] == me.node) {
        ++cnt;
        const int64_t ot = s_t[u];
        rank += (ot < me.t) || (ot == me.t && base + u < i);
      }
    }
    __syncthreads();
  }
  if (i >= n) return;
  int64_t d = -1;
  int32_t add = 0;
  if (me.node >= 0 && me.node < N) {
    const int first_kept = cnt > B ? cnt - B : 0;
    if (rank >= first_kept) {
      const int wp = int(uint32_t(wpos[me.node]) % uint32_t(B));
      const int slot = (wp + (rank - first_kept)) % B;
      d = int64_t(me.node) * B + slot;
      ids[d] = me.nbr;
      times[d] = me.t;
    }
    if (rank == cnt - 1) add = cnt < B ? cnt : B;  // :397-399 counts the kept entries only
  }
  dest[i] = d;
  inc[i] = add;
}

// Bulk path (n > kRankDirectMax): the same ranking from two stable radix sorts -- by time, then by
// node -- so entries of a node end up contiguous in (time, position) order.
constexpr int64_t kRankDirectMax = 8192;

__global__ void bulk_time_keys_kernel(const int64_t *__restrict__ t, int64_t Eb, int64_t n,
                                      uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    // order-preserving map of int64 to uint64 (times are >= 0 in valid streams, but stay total)
    keys[i] = uint64_t(t[i < Eb ? i : i - Eb]) ^ 0x8000000000000000ull;
    vals[i] = uint32_t(i);
  }
}
__global__ void bulk_node_keys_kernel(const int32_t *__restrict__ src,
                                      const int32_t *__restrict__ dst, int64_t Eb, int64_t n,
                                      const uint32_t *__restrict__ vals,
                                      uint32_t *__restrict__ keys) {
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = vals[j];
    keys[j] = uint32_t(i < Eb ? src[i] : dst[i - Eb]);
  }
}
// sorted position j -> rank and count inside its node run, then the same placement rule as
// ring_update_rank_kernel
__global__ void bulk_place_kernel(int32_t *__restrict__ ids, int64_t *__restrict__ times,
                                  const int32_t *__restrict__ wpos, int N, int B,
                                  const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                  const int64_t *__restrict__ t, int64_t Eb, int64_t n,
                                  const uint32_t *__restrict__ keys,
                                  const uint32_t *__restrict__ vals, int64_t *__restrict__ dest,
                                  int32_t *__restrict__ inc) {
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t node = keys[j];
    int64_t lo = 0, hi = j;  // first position of the run
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < node) lo = mid + 1; else hi = mid;
    }
    const int64_t start = lo;
    lo = j + 1, hi = n;  // one past the run
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] <= node) lo = mid + 1; else hi = mid;
    }
    const int64_t cnt = lo - start, rank = j - start;
    const int64_t i = vals[j];
    const UpdEntry me = load_entry(src, dst, t, Eb, i);
    int64_t d = -1;
    int32_t add = 0;
    if (me.node >= 0 && me.node < N) {
      const int64_t first_kept = cnt > B ? cnt - B : 0;
      if (rank >= first_kept) {
        const int wp = int(uint32_t(wpos[me.node]) % uint32_t(B));
        const int slot = int((wp + (rank - first_kept)) % B);
        d = int64_t(me.node) * B + slot;
        ids[d] = me.nbr;
        times[d] = me.t;
      }
      if (rank == cnt - 1) add = int32_t(cnt < B ? cnt : B);
    }
    dest[i] = d;
    inc[i] = add;
  }
}

// Phase B: copy the kept feature rows (a warp per entry, coalesced) and advance write_pos.
__global__ void __launch_bounds__(256)
ring_update_commit_kernel(float *__restrict__ feats, int32_t *__restrict__ wpos, int D,
                          const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                          const float *__restrict__ x, int64_t Eb, int64_t n,
                          const int64_t *__restrict__ dest, const int32_t *__restrict__ inc) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t i = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); i < n;
       i += int64_t(gridDim.x) * wpb) {
    const int64_t d = dest[i];
    const int64_t e = i < Eb ? i : i - Eb;
    if (d >= 0 && D > 0) {
      float *o = feats + d * D;
      if (x) {
        const float *r = x + e * D;
        for (int c = lane; c < D; c += 32) o[c] = __ldg(r + c);
      } else {
        for (int c = lane; c < D; c += 32) o[c] = 0.f;  // missing edge_x pushes zeros (:325-329)
      }
    }
    if (lane == 0) {
      const int32_t a = inc[i];
      if (a > 0) {
        const int32_t node = i < Eb ? src[e] : dst[e];
        wpos[node] += a;  // exactly one entry per node carries a non-zero increment
      }
    }
  }
}

// hop-0 seeds of the standard hook configuration: [edge_src | edge_dst], [edge_time | edge_time]
__global__ void __launch_bounds__(256)
step_seeds_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                  const int64_t *__restrict__ t, int64_t n, int32_t *__restrict__ seeds,
                  int64_t *__restrict__ times) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const bool second = i >= n;
    const int64_t e = second ? i - n : i;
    seeds[i] = second ? dst[e] : src[e];
    times[i] = t[e];
  }
}

}  // namespace

extern "C" int tgm_recency_create(tgm_recency **out, int32_t num_nodes, int32_t B, int32_t D,
                                  int device) {
  TGM_REQUIRE(out != nullptr, "tgm_recency_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(num_nodes > 0, "tgm_recency_create: num_nodes must be > 0");
  TGM_REQUIRE(B > 0, "tgm_recency_create: B must be > 0");
  TGM_REQUIRE(D >= 0, "tgm_recency_create: D must be >= 0");
  TGM_REQUIRE(device >= 0, "tgm_recency_create: a CUDA device is required (no CPU fallback)");
  DeviceGuard g(device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_recency_create: cannot select device");
  tgm_recency *h = new (std::nothrow) tgm_recency();
  if (!h) return fail(TGM_ERR_OOM, "tgm_recency_create: host allocation failed");
  h->N = num_nodes;
  h->B = B;
  h->D = D;
  h->device = device;
  const size_t nb = size_t(num_nodes) * size_t(B);
  cudaError_t e = cudaMalloc(&h->ids, nb * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->times, nb * 8);
  if (e == cudaSuccess && D > 0) e = cudaMalloc(&h->feats, nb * size_t(D) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->wpos, size_t(num_nodes) * 4);
  if (e != cudaSuccess) {
    delete h;
    return cuda_fail(e, "ring allocation", __FILE__, __LINE__);
  }
  int rc = tgm_recency_reset(h, nullptr);
  if (rc == TGM_OK) {
    e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) rc = cuda_fail(e, "ring reset", __FILE__, __LINE__);
  }
  if (rc != TGM_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return TGM_OK;
}

extern "C" void tgm_recency_destroy(tgm_recency *h) { delete h; }

extern "C" int tgm_recency_reset(tgm_recency *h, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_reset: handle is NULL");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  const size_t nb = size_t(h->N) * size_t(h->B);
  TGM_CUDA(cudaMemsetAsync(h->ids, 0xFF, nb * 4, st));  // int32 -1 == PADDED_NODE_ID
  TGM_CUDA(cudaMemsetAsync(h->times, 0, nb * 8, st));
  if (h->D > 0) TGM_CUDA(cudaMemsetAsync(h->feats, 0, nb * size_t(h->D) * 4, st));
  TGM_CUDA(cudaMemsetAsync(h->wpos, 0, size_t(h->N) * 4, st));
  return TGM_OK;
}

extern "C" int tgm_recency_query(const tgm_recency *h, const int32_t *seeds, const int64_t *tq,
                                 int64_t S, int32_t k, int32_t *out_nid, int64_t *out_t,
                                 float *out_x, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_query: handle is NULL");
  TGM_REQUIRE(S >= 0, "tgm_recency_query: S must be >= 0");
  TGM_REQUIRE(k >= 1 && k <= h->B, "tgm_recency_query: k must be in [1, B]");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && tq && out_nid && out_t, "tgm_recency_query: NULL array argument");
  TGM_REQUIRE(h->D == 0 || out_x != nullptr, "tgm_recency_query: out_x is NULL but D > 0");
  DeviceGuard g(h->device);
  const int wpb = kQueryThreads / 32;
  // feature block moved by bulk copies where the shapes allow (see ring_query_tma_kernel)
  if (h->D > 0 && h->D % 4 == 0 && h->B <= 32 && k * h->D * 4 <= kRqMaxStageBytes &&
      aligned16(h->feats) && aligned16(out_x) && S >= 4096) {
    const int stage_bytes = k * h->D * 4;
    const size_t smem = size_t(stage_bytes) + size_t(wpb) * kRqStages * (size_t(stage_bytes) + 8 + 16);
    if (smem > 48 * 1024)
      TGM_CUDA(cudaFuncSetAttribute(ring_query_tma_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int per_sm = int(std::max<size_t>(1, std::min<size_t>(6, (220 * 1024) / (smem + 1024))));
    ring_query_tma_kernel<<<grid_for((S + 31) / 32, wpb, per_sm), kQueryThreads, smem,
                            as_stream(stream)>>>(h->ids, h->times, h->feats, h->wpos, h->N, h->B,
                                                 h->D, seeds, tq, S, k, out_nid, out_t, out_x,
                                                 stage_bytes);
    TGM_LAUNCH_CHECK();
    return TGM_OK;
  }
  const size_t smem = size_t(wpb) * size_t(k) * sizeof(int32_t);
  TGM_REQUIRE(smem <= 48 * 1024, "tgm_recency_query: k too large");
  const int grid = grid_for(S, wpb, 8);
  const bool vec4 = h->D > 0 && (h->D % 4 == 0) && aligned16(h->feats) && aligned16(out_x);
  if (vec4)
    ring_query_kernel<true><<<grid, kQueryThreads, smem, as_stream(stream)>>>(
        h->ids, h->times, h->feats, h->wpos, h->N, h->B, h->D, seeds, tq, S, k, out_nid, out_t,
        out_x);
  else
    ring_query_kernel<false><<<grid, kQueryThreads, smem, as_stream(stream)>>>(
        h->ids, h->times, h->feats, h->wpos, h->N, h->B, h->D, seeds, tq, S, k, out_nid, out_t,
        out_x);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_recency_update(tgm_recency *h, const int32_t *src, const int32_t *dst,
                                  const int64_t *t, const float *x, int64_t Eb, int directed,
                                  tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_update: handle is NULL");
  TGM_REQUIRE(Eb >= 0, "tgm_recency_update: Eb must be >= 0");
  if (Eb == 0) return TGM_OK;
  TGM_REQUIRE(src && dst && t, "tgm_recency_update: NULL array argument");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  const int64_t n = directed ? Eb : 2 * Eb;
  TGM_REQUIRE(n < (int64_t(1) << 31), "tgm_recency_update: batch too large");
  if (n > h->cap) {
    // growing the scratch is the only synchronising path; steady state never reallocates
    TGM_CUDA(cudaStreamSynchronize(st));
    cudaFree(h->dest);
    cudaFree(h->inc);
    h->dest = nullptr;
    h->inc = nullptr;
    h->cap = 0;
    const int64_t cap = n < 1024 ? 1024 : n + n / 2;
    TGM_CUDA(cudaMalloc(&h->dest, size_t(cap) * 8));
    TGM_CUDA(cudaMalloc(&h->inc, size_t(cap) * 4));
    h->cap = cap;
  }
  if (n <= kRankDirectMax) {  // loader batches: O(n^2 / tile) ranking over shared-memory tiles
    const int gridA = int((n + kRankThreads - 1) / kRankThreads);
    ring_update_rank_kernel<<<gridA, kRankThreads, 0, st>>>(h->ids, h->times, h->wpos, h->N, h->B,
                                                            src, dst, t, Eb, n, h->dest, h->inc);
    TGM_LAUNCH_CHECK();
  } else {  // bulk load: sort-based ranking (scratch is allocated per call; synchronises)
    uint64_t *kt_a = nullptr, *kt_b = nullptr;
    uint32_t *v_a = nullptr, *v_b = nullptr, *kn_a = nullptr, *kn_b = nullptr;
    void *tmp = nullptr;
    auto done = [&](int rc) {
      cudaStreamSynchronize(st);
      cudaFree(kt_a), cudaFree(kt_b), cudaFree(v_a), cudaFree(v_b), cudaFree(kn_a), cudaFree(kn_b);
      cudaFree(tmp);
      return rc;
    };
#define BULK_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return done(cuda_fail(_e, #expr, __FILE__, __LINE__)); \
  } while (0)
    const size_t nn = size_t(n);
    BULK_CUDA(cudaMalloc(&kt_a, nn * 8));
    BULK_CUDA(cudaMalloc(&kt_b, nn * 8));
    BULK_CUDA(cudaMalloc(&v_a, nn * 4));
    BULK_CUDA(cudaMalloc(&v_b, nn * 4));
    BULK_CUDA(cudaMalloc(&kn_a, nn * 4));
    BULK_CUDA(cudaMalloc(&kn_b, nn * 4));
    size_t b1 = 0, b2 = 0;
    BULK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, kt_a, kt_b, v_a, v_b, n, 0, 64, st));
    BULK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, kn_a, kn_b, v_b, v_a, n, 0, 32, st));
    size_t tmp_bytes = b1 > b2 ? b1 : b2;
    BULK_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    bulk_time_keys_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(t, Eb, n, kt_a, v_a);
    BULK_CUDA(cudaGetLastError());
    BULK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kt_a, kt_b, v_a, v_b, n, 0, 64, st));
    bulk_node_keys_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(src, dst, Eb, n, v_b, kn_a);
    BULK_CUDA(cudaGetLastError());
    BULK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kn_a, kn_b, v_b, v_a, n, 0, 32, st));
    bulk_place_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(h->ids, h->times, h->wpos, h->N, h->B,
                                                           src, dst, t, Eb, n, kn_b, v_a, h->dest,
                                                           h->inc);
    BULK_CUDA(cudaGetLastError());
#undef BULK_CUDA
    int rc = done(TGM_OK);
    if (rc != TGM_OK) return rc;
  }
  const int gridB = grid_for(n, 8, 8);
  ring_update_commit_kernel<<<gridB, 256, 0, st>>>(h->feats, h->wpos, h->D, src, dst, x, Eb, n,
                                                   h->dest, h->inc);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_recency_step(tgm_recency *h, const int32_t *src, const int32_t *dst,
                                const int64_t *t, const float *x, int64_t Eb, int directed,
                                int32_t num_hops, const int32_t *num_nbrs, int32_t *seed_nids0,
                                int64_t *seed_times0, int32_t *const *out_nid,
                                int64_t *const *out_t, float *const *out_x, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_step: handle is NULL");
  TGM_REQUIRE(Eb > 0 && num_hops >= 1, "tgm_recency_step: needs a non-empty batch and >= 1 hop");
  TGM_REQUIRE(src && dst && t && num_nbrs && seed_nids0 && seed_times0 && out_nid && out_t && out_x,
              "tgm_recency_step: NULL argument");
  DeviceGuard g(h->device);
  step_seeds_kernel<<<grid_for(2 * Eb, 256, 8), 256, 0, as_stream(stream)>>>(src, dst, t, Eb,
                                                                           seed_nids0, seed_times0);
  TGM_LAUNCH_CHECK();
  const int32_t *seeds = seed_nids0;
  const int64_t *times = seed_times0;
  int64_t S = 2 * Eb;
  for (int32_t hop = 0; hop < num_hops; ++hop) {  // query BEFORE update (recency.py:161-163)
    const int32_t k = num_nbrs[hop];
    int rc = tgm_recency_query(h, seeds, times, S, k, out_nid[hop], out_t[hop], out_x[hop], stream);
    if (rc) return rc;
    seeds = out_nid[hop];  // hop h+1 seeds = flattened hop-h neighbours (recency.py:141-143)
    times = out_t[hop];
    S *= k;
  }
  return tgm_recency_update(h, src, dst, t, x, Eb, directed, stream);
}

extern "C" int tgm_recency_state(const tgm_recency *h, int32_t **ids, int64_t **times,
                                 float **feats, int32_t **write_pos) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_state: handle is NULL");
  if (ids) *ids = h->ids;
  if (times) *times = h->times;
  if (feats) *feats = h->feats;
  if (write_pos) *write_pos = h->wpos;
  return TGM_OK;
}


extern "C" int tgm_recency_dims(const tgm_recency *h, int32_t *num_nodes, int32_t *B, int32_t *D) {
  TGM_REQUIRE(h != nullptr, "tgm_recency_dims: handle is NULL");
  if (num_nodes) *num_nodes = h->N;
  if (B) *B = h->B;
  if (D) *D = h->D;
  return TGM_OK;
}
