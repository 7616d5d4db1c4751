// Batch de-duplication: sorted unique node ids of a batch and the global -> local index map.
//
// Replaces (reference tgm-team/tgm @ 5183dc9) tgm/hooks/dedup.py:35-67: a boolean-mask gather per
// hop, torch.cat, torch.unique(sorted=True) (a device sort) and, per lookup, torch.searchsorted.
// Node ids are dense in [0, num_nodes), so the set is a bitmap: no sort, and the rank of an id
//     local(v) = #{u in set : u < v} = prefix[v >> 5] + popc(bitmap[v >> 5] & ((1 << (v & 31)) - 1))
// is exactly searchsorted(unique, v) (left) for EVERY v, member or not -- an O(1) lookup.
//   1. memset bitmap (num_nodes / 8 bytes: 125 KB at 1M nodes)
//   2. dedup_mark: atomicOr one bit per id; padded (-1) neighbour slots are skipped in flight
//      (no separate compaction, no host sync per hop); a -1 in a seed array is kept as an element,
//      as torch.unique would
//   3. exclusive scan of the per-word popcounts (cub::DeviceScan, one decoupled-look-back launch)
//   4. dedup_emit: every word writes its ids at its prefix -> ascending order for free
// The caller owns bitmap/prefix (one pair per batch, so an older batch's map stays valid).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

using namespace tgm;

namespace {

constexpr int kMaxParts = 8;
struct Parts {
  const int32_t *p[kMaxParts];
  int64_t end[kMaxParts];  // cumulative sizes
  int skip_padded[kMaxParts];
  int n;
};

// flags[0] |= 1 when a kept -1 was seen, |= 2 when an id is outside [-1, num_nodes)
__global__ void __launch_bounds__(256)
dedup_mark_kernel(Parts parts, int64_t total, int32_t num_nodes, uint32_t *__restrict__ bitmap,
                  int32_t *__restrict__ flags) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    int q = 0;
    while (q + 1 < parts.n && i >= parts.end[q]) ++q;
    const int64_t off = i - (q ? parts.end[q - 1] : 0);
    const int32_t v = __ldg(parts.p[q] + off);
    if (v == TGM_PADDED_NODE_ID) {
      if (!parts.skip_padded[q]) atomicOr(flags, 1);
    } else if (v < 0 || v >= num_nodes) {
      atomicOr(flags, 2);
    } else {
      atomicOr(bitmap + (v >> 5), 1u << (v & 31));
    }
  }
}

struct PopcOp {
  __host__ __device__ __forceinline__ int32_t operator()(const uint32_t &w) const {
#ifdef __CUDA_ARCH__
    return __popc(w);
#else
    return __builtin_popcount(w);
#endif
  }
};

__global__ void __launch_bounds__(256)
dedup_emit_kernel(const uint32_t *__restrict__ bitmap, const int32_t *__restrict__ prefix,
                  int64_t words, const int32_t *__restrict__ flags, int32_t *__restrict__ out,
                  int64_t *__restrict__ out_count) {
  const int shift = flags[0] & 1;
  for (int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < words;
       w += int64_t(gridDim.x) * blockDim.x) {
    uint32_t bits = bitmap[w];
    int32_t at = prefix[w] + shift;
    while (bits) {
      const int b = __ffs(bits) - 1;
      out[at++] = int32_t(w * 32 + b);
      bits &= bits - 1;
    }
    if (w == 0) {
      if (shift) out[0] = TGM_PADDED_NODE_ID;
      // prefix[words] is the scan's total (the bitmap carries one trailing zero word)
      out_count[0] = (flags[0] & 2) ? int64_t(-1) : int64_t(prefix[words]) + shift;
    }
  }
}

__global__ void __launch_bounds__(256)
dedup_map_kernel(const uint32_t *__restrict__ bitmap, const int32_t *__restrict__ prefix,
                 const int32_t *__restrict__ flags, int32_t num_nodes, int64_t words,
                 const int32_t *__restrict__ ids, int64_t n, int32_t *__restrict__ out) {
  const int shift = flags[0] & 1;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int32_t v = ids[i];
    int32_t r;
    if (v < 0) r = (v == TGM_PADDED_NODE_ID) ? 0 : 0;  // searchsorted(unique, negative) == 0
    else if (v >= num_nodes) r = prefix[words] + shift;
    else r = prefix[v >> 5] + __popc(bitmap[v >> 5] & ((1u << (v & 31)) - 1u)) + shift;
    out[i] = r;
  }
}

inline int64_t dedup_words(int32_t num_nodes) { return (int64_t(num_nodes) + 31) / 32; }

}  // namespace

extern "C" int tgm_dedup_sizes(int32_t num_nodes, int64_t *bitmap_words, int64_t *prefix_len,
                               int64_t *tmp_bytes) {
  TGM_REQUIRE(num_nodes > 0 && bitmap_words && prefix_len && tmp_bytes, "tgm_dedup_sizes: bad arguments");
  const int64_t words = dedup_words(num_nodes);
  *bitmap_words = words + 1;  // one trailing zero word so the scan also yields the total
  *prefix_len = words + 1;
  size_t bytes = 0;
  cub::TransformInputIterator<int32_t, PopcOp, const uint32_t *> it(nullptr, PopcOp());
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (int32_t *)nullptr, int(words + 1));
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceScan size query", __FILE__, __LINE__);
  *tmp_bytes = int64_t(bytes) + 16;  // + the flag word (kept 16-byte aligned at the front)
  return TGM_OK;
}

extern "C" int tgm_dedup_unique(const int32_t *const *parts, const int64_t *sizes,
                                const int32_t *skip_padded, int32_t n_parts, int32_t num_nodes,
                                uint32_t *bitmap, int32_t *prefix, void *tmp, int64_t tmp_bytes,
                                int32_t *out_unique, int64_t *out_count, tgm_stream stream) {
  TGM_REQUIRE(n_parts >= 1 && n_parts <= kMaxParts, "tgm_dedup_unique: 1..8 id arrays");
  TGM_REQUIRE(num_nodes > 0 && parts && sizes && skip_padded && bitmap && prefix && tmp &&
                  out_unique && out_count, "tgm_dedup_unique: NULL argument");
  Parts ps;
  int64_t total = 0;
  for (int q = 0; q < kMaxParts; ++q) {
    const bool live = q < n_parts;
    TGM_REQUIRE(!live || sizes[q] >= 0, "tgm_dedup_unique: negative size");
    TGM_REQUIRE(!live || sizes[q] == 0 || parts[q], "tgm_dedup_unique: NULL id array");
    total += live ? sizes[q] : 0;
    ps.p[q] = live ? parts[q] : nullptr;
    ps.end[q] = total;
    ps.skip_padded[q] = live ? skip_padded[q] : 1;
  }
  ps.n = n_parts;
  const int64_t words = dedup_words(num_nodes);
  TGM_REQUIRE(tmp_bytes >= 16, "tgm_dedup_unique: tmp too small (see tgm_dedup_sizes)");
  cudaStream_t st = as_stream(stream);
  int32_t *flags = static_cast<int32_t *>(tmp);
  void *cub_tmp = static_cast<char *>(tmp) + 16;
  size_t cub_bytes = size_t(tmp_bytes - 16);
  TGM_CUDA(cudaMemsetAsync(bitmap, 0, size_t(words + 1) * sizeof(uint32_t), st));
  TGM_CUDA(cudaMemsetAsync(flags, 0, 16, st));
  if (total > 0) {
    dedup_mark_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(ps, total, num_nodes, bitmap, flags);
    TGM_LAUNCH_CHECK();
  }
  cub::TransformInputIterator<int32_t, PopcOp, const uint32_t *> it(bitmap, PopcOp());
  TGM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, it, prefix, int(words + 1), st));
  dedup_emit_kernel<<<grid_for(words, 256, 8), 256, 0, st>>>(bitmap, prefix, words, flags,
                                                             out_unique, out_count);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_dedup_map(const uint32_t *bitmap, const int32_t *prefix, const void *tmp,
                             int32_t num_nodes, const int32_t *ids, int64_t n, int32_t *out_local,
                             tgm_stream stream) {
  TGM_REQUIRE(n >= 0 && num_nodes > 0, "tgm_dedup_map: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(bitmap && prefix && tmp && ids && out_local, "tgm_dedup_map: NULL argument");
  dedup_map_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      bitmap, prefix, static_cast<const int32_t *>(tmp), num_nodes, dedup_words(num_nodes), ids, n,
      out_local);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
