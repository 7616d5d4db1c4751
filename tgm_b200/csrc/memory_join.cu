// TGN node-memory join at time-shard boundaries (BASELINE config 4): the one exchange of the path.
//
// The reference keeps `memory [N, M] f32` and `last_update [N] i64` in one process
// (tgm/nn/encoder/tgn.py:95-110, written by _update_memory :192-216).  Time-sharded over GPUs,
// every rank advances its shard from a common snapshot and touches a subset of the rows; at the
// join each rank needs the rows the others touched.  Instead of dense all-reduces over all N rows
// the touched rows travel as packed records
//     { int32 id | int32 0 | int64 last_update | float memory[M] }          16 + 4 M bytes
// through ONE all-gather; these two kernels are its ends: pack gathers the touched rows of this
// rank into the send block, scatter applies a received block (ranks in ascending order, so the
// later shard wins a row two shards touched).  Both are HBM-bound row copies: a warp per row,
// 128-bit accesses.
#include "common.cuh"

using namespace tgm;

namespace {

__global__ void __launch_bounds__(256)
join_pack_kernel(const float *__restrict__ memory, const int64_t *__restrict__ last_update, int M,
                 const int32_t *__restrict__ ids, int64_t n, int64_t cap,
                 unsigned char *__restrict__ rows) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t rb = 16 + 4 * int64_t(M);
  const int M4 = M >> 2;
  for (int64_t i = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); i < cap;
       i += int64_t(gridDim.x) * wpb) {
    unsigned char *row = rows + i * rb;
    const int32_t v = i < n ? ids[i] : -1;  // rows beyond n pad the block up to the common size
    if (lane == 0) {
      *reinterpret_cast<int2 *>(row) = make_int2(v, 0);
      *reinterpret_cast<int64_t *>(row + 8) = v >= 0 ? last_update[v] : 0;
    }
    if (v < 0) continue;
    const float4 *src = reinterpret_cast<const float4 *>(memory + int64_t(v) * M);
    float4 *dst = reinterpret_cast<float4 *>(row + 16);
    for (int c = lane; c < M4; c += 32) dst[c] = src[c];
  }
}

__global__ void __launch_bounds__(256)
join_scatter_kernel(const unsigned char *__restrict__ rows, int64_t n, int M, int32_t N,
                    float *__restrict__ memory, int64_t *__restrict__ last_update) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t rb = 16 + 4 * int64_t(M);
  const int M4 = M >> 2;
  for (int64_t i = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); i < n;
       i += int64_t(gridDim.x) * wpb) {
    const unsigned char *row = rows + i * rb;
    const int32_t v = *reinterpret_cast<const int32_t *>(row);
    if (v < 0 || v >= N) continue;
    if (lane == 0) last_update[v] = *reinterpret_cast<const int64_t *>(row + 8);
    const float4 *src = reinterpret_cast<const float4 *>(row + 16);
    float4 *dst = reinterpret_cast<float4 *>(memory + int64_t(v) * M);
    for (int c = lane; c < M4; c += 32) dst[c] = ldg_stream_f4(src + c);
  }
}

}  // namespace

extern "C" int64_t tgm_join_row_bytes(int32_t M) { return 16 + 4 * int64_t(M); }

extern "C" int tgm_join_pack(const float *memory, const int64_t *last_update, int32_t M,
                             const int32_t *ids, int64_t n, int64_t cap, void *rows,
                             tgm_stream stream) {
  TGM_REQUIRE(M > 0 && M % 4 == 0, "tgm_join_pack: memory_dim must be a positive multiple of 4");
  TGM_REQUIRE(0 <= n && n <= cap, "tgm_join_pack: need 0 <= n <= cap");
  if (cap == 0) return TGM_OK;
  TGM_REQUIRE(memory && last_update && rows && (ids || n == 0), "tgm_join_pack: NULL argument");
  TGM_REQUIRE(aligned16(memory) && aligned16(rows), "tgm_join_pack: arrays must be 16-byte aligned");
  join_pack_kernel<<<grid_for(cap, 8, 8), 256, 0, as_stream(stream)>>>(
      memory, last_update, M, ids, n, cap, static_cast<unsigned char *>(rows));
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_join_scatter(const void *rows, int64_t n, int32_t M, int32_t num_nodes,
                                float *memory, int64_t *last_update, tgm_stream stream) {
  TGM_REQUIRE(M > 0 && M % 4 == 0, "tgm_join_scatter: memory_dim must be a positive multiple of 4");
  TGM_REQUIRE(n >= 0, "tgm_join_scatter: n must be >= 0");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(rows && memory && last_update, "tgm_join_scatter: NULL argument");
  TGM_REQUIRE(aligned16(memory) && aligned16(rows),
              "tgm_join_scatter: arrays must be 16-byte aligned");
  join_scatter_kernel<<<grid_for(n, 8, 8), 256, 0, as_stream(stream)>>>(
      static_cast<const unsigned char *>(rows), n, M, num_nodes, memory, last_update);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
