// TGAT's hop recursion (reference tgm-team/tgm @ 5183dc9: tgm/nn/encoder/tgat.py:122-149) as ONE
// native call for inference.
//
// The reference walks z[j][i] = merge_j(attn_j(z[j-1][i], z[j-1][i+1], hop i's edges / times), z[0][i])
// hop by hop: L(L+1)/2 attention calls and as many merge calls, each a handful of launches and
// Python round trips.  Here the rows of all hops live back to back -- hop i+1's nodes ARE hop i's
// neighbour slots, so rows [off[i], off[i+1]) of one buffer serve as hop i's seeds and, shifted by
// off[1], as the neighbours of the hops before -- and layer j is one folded attention chain
// (attn_fold.cu) plus one merge layer over the off[L-j+1] rows of hops 0..L-j:
//   gather (all hops' node rows, one launch)
//   per layer: [x-side qk product] -> attn_warp_kernel -> output product -> LayerNorm writing
//              straight into the merge layer's concatenated input -> fc1+ReLU -> fc2
// 12 launches for two layers (45 through the per-hop Python path at the start of round 2), no
// concatenations: the per-hop id / time / edge-feature arrays are read where the sampler wrote them.
#include <algorithm>
#include <new>

#include "common.cuh"

using namespace tgm;

#include "attention.cuh"

struct tgm_tgat {
  int device = -1;
  int L = 0;
  tgm_attn *attn[4] = {};
  tgm_mlp2 *merge[4] = {};
  int node_dim = 0, width = 0;  // widest merge output
  int64_t cap = 0;              // rows
  float *z0 = nullptr, *buf[2] = {nullptr, nullptr};
  ~tgm_tgat() {
    if (device >= 0) {
      DeviceGuard g(device);
      cudaFree(z0), cudaFree(buf[0]), cudaFree(buf[1]);
    }
  }
};

namespace {

struct IdSegs {
  const int32_t *p[5];
  int64_t end[5];
  int n;
};

// out[r, :] = table[id < 0 ? id + N : id] for the ids of all hops (torch negative indexing,
// tgat.py:131-134); out-of-range ids give zeros
__global__ void gather_hops_kernel(const float *__restrict__ table, int64_t N, int dim, IdSegs ids,
                                   int64_t rows, float *__restrict__ out) {
  const int64_t total = rows * dim;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int c = int(i - r * dim);
    int seg = 0;
    int64_t first = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (s + 1 < ids.n && r >= ids.end[s]) seg = s + 1, first = ids.end[s];
    int64_t v = ids.p[seg][r - first];
    if (v < 0) v += N;
    out[i] = (v >= 0 && v < N) ? __ldg(table + v * dim + c) : 0.f;
  }
}

}  // namespace

extern "C" int tgm_tgat_create(tgm_tgat **out, int32_t num_layers, tgm_attn *const *attn,
                               tgm_mlp2 *const *merge, int device) {
  TGM_REQUIRE(out != nullptr, "tgm_tgat_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(num_layers >= 1 && num_layers <= 4, "tgm_tgat_create: 1..4 layers");
  TGM_REQUIRE(attn && merge, "tgm_tgat_create: NULL handle array");
  TGM_REQUIRE(device >= 0, "tgm_tgat_create: a CUDA device is required (no CPU fallback)");
  tgm_tgat *t = new (std::nothrow) tgm_tgat();
  if (!t) return fail(TGM_ERR_OOM, "tgm_tgat_create: host allocation failed");
  t->device = device, t->L = num_layers;
  for (int j = 0; j < num_layers; ++j) {
    t->attn[j] = attn[j], t->merge[j] = merge[j];
    if (!attn[j] || !merge[j] || attn[j]->device != device || merge[j]->device != device) {
      delete t;
      return fail(TGM_ERR_INVALID, "tgm_tgat_create: layer handles must exist on `device`");
    }
    t->width = std::max(t->width, merge[j]->out);
  }
  // the shapes the recursion needs (tgat.py:52-66): layer j reads what layer j-1's merge wrote,
  // every merge concatenates the attention output with the raw node features
  t->node_dim = attn[0]->node_dim;
  for (int j = 0; j < num_layers; ++j) {
    const bool ok = merge[j]->in1 == attn[j]->out_dim && merge[j]->in2 == t->node_dim &&
                    (j == 0 || attn[j]->node_dim == merge[j - 1]->out);
    if (!ok) {
      delete t;
      return fail(TGM_ERR_INVALID, "tgm_tgat_create: layer dimensions do not chain");
    }
  }
  *out = t;
  return TGM_OK;
}

extern "C" void tgm_tgat_destroy(tgm_tgat *t) { delete t; }

extern "C" int tgm_tgat_forward(tgm_tgat *t, const float *node_x, int64_t num_nodes,
                                const int32_t *seed_ids, int64_t S0,
                                const int32_t *const *nbr_ids, const int64_t *const *seed_t,
                                const int64_t *const *nbr_t, const float *const *edge_feat,
                                const float *edge_table, const int32_t *const *edge_rows,
                                int32_t k, float *out, tgm_stream stream) {
  TGM_REQUIRE(t != nullptr, "tgm_tgat_forward: handle is NULL");
  TGM_REQUIRE(S0 >= 0 && k >= 1 && num_nodes >= 0, "tgm_tgat_forward: bad sizes");
  if (S0 == 0) return TGM_OK;
  TGM_REQUIRE(node_x && seed_ids && nbr_ids && seed_t && nbr_t && out,
              "tgm_tgat_forward: NULL array argument");
  TGM_REQUIRE((edge_feat != nullptr) != (edge_table != nullptr && edge_rows != nullptr),
              "tgm_tgat_forward: give either dense edge-feature blocks or (edge_table, edge_rows)");
  const int L = t->L;
  for (int j = 0; j < L; ++j)
    TGM_REQUIRE(attn_folded_covers(t->attn[j], k),
                "tgm_tgat_forward: shape outside the folded chain (see tgm_attn_folded_covers)");
  // rows of hop i: S0 k^i
  int64_t off[6] = {0, S0};
  for (int i = 1; i <= L; ++i) {
    const int64_t hop = (off[i] - off[i - 1]) * k;
    TGM_REQUIRE(off[i] + hop < (int64_t(1) << 31), "tgm_tgat_forward: too many rows");
    off[i + 1] = off[i] + hop;
  }
  for (int i = 0; i < L; ++i) {
    TGM_REQUIRE(nbr_ids[i] && seed_t[i] && nbr_t[i], "tgm_tgat_forward: NULL hop array");
    TGM_REQUIRE(edge_feat ? edge_feat[i] != nullptr : edge_rows[i] != nullptr,
                "tgm_tgat_forward: NULL edge-feature array");
  }
  DeviceGuard g(t->device);
  cudaStream_t st = as_stream(stream);
  const int64_t rows_all = off[L + 1];
  if (rows_all > t->cap) {
    TGM_CUDA(cudaStreamSynchronize(st));
    cudaFree(t->z0), cudaFree(t->buf[0]), cudaFree(t->buf[1]);
    t->z0 = t->buf[0] = t->buf[1] = nullptr;
    t->cap = 0;
    const size_t rows = size_t(rows_all + rows_all / 4);
    TGM_CUDA(cudaMalloc(&t->z0, rows * t->node_dim * 4));
    TGM_CUDA(cudaMalloc(&t->buf[0], rows * t->width * 4));
    TGM_CUDA(cudaMalloc(&t->buf[1], rows * t->width * 4));
    t->cap = int64_t(rows);
  }
  {  // z[0][i] for every hop: rows [off[i], off[i+1])
    IdSegs ids{};
    ids.n = L + 1;
    ids.p[0] = seed_ids, ids.end[0] = off[1];
    for (int i = 1; i <= L; ++i) ids.p[i] = nbr_ids[i - 1], ids.end[i] = off[i + 1];
    for (int i = L + 1; i < 5; ++i) ids.p[i] = nullptr, ids.end[i] = rows_all;
    gather_hops_kernel<<<grid_for(rows_all * t->node_dim, 256, 8), 256, 0, st>>>(
        node_x, num_nodes, t->node_dim, ids, rows_all, t->z0);
    TGM_LAUNCH_CHECK();
  }
  const float *prev = t->z0;  // rows of hops 0 .. L-j+1 after j-1 layers
  for (int j = 1; j <= L; ++j) {
    tgm_attn *a = t->attn[j - 1];
    tgm_mlp2 *m = t->merge[j - 1];
    const int n_hops = L - j + 1;
    const int64_t rows = off[n_hops];
    HopSegs hops{};
    hops.n = n_hops;
    hops.table = edge_table;
    for (int i = 0; i < 4; ++i) {
      if (i < n_hops) {
        hops.nid[i] = nbr_ids[i], hops.nt[i] = nbr_t[i], hops.st[i] = seed_t[i];
        if (edge_table) hops.er[i] = edge_rows[i];
        else hops.ef[i] = edge_feat[i];
      }
      hops.end[i] = off[std::min(i + 1, n_hops)];
    }
    if (int rc = attn_workspace(a, rows, st)) return rc;
    if (int rc = mlp2_workspace(m, rows, st)) return rc;
    // attention -> LayerNorm writes [out | z0 | 0] rows of the merge layer's input
    const int width_in = a->node_dim;
    if (int rc = attn_forward_folded(a, prev, prev + off[1] * width_in, hops, rows, k,
                                     LnTarget{m->cat, m->inp, t->z0, t->node_dim}, st))
      return rc;
    float *dst = j == L ? out : t->buf[j & 1];
    if (int rc = mlp2_forward_cat(m, rows, dst, st)) return rc;
    prev = dst;
  }
  return TGM_OK;
}
