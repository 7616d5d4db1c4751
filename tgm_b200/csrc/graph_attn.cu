// TGN embedding: GraphAttentionEmbedding = Time2Vec on (last_update[source] - t) + TransformerConv
// over the sampled-neighbour edge list of one batch.
//
// Replaces (reference tgm-team/tgm @ 5183dc9): tgm/nn/encoder/tgn.py:14-40
// (GraphAttentionEmbedding.__init__/forward) as it is called from
// examples/linkproppred/tgn.py:74-98.  The convolution itself is third-party code,
// torch_geometric.nn.TransformerConv (torch-geometric 2.6.1 per uv.lock:1925-1926, not vendored
// and not installed where this was built; the reference's own test, test/unit/test_nn/test_tgn.py:
// 12-93, is shape/NaN-only), restated from its published algorithm with the arguments tgn.py:25-27
// passes (heads=2, concat, root_weight, bias, no beta; eval mode: attention dropout = identity):
//     q_i = Wq x_i + bq,  k_j = Wk x_j + bk,  v_j = Wv x_j + bv,  e_ij = We [Time2Vec(rel_t) | msg]
//     alpha_ij = softmax over the edges entering i of  q_i . (k_j + e_ij) / sqrt(C)    per head
//     out_i = concat_h sum_j alpha_ij (v_j + e_ij)  +  Wskip x_i + bskip
// with j = edge_index[0] (source), i = edge_index[1] (target); PyG's softmax divides by
// (sum + 1e-16).  PARITY UNPINNED (oracle/tgn_oracle.py::transformer_conv says the same).
//
// Device plan (a batch is ~6,000 edges over ~5,000 nodes: launch-bound, so few launches):
//   1. one SGEMM  P[n, 4HC] = x [Wq;Wk;Wv;Wskip]^T                      (cuBLAS, true fp32)
//   2. gae_edge_prep: edge attributes A[m, TD+D] (Time2Vec in registers) + sort keys
//   3. one SGEMM  Ep[m, HC] = A We^T
//   4. stable radix sort of the edges by target (CUB) + row pointers by binary search
//   5. gae_node_kernel: one warp per target node walks its incoming edges IN EDGE ORDER
//      (deterministic, the order a CPU index_add uses): logits -> max -> exp/sum -> weighted sum
#include <cublas_v2.h>

#include <cub/device/device_radix_sort.cuh>
#include <new>

#include "bwd_common.cuh"
#include "common.cuh"

using namespace tgm;

struct tgm_gae {
  int device = -1;
  int32_t in = 0, HC = 0, H = 0, C = 0, D = 0, TD = 0, A = 0;
  float *Wall = nullptr, *ball = nullptr, *We = nullptr, *tw = nullptr, *tb = nullptr;
  cublasHandle_t blas = nullptr;
  int64_t cap_n = 0, cap_m = 0;
  float *P = nullptr, *attr = nullptr, *Ep = nullptr, *logit = nullptr;
  int32_t *key_in = nullptr, *key_out = nullptr, *perm_in = nullptr, *perm_out = nullptr;
  int64_t *rowptr = nullptr;
  void *cub_tmp = nullptr;
  size_t cub_bytes = 0;
  // backward workspaces (tgm_gae_backward)
  int64_t bcap_n = 0, bcap_m = 0;
  float *dP = nullptr, *out_fwd = nullptr;                                // [n,4HC], [n,HC]
  float *dE = nullptr, *dA = nullptr, *dAttr = nullptr, *rel = nullptr;   // [m,HC], [m,H], [m,TD], [m]
  ~tgm_gae() {
    if (device >= 0) {
      DeviceGuard g(device);
      for (float *p : {dP, out_fwd, dE, dA, dAttr, rel}) cudaFree(p);
      for (float *p : {Wall, ball, We, tw, tb, P, attr, Ep, logit}) cudaFree(p);
      for (int32_t *p : {key_in, key_out, perm_in, perm_out}) cudaFree(p);
      cudaFree(rowptr), cudaFree(cub_tmp);
      if (blas) cublasDestroy(blas);
    }
  }
};

namespace {

int gae_blas_fail(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define GAE_BLAS(expr)                                                \
  do {                                                                \
    cublasStatus_t _s = (expr);                                       \
    if (_s != CUBLAS_STATUS_SUCCESS) return gae_blas_fail(_s, #expr); \
  } while (0)

int gae_copy(float **dst, const float *src, size_t n, size_t offset_elems = 0, float *into = nullptr) {
  if (into == nullptr) {
    cudaError_t e = cudaMalloc(dst, (n ? n : 1) * sizeof(float));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(gae parameter)", __FILE__, __LINE__);
    into = *dst;
  }
  if (n) TGM_CUDA(cudaMemcpy(into + offset_elems, src, n * sizeof(float), cudaMemcpyDefault));
  return TGM_OK;
}

// A[e, :TD] = cos(fma(float(last_update[src[e]] - t[e]), w, b))  (tgn.py:37-38 + Time2Vec),
// A[e, TD:] = msg[e, :]  (tgn.py:39);  key[e] = target of e, perm[e] = e.
__global__ void __launch_bounds__(256)
gae_edge_prep_kernel(const int64_t *__restrict__ esrc, const int64_t *__restrict__ edst,
                     const int64_t *__restrict__ t, const int64_t *__restrict__ last_update,
                     const float *__restrict__ msg, const float *__restrict__ tw,
                     const float *__restrict__ tb, int64_t m, int TD, int D,
                     float *__restrict__ attr, int32_t *__restrict__ key,
                     int32_t *__restrict__ perm) {
  const int A = TD + D;
  const int64_t total = m * A;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = i / A;
    const int c = int(i - e * A);
    if (c < TD) {
      const float rel = float(__ldg(last_update + __ldg(esrc + e)) - __ldg(t + e));
      attr[i] = cosf(__fmaf_rn(rel, __ldg(tw + c), __ldg(tb + c)));
    } else {
      attr[i] = __ldg(msg + e * D + (c - TD));
    }
    if (c == 0) {
      key[e] = int32_t(__ldg(edst + e));
      perm[e] = int32_t(e);
    }
  }
}

// rowptr[i] = first sorted position whose key is >= i   (i in [0, n])
__global__ void __launch_bounds__(256)
gae_rowptr_kernel(const int32_t *__restrict__ sorted_key, int64_t m, int64_t n,
                  int64_t *__restrict__ rowptr) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i <= n;
       i += int64_t(gridDim.x) * blockDim.x) {
    int64_t lo = 0, hi = m;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (int64_t(__ldg(sorted_key + mid)) < i) lo = mid + 1; else hi = mid;
    }
    rowptr[i] = lo;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per target node.  P rows are [q | k | v | skip] (each HC wide, biases in ball).
__global__ void __launch_bounds__(256)
gae_node_kernel(const float *__restrict__ P, const float *__restrict__ ball,
                const float *__restrict__ Ep, const int64_t *__restrict__ esrc,
                const int32_t *__restrict__ perm, const int64_t *__restrict__ rowptr, int64_t n,
                int H, int C, float *__restrict__ logit, float *__restrict__ out) {
  const int HC = H * C, ld = 4 * HC;
  const int lane = lane_id();
  const float scale = sqrtf(float(C));
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < n;
       i += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int64_t lo = rowptr[i], hi = rowptr[i + 1];
    const float *Pi = P + i * ld;
    for (int h = 0; h < H; ++h) {
      const int c0 = h * C;
      // pass 1: logits and their maximum
      float mx = -INFINITY;
      for (int64_t p = lo; p < hi; ++p) {
        const int32_t e = perm[p];
        const float *Pj = P + esrc[e] * ld + HC;  // k_j
        const float *Ee = Ep + int64_t(e) * HC;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) {
          const float q = Pi[c0 + c] + ball[c0 + c];
          const float kk = Pj[c0 + c] + ball[HC + c0 + c] + Ee[c0 + c];
          acc = __fmaf_rn(q, kk, acc);
        }
        acc = warp_sum(acc) / scale;
        if (lane == 0) logit[int64_t(e) * H + h] = acc;
        mx = fmaxf(mx, acc);
      }
      __syncwarp();
      // pass 2: sum of exponentials (edge order)
      float sum = 0.f;
      for (int64_t p = lo; p < hi; ++p) sum += expf(logit[int64_t(perm[p]) * H + h] - mx);
      const float den = sum + 1e-16f;
      // pass 3: weighted sum of (v_j + e_ij), edge order; lanes own channels
      for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int64_t p = lo; p < hi; ++p) {
          const int32_t e = perm[p];
          const float a = expf(logit[int64_t(e) * H + h] - mx) / den;
          const float v = P[esrc[e] * ld + 2 * HC + c0 + c] + ball[2 * HC + c0 + c] +
                          Ep[int64_t(e) * HC + c0 + c];
          acc += v * a;
        }
        out[i * HC + c0 + c] = acc + (Pi[3 * HC + c0 + c] + ball[3 * HC + c0 + c]);
      }
      __syncwarp();
    }
  }
}

int gae_reserve(tgm_gae *g, int64_t n, int64_t m) {
  if (n > g->cap_n) {
    cudaFree(g->P), cudaFree(g->rowptr);
    g->P = nullptr, g->rowptr = nullptr, g->cap_n = 0;
    const int64_t cap = n + n / 4 + 64;
    TGM_CUDA(cudaMalloc(&g->P, size_t(cap) * 4 * g->HC * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->rowptr, size_t(cap + 1) * sizeof(int64_t)));
    g->cap_n = cap;
  }
  if (m > g->cap_m) {
    for (float **p : {&g->attr, &g->Ep, &g->logit}) cudaFree(*p), *p = nullptr;
    for (int32_t **p : {&g->key_in, &g->key_out, &g->perm_in, &g->perm_out}) cudaFree(*p), *p = nullptr;
    cudaFree(g->cub_tmp), g->cub_tmp = nullptr, g->cap_m = 0;
    const int64_t cap = m + m / 4 + 64;
    TGM_CUDA(cudaMalloc(&g->attr, size_t(cap) * g->A * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->Ep, size_t(cap) * g->HC * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->logit, size_t(cap) * g->H * sizeof(float)));
    for (int32_t **p : {&g->key_in, &g->key_out, &g->perm_in, &g->perm_out})
      TGM_CUDA(cudaMalloc(p, size_t(cap) * sizeof(int32_t)));
    size_t bytes = 0;
    TGM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, g->key_in, g->key_out, g->perm_in,
                                             g->perm_out, int(cap)));
    TGM_CUDA(cudaMalloc(&g->cub_tmp, bytes ? bytes : 1));
    g->cub_bytes = bytes;
    g->cap_m = cap;
  }
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_gae_create(tgm_gae **out, int32_t in_channels, int32_t out_channels,
                              int32_t heads, int32_t msg_dim, int32_t time_dim, const float *W_query,
                              const float *b_query, const float *W_key, const float *b_key,
                              const float *W_value, const float *b_value, const float *W_edge,
                              const float *W_skip, const float *b_skip, const float *t2v_w,
                              const float *t2v_b, int device) {
  TGM_REQUIRE(out != nullptr, "tgm_gae_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(in_channels > 0 && out_channels > 0 && heads > 0 && msg_dim >= 0 && time_dim > 0,
              "tgm_gae_create: bad dimensions");
  TGM_REQUIRE(out_channels % heads == 0,
              "tgm_gae_create: out_channels (= heads * per-head channels) must divide by heads");
  TGM_REQUIRE(W_query && b_query && W_key && b_key && W_value && b_value && W_edge && W_skip &&
                  b_skip && t2v_w && t2v_b, "tgm_gae_create: NULL parameter");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(TGM_ERR_NO_DEVICE, "tgm_gae_create: no such CUDA device");
  }
  DeviceGuard guard(device);
  tgm_gae *g = new (std::nothrow) tgm_gae();
  if (!g) return fail(TGM_ERR_OOM, "tgm_gae_create: out of host memory");
  g->device = device;
  g->in = in_channels, g->HC = out_channels, g->H = heads, g->C = out_channels / heads;
  g->D = msg_dim, g->TD = time_dim, g->A = msg_dim + time_dim;
  const size_t HC = size_t(g->HC), in = size_t(g->in);
  int rc = TGM_OK;
  cudaError_t e = cudaMalloc(&g->Wall, 4 * HC * in * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&g->ball, 4 * HC * sizeof(float));
  if (e != cudaSuccess) rc = cuda_fail(e, "gae parameter allocation", __FILE__, __LINE__);
  const float *Ws[4] = {W_query, W_key, W_value, W_skip};
  const float *bs[4] = {b_query, b_key, b_value, b_skip};
  for (int q = 0; q < 4 && !rc; ++q) {
    rc = gae_copy(nullptr, Ws[q], HC * in, q * HC * in, g->Wall);
    if (!rc) rc = gae_copy(nullptr, bs[q], HC, q * HC, g->ball);
  }
  if (!rc) rc = gae_copy(&g->We, W_edge, HC * size_t(g->A));
  if (!rc) rc = gae_copy(&g->tw, t2v_w, size_t(time_dim));
  if (!rc) rc = gae_copy(&g->tb, t2v_b, size_t(time_dim));
  if (!rc) {
    cublasStatus_t s = cublasCreate(&g->blas);
    if (s != CUBLAS_STATUS_SUCCESS) rc = gae_blas_fail(s, "cublasCreate");
    else cublasSetMathMode(g->blas, CUBLAS_PEDANTIC_MATH);
  }
  if (rc) {
    delete g;
    return rc;
  }
  *out = g;
  return TGM_OK;
}

extern "C" void tgm_gae_destroy(tgm_gae *g) { delete g; }

extern "C" int tgm_gae_forward(tgm_gae *g, const float *x, const int64_t *last_update, int64_t n,
                               const int64_t *edge_src, const int64_t *edge_dst, const int64_t *t,
                               const float *msg, int64_t m, float *out, tgm_stream stream) {
  TGM_REQUIRE(g != nullptr, "tgm_gae_forward: handle is NULL");
  TGM_REQUIRE(n >= 0 && m >= 0 && n < (int64_t(1) << 31) && m < (int64_t(1) << 31),
              "tgm_gae_forward: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(x && last_update && out, "tgm_gae_forward: NULL node argument");
  TGM_REQUIRE(m == 0 || (edge_src && edge_dst && t && (msg || g->D == 0)),
              "tgm_gae_forward: NULL edge argument");
  DeviceGuard guard(g->device);
  cudaStream_t st = as_stream(stream);
  int rc = gae_reserve(g, n, m);
  if (rc) return rc;
  const int HC = g->HC;
  const float one = 1.f, zero = 0.f;
  GAE_BLAS(cublasSetStream(g->blas, st));
  // P[n, 4HC] = x[n, in] Wall[4HC, in]^T   (row-major)
  GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_T, CUBLAS_OP_N, 4 * HC, int(n), g->in, &one, g->Wall,
                       g->in, x, g->in, &zero, g->P, 4 * HC));
  if (m > 0) {
    gae_edge_prep_kernel<<<grid_for(m * g->A, 256, 8), 256, 0, st>>>(
        edge_src, edge_dst, t, last_update, msg, g->tw, g->tb, m, g->TD, g->D, g->attr, g->key_in,
        g->perm_in);
    TGM_LAUNCH_CHECK();
    // Ep[m, HC] = attr[m, A] We[HC, A]^T
    GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_T, CUBLAS_OP_N, HC, int(m), g->A, &one, g->We, g->A,
                         g->attr, g->A, &zero, g->Ep, HC));
    int end_bit = 1;
    while (end_bit < 31 && (int64_t(1) << end_bit) < n) ++end_bit;
    size_t bytes = g->cub_bytes;
    TGM_CUDA(cub::DeviceRadixSort::SortPairs(g->cub_tmp, bytes, g->key_in, g->key_out, g->perm_in,
                                             g->perm_out, int(m), 0, end_bit, st));
  }
  gae_rowptr_kernel<<<grid_for(n + 1, 256, 8), 256, 0, st>>>(g->key_out, m, n, g->rowptr);
  TGM_LAUNCH_CHECK();
  gae_node_kernel<<<grid_for(n, 8, 8), 256, 0, st>>>(g->P, g->ball, g->Ep, edge_src, g->perm_out,
                                                     g->rowptr, n, g->H, g->C, g->logit, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

// ---- training: parameter refresh and backward ----------------------------------------------------
// Chain rule of the forward above (oracle/tgn_oracle.py::transformer_conv_backward states it in
// float64).  With s_ij = q_i.(k_j + e_ij)/sqrt(C) and a = softmax over the edges entering i:
//     d a_ij = dOut_i . (v_j + e_ij)             d s_ij = a_ij (d a_ij - sum_j' a_ij' d a_ij') / sqrt(C)
//     d v_j += a_ij dOut_i     d k_j += d s_ij q_i     d q_i += d s_ij (k_j + e_ij)
//     d e_ij = a_ij dOut_i + d s_ij q_i          d skip_i = dOut_i
// (PyG's +1e-16 in the softmax denominator leaves the Jacobian a (delta - a) unchanged.)
// The forward is recomputed (nothing is saved but the inputs); one warp per target node walks its
// incoming edges again: d q / d skip rows are owned by the warp, d k / d v rows of the SOURCE nodes
// are accumulated with atomics; the linears' gradients are cuBLAS GEMMs over dP / dE.
namespace {

__global__ void __launch_bounds__(256)
gae_node_bwd_kernel(const float *__restrict__ P, const float *__restrict__ ball,
                    const float *__restrict__ Ep, const int64_t *__restrict__ esrc,
                    const int32_t *__restrict__ perm, const int64_t *__restrict__ rowptr, int64_t n,
                    int H, int C, const float *__restrict__ logit, const float *__restrict__ d_out,
                    float *__restrict__ dA, float *__restrict__ dP, float *__restrict__ dE) {
  const int HC = H * C, ld = 4 * HC;
  const int lane = lane_id();
  const float scale = sqrtf(float(C));
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < n;
       i += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int64_t lo = rowptr[i], hi = rowptr[i + 1];
    const float *Pi = P + i * ld;
    const float *go = d_out + i * HC;
    for (int h = 0; h < H; ++h) {
      const int c0 = h * C;
      // softmax statistics, as the forward computes them (edge order)
      float mx = -INFINITY;
      for (int64_t p = lo; p < hi; ++p) mx = fmaxf(mx, logit[int64_t(perm[p]) * H + h]);
      float sum = 0.f;
      for (int64_t p = lo; p < hi; ++p) sum += expf(logit[int64_t(perm[p]) * H + h] - mx);
      const float den = sum + 1e-16f;
      // pass A: d a_ij for every incoming edge and dot = sum_j a_ij d a_ij
      float dot = 0.f;
      for (int64_t p = lo; p < hi; ++p) {
        const int32_t e = perm[p];
        const float *Pj = P + esrc[e] * ld + 2 * HC;  // v_j
        const float *Ee = Ep + int64_t(e) * HC;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32)
          acc = __fmaf_rn(go[c0 + c], Pj[c0 + c] + ball[2 * HC + c0 + c] + Ee[c0 + c], acc);
        acc = warp_sum(acc);
        if (lane == 0) dA[int64_t(e) * H + h] = acc;
        dot = __fmaf_rn(expf(logit[int64_t(e) * H + h] - mx) / den, acc, dot);
      }
      __syncwarp();
      // pass B: lanes own channels
      for (int c = lane; c < C; c += 32) {
        const float q = Pi[c0 + c] + ball[c0 + c];
        const float g = go[c0 + c];
        float dq = 0.f;
        for (int64_t p = lo; p < hi; ++p) {
          const int32_t e = perm[p];
          const int64_t j = esrc[e];
          const float a = expf(logit[int64_t(e) * H + h] - mx) / den;
          const float ds = a * (dA[int64_t(e) * H + h] - dot) / scale;
          const float kk = P[j * ld + HC + c0 + c] + ball[HC + c0 + c] + Ep[int64_t(e) * HC + c0 + c];
          dq = __fmaf_rn(ds, kk, dq);
          dE[int64_t(e) * HC + c0 + c] = a * g + ds * q;
          atomicAdd(dP + j * ld + HC + c0 + c, ds * q);
          atomicAdd(dP + j * ld + 2 * HC + c0 + c, a * g);
        }
        dP[i * ld + c0 + c] = dq;
        dP[i * ld + 3 * HC + c0 + c] = g;
      }
      __syncwarp();
    }
  }
}

// rel[e] = float(last_update[src[e]] - t[e])  (the Time2Vec input of the forward)
__global__ void gae_rel_kernel(const int64_t *__restrict__ esrc, const int64_t *__restrict__ t,
                               const int64_t *__restrict__ last_update, int64_t m,
                               float *__restrict__ rel) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < m;
       e += int64_t(gridDim.x) * blockDim.x)
    rel[e] = float(last_update[esrc[e]] - t[e]);
}

int gae_reserve_bwd(tgm_gae *g, int64_t n, int64_t m) {
  if (n > g->bcap_n) {
    cudaFree(g->dP), cudaFree(g->out_fwd);
    g->dP = g->out_fwd = nullptr, g->bcap_n = 0;
    const int64_t cap = n + n / 4 + 64;
    TGM_CUDA(cudaMalloc(&g->dP, size_t(cap) * 4 * g->HC * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->out_fwd, size_t(cap) * g->HC * sizeof(float)));
    g->bcap_n = cap;
  }
  if (m > g->bcap_m) {
    for (float **p : {&g->dE, &g->dA, &g->dAttr, &g->rel}) cudaFree(*p), *p = nullptr;
    g->bcap_m = 0;
    const int64_t cap = m + m / 4 + 64;
    TGM_CUDA(cudaMalloc(&g->dE, size_t(cap) * g->HC * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->dA, size_t(cap) * g->H * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->dAttr, size_t(cap) * g->TD * sizeof(float)));
    TGM_CUDA(cudaMalloc(&g->rel, size_t(cap) * sizeof(float)));
    g->bcap_m = cap;
  }
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_gae_set_params(tgm_gae *g, const float *W_query, const float *b_query,
                                  const float *W_key, const float *b_key, const float *W_value,
                                  const float *b_value, const float *W_edge, const float *W_skip,
                                  const float *b_skip, const float *t2v_w, const float *t2v_b,
                                  tgm_stream stream) {
  TGM_REQUIRE(g != nullptr, "tgm_gae_set_params: handle is NULL");
  TGM_REQUIRE(W_query && b_query && W_key && b_key && W_value && b_value && W_edge && W_skip &&
                  b_skip && t2v_w && t2v_b, "tgm_gae_set_params: NULL parameter");
  DeviceGuard guard(g->device);
  cudaStream_t st = as_stream(stream);
  const size_t HC = size_t(g->HC), in = size_t(g->in);
  const float *Ws[4] = {W_query, W_key, W_value, W_skip};
  const float *bs[4] = {b_query, b_key, b_value, b_skip};
  for (int q = 0; q < 4; ++q) {
    TGM_CUDA(cudaMemcpyAsync(g->Wall + q * HC * in, Ws[q], HC * in * 4, cudaMemcpyDefault, st));
    TGM_CUDA(cudaMemcpyAsync(g->ball + q * HC, bs[q], HC * 4, cudaMemcpyDefault, st));
  }
  TGM_CUDA(cudaMemcpyAsync(g->We, W_edge, HC * size_t(g->A) * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(g->tw, t2v_w, size_t(g->TD) * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(g->tb, t2v_b, size_t(g->TD) * 4, cudaMemcpyDefault, st));
  return TGM_OK;
}

extern "C" int tgm_gae_backward(tgm_gae *g, const float *x, const int64_t *last_update, int64_t n,
                                const int64_t *edge_src, const int64_t *edge_dst, const int64_t *t,
                                const float *msg, int64_t m, const float *d_out, float *d_x,
                                float *g_W_qkvs, float *g_b_qkvs, float *g_W_edge, float *g_t2v_w,
                                float *g_t2v_b, tgm_stream stream) {
  TGM_REQUIRE(g != nullptr, "tgm_gae_backward: handle is NULL");
  TGM_REQUIRE(n >= 0 && m >= 0 && n < (int64_t(1) << 31) && m < (int64_t(1) << 31),
              "tgm_gae_backward: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(x && last_update && d_out, "tgm_gae_backward: NULL node argument");
  TGM_REQUIRE(m == 0 || (edge_src && edge_dst && t && (msg || g->D == 0)),
              "tgm_gae_backward: NULL edge argument");
  TGM_REQUIRE(g_W_qkvs && g_b_qkvs && g_W_edge && g_t2v_w && g_t2v_b,
              "tgm_gae_backward: NULL gradient buffer");
  DeviceGuard guard(g->device);
  cudaStream_t st = as_stream(stream);
  int rc = gae_reserve_bwd(g, n, m);
  if (rc) return rc;
  // recompute P, attr, Ep, the target-sorted edge order and the logits
  rc = tgm_gae_forward(g, x, last_update, n, edge_src, edge_dst, t, msg, m, g->out_fwd, stream);
  if (rc) return rc;
  const int HC = g->HC, HC4 = 4 * g->HC, in = g->in, A = g->A, TD = g->TD;
  const float one = 1.f, zero = 0.f;
  TGM_CUDA(cudaMemsetAsync(g->dP, 0, size_t(n) * HC4 * sizeof(float), st));
  gae_node_bwd_kernel<<<grid_for(n, 8, 8), 256, 0, st>>>(g->P, g->ball, g->Ep, edge_src,
                                                         g->perm_out, g->rowptr, n, g->H, g->C,
                                                         g->logit, d_out, g->dA, g->dP, g->dE);
  TGM_LAUNCH_CHECK();
  GAE_BLAS(cublasSetStream(g->blas, st));
  // g_W_qkvs[4HC,in] += dP^T x ; g_b_qkvs[4HC] += colsum(dP) ; d_x[n,in] += dP Wall   (row-major)
  GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_N, CUBLAS_OP_T, in, HC4, int(n), &one, x, in, g->dP, HC4,
                       &one, g_W_qkvs, in));
  colsum_add_kernel<<<colsum_grid(n, HC4), 128, 0, st>>>(g->dP, n, HC4, HC4, g_b_qkvs);
  TGM_LAUNCH_CHECK();
  if (d_x != nullptr)
    GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_N, CUBLAS_OP_N, in, int(n), HC4, &one, g->Wall, in,
                         g->dP, HC4, &one, d_x, in));
  if (m > 0) {
    // g_W_edge[HC,A] += dE^T attr ; d(attr)[:, :TD] = dE W_edge[:, :TD] ; then Time2Vec
    GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_N, CUBLAS_OP_T, A, HC, int(m), &one, g->attr, A, g->dE,
                         HC, &one, g_W_edge, A));
    GAE_BLAS(cublasSgemm(g->blas, CUBLAS_OP_N, CUBLAS_OP_N, TD, int(m), HC, &one, g->We, A, g->dE,
                         HC, &zero, g->dAttr, TD));
    gae_rel_kernel<<<grid_for(m, 256, 8), 256, 0, st>>>(edge_src, t, last_update, m, g->rel);
    TGM_LAUNCH_CHECK();
    t2v_grad_kernel<<<colsum_grid(m, TD), 128, 0, st>>>(g->rel, nullptr, 1, g->dAttr, TD, m, TD,
                                                        g->tw, g->tb, g_t2v_w, g_t2v_b);
    TGM_LAUNCH_CHECK();
  }
  return TGM_OK;
}
