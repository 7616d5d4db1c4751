// DyGFormer forward on the device.
//
// Replaces DyGFormer.forward (reference tgm-team/tgm @ 5183dc9, tgm/nn/encoder/dygformer.py:243-431)
// with NeighborCooccurrenceEncoder (:13-77), TransformerEncoder (:80-143) and _get_patches
// (:433-444).  Eval mode (dropout = identity), fp32 throughout.
//
//   frontend kernel   sequences [self | k sampled neighbours] of the source and destination of
//                     every edge -> the four per-position channels: node features (0 for padding,
//                     :296-299), edge features ([0 | nbr_edge_x], :284-292), Time2Vec(t_edge -
//                     t_nbr) (0 for padding, :301-310) and the co-occurrence encoding: exact
//                     integer counts of each id in its own and in the partner sequence (:33-51)
//                     through Linear(1,C)-ReLU-Linear(C,C), summed over the two counts (:68-76).
//                     Patching (:433-444) is a pure reshape of these [rows, L, d] tensors.
//   dense part        4 channel projections written straight into the [B, 2*NP, 4C] token
//                     layout (strided-batched SGEMM), L pre-LN transformer layers (cuBLAS SGEMM
//                     + LayerNorm / softmax / GELU kernels), mean-pool per side, output Linear.
#include <cublas_v2.h>

#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"

using namespace tgm;

struct DygLayerDev {
  float *in_w, *in_b, *out_w, *out_b, *f1_w, *f1_b, *f2_w, *f2_b, *ln0_w, *ln0_b, *ln1_w, *ln1_b;
};

struct tgm_dyg {
  int device = -1;
  int dN = 0, dE = 0, dT = 0, C = 0, out = 0, P = 0, layers = 0, H = 0, L = 0, NP = 0, E = 0;
  float eps = 1e-5f;
  std::vector<float *> owned;
  float *tw = nullptr, *tb = nullptr, *c_w1 = nullptr, *c_b1 = nullptr, *c_w2 = nullptr,
        *c_b2 = nullptr, *proj_w[4] = {}, *proj_b = nullptr /* [4C] */, *out_w = nullptr,
        *out_b = nullptr;
  std::vector<DygLayerDev> lay;
  cublasHandle_t blas = nullptr;
  int64_t cap = 0;  // pairs
  float *feat[4] = {};  // [2B, L, d_c]
  float *X = nullptr, *Xn = nullptr, *QKV = nullptr, *S = nullptr, *O = nullptr, *F1 = nullptr,
        *tmp = nullptr, *pooled = nullptr;
  ~tgm_dyg() {
    if (device >= 0) {
      DeviceGuard g(device);
      for (float *p : owned) cudaFree(p);
      for (float *p : {feat[0], feat[1], feat[2], feat[3], X, Xn, QKV, S, O, F1, tmp, pooled})
        cudaFree(p);
      if (blas) cublasDestroy(blas);
    }
  }
};

namespace {

int blas_fail3(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define DYG_BLAS(expr)                                             \
  do {                                                             \
    cublasStatus_t _s = (expr);                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) return blas_fail3(_s, #expr); \
  } while (0)

// C[S,N] = A[S,K] . W[N,K]^T (row-major), cuBLAS true-fp32 SGEMM
cublasStatus_t gemm_nt3(cublasHandle_t h, int64_t S, int N, int K, const float *A, const float *W,
                        float *C) {
  const float one = 1.f, zero = 0.f;
  return cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, int(S), K, &one, W, K, A, K, &zero, C, N);
}

// id / time of position l of sequence row r (rows [0,B) = sources, [B,2B) = destinations)
__device__ __forceinline__ int32_t seq_id(const int32_t *src, const int32_t *dst,
                                          const int32_t *nbrs, int64_t B, int k, int64_t r, int l) {
  if (l == 0) return r < B ? src[r] : dst[r - B];
  return nbrs[r * k + (l - 1)];
}

__global__ void __launch_bounds__(256)
dyg_frontend_kernel(const float *__restrict__ node_x, int64_t num_nodes,
                    const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                    const int64_t *__restrict__ edge_time, const int32_t *__restrict__ nbrs,
                    const int64_t *__restrict__ nbr_t, const float *__restrict__ nbr_x, int64_t B,
                    int L, int dN, int dE, int dT, int C, const float *__restrict__ tw,
                    const float *__restrict__ tb, const float *__restrict__ w1,
                    const float *__restrict__ b1, const float *__restrict__ w2,
                    const float *__restrict__ b2, float *__restrict__ f_node,
                    float *__restrict__ f_edge, float *__restrict__ f_time,
                    float *__restrict__ f_cooc) {
  extern __shared__ float s_h[];  // [warps][C] hidden activations of the co-occurrence MLP
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float *hsum = s_h + warp * C;
  const int k = L - 1;
  const int64_t total = 2 * B * L;
  for (int64_t i = int64_t(blockIdx.x) * wpb + warp; i < total; i += int64_t(gridDim.x) * wpb) {
    const int64_t r = i / L;
    const int l = int(i - r * L);
    const int64_t pair = r < B ? r : r - B, other = r < B ? r + B : r - B;
    const int32_t id = seq_id(src, dst, nbrs, B, k, r, l);
    const bool padded = id == TGM_PADDED_NODE_ID;
    // node channel (dygformer.py:296-299)
    int64_t row = id;
    if (row < 0) row += num_nodes;
    for (int c = lane; c < dN; c += 32)
      f_node[i * dN + c] = (padded || row < 0 || row >= num_nodes) ? 0.f : node_x[row * dN + c];
    // edge channel: zero row for the seed itself (:284-292)
    for (int c = lane; c < dE; c += 32)
      f_edge[i * dE + c] = l == 0 ? 0.f : nbr_x[(r * k + (l - 1)) * dE + c];
    // time channel (:301-310)
    const int64_t tq = edge_time[pair];
    const float dt = float(tq - (l == 0 ? tq : nbr_t[r * k + (l - 1)]));
    for (int c = lane; c < dT; c += 32)
      f_time[i * dT + c] = padded ? 0.f : cosf(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c)));
    // co-occurrence counts (:33-51)
    int own = 0, cross = 0;
    if (!padded) {
      for (int base = 0; base < L; base += 32) {
        const int j = base + lane;
        const bool in = j < L;
        const bool a = in && seq_id(src, dst, nbrs, B, k, r, j) == id;
        const bool b = in && seq_id(src, dst, nbrs, B, k, other, j) == id;
        own += __popc(__ballot_sync(0xffffffffu, a));
        cross += __popc(__ballot_sync(0xffffffffu, b));
      }
    }
    const float f0 = float(own), f1 = float(cross);
    for (int j = lane; j < C; j += 32) {
      const float w = __ldg(w1 + j), bb = __ldg(b1 + j);
      hsum[j] = fmaxf(fmaf(w, f0, bb), 0.f) + fmaxf(fmaf(w, f1, bb), 0.f);
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      float acc = 2.f * __ldg(b2 + c);  // MLP(f0) + MLP(f1): the output bias counts twice (:68-76)
      for (int j = 0; j < C; ++j) acc = fmaf(__ldg(w2 + c * C + j), hsum[j], acc);
      f_cooc[i * C + c] = acc;
    }
    __syncwarp();
  }
}

// X[row, c] (+)= ... helpers over a [rows, cols] matrix
__global__ void add_bias_kernel(float *__restrict__ x, const float *__restrict__ b, int64_t rows,
                                int cols, int gelu) {
  const int64_t total = rows * cols;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    float v = x[i] + __ldg(b + int(i % cols));
    if (gelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // exact GELU (F.gelu)
    x[i] = v;
  }
}
// out = r + (y + b)   (residual connection with the bias of the preceding Linear; out may alias r)
__global__ void residual_bias_kernel(float *out, const float *r, const float *__restrict__ y,
                                     const float *__restrict__ b, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x)
    out[i] = r[i] + (y[i] + __ldg(b + int(i % cols)));
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ x, const float *__restrict__ w,
                 const float *__restrict__ b, int64_t rows, int cols, float eps,
                 float *__restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows;
       r += int64_t(gridDim.x) * wpb) {
    const float *xr = x + r * cols;
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += xr[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / float(cols);
    float var = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float d = xr[c] - mean;
      var = fmaf(d, d, var);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / float(cols) + eps);
    for (int c = lane; c < cols; c += 32)
      out[r * cols + c] = (xr[c] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
  }
}

// softmax(scale * row) in place; one warp per row
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float *__restrict__ s, int64_t rows, int cols, float scale) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows;
       r += int64_t(gridDim.x) * wpb) {
    float *row = s + r * cols;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, row[c] * scale);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float e = expf(row[c] * scale - m);
      row[c] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int c = lane; c < cols; c += 32) row[c] *= inv;
  }
}

// pooled[side*B + b, :] = mean_p X[b, side*NP + p, :]   (dygformer.py:419-425)
__global__ void meanpool_kernel(const float *__restrict__ X, int64_t B, int NP, int E,
                                float *__restrict__ pooled) {
  const int64_t total = 2 * B * E;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t row = i / E;
    const int c = int(i - row * E);
    const int64_t side = row / B, b = row - side * B;
    const float *x = X + (b * 2 * NP + side * NP) * E + c;
    float acc = 0.f;
    for (int p = 0; p < NP; ++p) acc += x[int64_t(p) * E];
    pooled[i] = acc / float(NP);
  }
}

int dup(tgm_dyg *m, float **dst, const float *src, size_t n) {
  TGM_REQUIRE(src != nullptr, "tgm_dyg_create: NULL parameter");
  TGM_CUDA(cudaMalloc(dst, (n ? n : 1) * sizeof(float)));
  m->owned.push_back(*dst);
  if (n) TGM_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyDefault));
  return TGM_OK;
}

// out[S,N] = act(A[S,K] W[N,K]^T + b (+ residual)); residual may alias out, `tmp` is [S,N] scratch.
// Token-sized linears run on the tensor cores with the bias / residual / GELU fused into the
// epilogue (gemm_fastf32.cu: fp32-accurate 9xBF16 emulation on tcgen05, 1.6x cuBLAS's SIMT SGEMM at
// a quarter of its rounding error); small or unaligned ones use cuBLAS (true fp32) + one
// elementwise pass.
int linear(cublasHandle_t blas, int64_t S, int N, int K, const float *A, const float *W,
           const float *b, const float *residual, int gelu, float *out, float *tmp,
           cudaStream_t st) {
  if (g_gemm_fastf32 && S >= 2048) {
    const int rc = fastf32_linear(S, N, K, A, W, b, residual, gelu, out, st);
    if (rc != 0) return rc < 0 ? rc : TGM_OK;
  }
  float *dst = residual ? tmp : out;
  DYG_BLAS(gemm_nt3(blas, S, N, K, A, W, dst));
  if (residual)
    residual_bias_kernel<<<grid_for(S * N, 256, 8), 256, 0, st>>>(out, residual, tmp, b, S, N);
  else
    add_bias_kernel<<<grid_for(S * N, 256, 8), 256, 0, st>>>(out, b, S, N, gelu);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_dyg_create(tgm_dyg **out, const tgm_dyg_params *p, int device) {
  TGM_REQUIRE(out != nullptr && p != nullptr, "tgm_dyg_create: NULL argument");
  *out = nullptr;
  TGM_REQUIRE(p->node_dim > 0 && p->edge_dim > 0 && p->time_dim > 0 && p->channel_dim > 0 &&
                  p->out_dim > 0 && p->patch_size > 0 && p->num_layers >= 0 && p->num_heads > 0 &&
                  p->seq_len > 1,
              "tgm_dyg_create: bad sizes");
  if (p->seq_len % p->patch_size)  // dygformer.py:187-188
    return fail(TGM_ERR_INVALID, "Max sequence length must be a multiple of path size");
  TGM_REQUIRE((4 * p->channel_dim) % p->num_heads == 0,
              "tgm_dyg_create: embed_dim must be divisible by num_heads");
  TGM_REQUIRE(device >= 0, "tgm_dyg_create: a CUDA device is required (no CPU fallback)");
  TGM_REQUIRE(p->num_layers == 0 || p->layers != nullptr, "tgm_dyg_create: layers is NULL");
  DeviceGuard g(device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_dyg_create: cannot select device");
  tgm_dyg *m = new (std::nothrow) tgm_dyg();
  if (!m) return fail(TGM_ERR_OOM, "tgm_dyg_create: host allocation failed");
  m->device = device;
  m->dN = p->node_dim, m->dE = p->edge_dim, m->dT = p->time_dim, m->C = p->channel_dim;
  m->out = p->out_dim, m->P = p->patch_size, m->layers = p->num_layers, m->H = p->num_heads;
  m->L = p->seq_len, m->NP = p->seq_len / p->patch_size, m->E = 4 * p->channel_dim;
  m->eps = p->ln_eps;
  const size_t C = m->C, E = m->E, P = m->P;
  const size_t dims[4] = {size_t(m->dN), size_t(m->dE), size_t(m->dT), C};
  int rc = dup(m, &m->tw, p->t2v_w, m->dT);
  if (!rc) rc = dup(m, &m->tb, p->t2v_b, m->dT);
  if (!rc) rc = dup(m, &m->c_w1, p->cooc_w1, C);
  if (!rc) rc = dup(m, &m->c_b1, p->cooc_b1, C);
  if (!rc) rc = dup(m, &m->c_w2, p->cooc_w2, C * C);
  if (!rc) rc = dup(m, &m->c_b2, p->cooc_b2, C);
  for (int c = 0; c < 4 && !rc; ++c) rc = dup(m, &m->proj_w[c], p->proj_w[c], C * P * dims[c]);
  if (!rc) {  // the four projection biases side by side: one add over the [tokens, 4C] layout
    cudaError_t e = cudaMalloc(&m->proj_b, 4 * C * sizeof(float));
    if (e != cudaSuccess) rc = cuda_fail(e, "projection bias", __FILE__, __LINE__);
    else m->owned.push_back(m->proj_b);
    for (int c = 0; c < 4 && !rc; ++c) {
      if (!p->proj_b[c]) rc = fail(TGM_ERR_INVALID, "tgm_dyg_create: NULL parameter");
      else if ((e = cudaMemcpy(m->proj_b + c * C, p->proj_b[c], C * sizeof(float),
                               cudaMemcpyDefault)) != cudaSuccess)
        rc = cuda_fail(e, "projection bias copy", __FILE__, __LINE__);
    }
  }
  if (!rc) rc = dup(m, &m->out_w, p->out_w, size_t(m->out) * E);
  if (!rc) rc = dup(m, &m->out_b, p->out_b, m->out);
  for (int l = 0; l < m->layers && !rc; ++l) {
    const tgm_dyg_layer &s = p->layers[l];
    DygLayerDev d{};
    rc = dup(m, &d.in_w, s.in_proj_w, 3 * E * E);
    if (!rc) rc = dup(m, &d.in_b, s.in_proj_b, 3 * E);
    if (!rc) rc = dup(m, &d.out_w, s.out_proj_w, E * E);
    if (!rc) rc = dup(m, &d.out_b, s.out_proj_b, E);
    if (!rc) rc = dup(m, &d.f1_w, s.ffn1_w, 4 * E * E);
    if (!rc) rc = dup(m, &d.f1_b, s.ffn1_b, 4 * E);
    if (!rc) rc = dup(m, &d.f2_w, s.ffn2_w, 4 * E * E);
    if (!rc) rc = dup(m, &d.f2_b, s.ffn2_b, E);
    if (!rc) rc = dup(m, &d.ln0_w, s.ln0_w, E);
    if (!rc) rc = dup(m, &d.ln0_b, s.ln0_b, E);
    if (!rc) rc = dup(m, &d.ln1_w, s.ln1_w, E);
    if (!rc) rc = dup(m, &d.ln1_b, s.ln1_b, E);
    m->lay.push_back(d);
  }
  if (!rc) {
    cublasStatus_t s = cublasCreate(&m->blas);
    if (s != CUBLAS_STATUS_SUCCESS) rc = blas_fail3(s, "cublasCreate");
    else cublasSetMathMode(m->blas, CUBLAS_PEDANTIC_MATH);
  }
  if (rc) {
    delete m;
    return rc;
  }
  *out = m;
  return TGM_OK;
}

extern "C" void tgm_dyg_destroy(tgm_dyg *m) { delete m; }

extern "C" int tgm_dyg_forward(tgm_dyg *m, const float *node_x, int64_t num_nodes,
                               const int32_t *src, const int32_t *dst, const int64_t *edge_time,
                               const int32_t *nbrs, const int64_t *nbr_t, const float *nbr_x,
                               int64_t B, float *out_src, float *out_dst, tgm_stream stream) {
  TGM_REQUIRE(m != nullptr, "tgm_dyg_forward: handle is NULL");
  TGM_REQUIRE(B >= 0 && num_nodes > 0, "tgm_dyg_forward: bad sizes");
  if (B == 0) return TGM_OK;
  TGM_REQUIRE(node_x && src && dst && edge_time && nbrs && nbr_t && nbr_x && out_src && out_dst,
              "tgm_dyg_forward: NULL array argument");
  DeviceGuard g(m->device);
  cudaStream_t st = as_stream(stream);
  const int L = m->L, NP = m->NP, P = m->P, C = m->C, E = m->E, H = m->H, hd = E / H, T = 2 * NP;
  const int dims[4] = {m->dN, m->dE, m->dT, C};
  const int64_t tokens = B * T;
  TGM_REQUIRE(tokens * 4 * E < (int64_t(1) << 31), "tgm_dyg_forward: batch too large");
  if (B > m->cap) {
    TGM_CUDA(cudaStreamSynchronize(st));
    for (float **p : {&m->feat[0], &m->feat[1], &m->feat[2], &m->feat[3], &m->X, &m->Xn, &m->QKV,
                      &m->S, &m->O, &m->F1, &m->tmp, &m->pooled}) {
      cudaFree(*p);
      *p = nullptr;
    }
    m->cap = 0;
    const size_t cb = size_t(B + B / 4 + 1), ct = cb * T;
    for (int c = 0; c < 4; ++c) TGM_CUDA(cudaMalloc(&m->feat[c], 2 * cb * L * dims[c] * 4));
    TGM_CUDA(cudaMalloc(&m->X, ct * E * 4));
    TGM_CUDA(cudaMalloc(&m->Xn, ct * E * 4));
    TGM_CUDA(cudaMalloc(&m->QKV, ct * 3 * E * 4));
    TGM_CUDA(cudaMalloc(&m->S, cb * H * T * T * 4));
    TGM_CUDA(cudaMalloc(&m->O, ct * E * 4));
    TGM_CUDA(cudaMalloc(&m->F1, ct * 4 * E * 4));
    TGM_CUDA(cudaMalloc(&m->tmp, ct * E * 4));
    TGM_CUDA(cudaMalloc(&m->pooled, 2 * cb * E * 4));
    m->cap = int64_t(cb);
  }
  DYG_BLAS(cublasSetStream(m->blas, st));
  const float one = 1.f, zero = 0.f;

  const size_t smem = size_t(8) * C * sizeof(float);
  TGM_REQUIRE(smem <= 48 * 1024, "tgm_dyg_forward: channel_dim too large");
  dyg_frontend_kernel<<<grid_for(2 * B * L, 8, 8), 256, smem, st>>>(
      node_x, num_nodes, src, dst, edge_time, nbrs, nbr_t, nbr_x, B, L, m->dN, m->dE, m->dT, C,
      m->tw, m->tb, m->c_w1, m->c_b1, m->c_w2, m->c_b2, m->feat[0], m->feat[1], m->feat[2],
      m->feat[3]);
  TGM_LAUNCH_CHECK();

  // channel projections into the token layout X[b, side*NP + p, c*C ...] (dygformer.py:330-413);
  // a patch is P consecutive positions, i.e. a row of the [rows*NP, P*d] view of the channel
  for (int c = 0; c < 4; ++c) {
    const int K = P * dims[c];
    for (int side = 0; side < 2; ++side) {
      const float *A = m->feat[c] + size_t(side) * B * L * dims[c];
      float *Cp = m->X + size_t(side) * NP * E + size_t(c) * C;
      DYG_BLAS(cublasSgemmStridedBatched(m->blas, CUBLAS_OP_T, CUBLAS_OP_N, C, NP, K, &one,
                                         m->proj_w[c], K, 0, A, K, int64_t(NP) * K, &zero, Cp, E,
                                         int64_t(T) * E, int(B)));
    }
  }
  add_bias_kernel<<<grid_for(tokens * E, 256, 8), 256, 0, st>>>(m->X, m->proj_b, tokens, E, 0);
  TGM_LAUNCH_CHECK();

  const float scale = 1.0f / sqrtf(float(hd));
  for (const DygLayerDev &ly : m->lay) {  // TransformerEncoder.forward (:117-143)
    layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(m->X, ly.ln0_w, ly.ln0_b, tokens, E,
                                                             m->eps, m->Xn);
    TGM_LAUNCH_CHECK();
    if (int rc = linear(m->blas, tokens, 3 * E, E, m->Xn, ly.in_w, ly.in_b, nullptr, 0, m->QKV,
                        nullptr, st))
      return rc;
    for (int h = 0; h < H; ++h)  // S[b,h,q,k] = Q_bh[q,:] . K_bh[k,:]
      DYG_BLAS(cublasSgemmStridedBatched(
          m->blas, CUBLAS_OP_T, CUBLAS_OP_N, T, T, hd, &one, m->QKV + E + h * hd, 3 * E,
          int64_t(T) * 3 * E, m->QKV + h * hd, 3 * E, int64_t(T) * 3 * E, &zero,
          m->S + size_t(h) * T * T, T, int64_t(H) * T * T, int(B)));
    softmax_rows_kernel<<<grid_for(B * H * T, 8, 8), 256, 0, st>>>(m->S, B * H * T, T, scale);
    TGM_LAUNCH_CHECK();
    for (int h = 0; h < H; ++h)  // O[b,q,h*hd+d] = sum_k P[b,h,q,k] V_bh[k,d]
      DYG_BLAS(cublasSgemmStridedBatched(
          m->blas, CUBLAS_OP_N, CUBLAS_OP_N, hd, T, T, &one, m->QKV + 2 * E + h * hd, 3 * E,
          int64_t(T) * 3 * E, m->S + size_t(h) * T * T, T, int64_t(H) * T * T, &zero,
          m->O + h * hd, E, int64_t(T) * E, int(B)));
    if (int rc = linear(m->blas, tokens, E, E, m->O, ly.out_w, ly.out_b, m->X, 0, m->X, m->tmp, st))
      return rc;
    layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(m->X, ly.ln1_w, ly.ln1_b, tokens, E,
                                                             m->eps, m->Xn);
    TGM_LAUNCH_CHECK();
    if (int rc = linear(m->blas, tokens, 4 * E, E, m->Xn, ly.f1_w, ly.f1_b, nullptr, 1, m->F1,
                        nullptr, st))
      return rc;
    if (int rc = linear(m->blas, tokens, E, 4 * E, m->F1, ly.f2_w, ly.f2_b, m->X, 0, m->X, m->tmp,
                        st))
      return rc;
  }
  meanpool_kernel<<<grid_for(2 * B * E, 256, 8), 256, 0, st>>>(m->X, B, NP, E, m->pooled);
  TGM_LAUNCH_CHECK();
  DYG_BLAS(gemm_nt3(m->blas, B, m->out, E, m->pooled, m->out_w, out_src));
  DYG_BLAS(gemm_nt3(m->blas, B, m->out, E, m->pooled + size_t(B) * E, m->out_w, out_dst));
  add_bias_kernel<<<grid_for(B * m->out, 256, 8), 256, 0, st>>>(out_src, m->out_b, B, m->out, 0);
  add_bias_kernel<<<grid_for(B * m->out, 256, 8), 256, 0, st>>>(out_dst, m->out_b, B, m->out, 0);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
