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

namespace tgm {
int g_dyg_fused_attn = 1;  // tgm_set_option("dyg_fused_attn", 0|1)
}

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
  // backward (tgm_dyg_backward): per-layer saved activations + gradient scratch, grown on demand
  int64_t bcap = 0;
  struct Saved {
    float *Xin = nullptr, *QKV = nullptr, *Pm = nullptr, *O = nullptr, *X1 = nullptr,
          *F1pre = nullptr;
  };
  std::vector<Saved> saved;
  float *dX = nullptr, *dXn = nullptr, *dQKV = nullptr, *dP = nullptr, *dO = nullptr,
        *dF1 = nullptr, *G1 = nullptr, *dCh = nullptr, *dFeat = nullptr, *Gst = nullptr,
        *dPool = nullptr;
  void free_backward() {
    for (Saved &sv : saved)
      for (float **q : {&sv.Xin, &sv.QKV, &sv.Pm, &sv.O, &sv.X1, &sv.F1pre}) cudaFree(*q), *q = nullptr;
    for (float **q : {&dX, &dXn, &dQKV, &dP, &dO, &dF1, &G1, &dCh, &dFeat, &Gst, &dPool})
      cudaFree(*q), *q = nullptr;
    bcap = 0;
  }
  ~tgm_dyg() {
    if (device >= 0) {
      DeviceGuard g(device);
      free_backward();
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

// Fused self-attention core of one (sequence, head): S = scale * Q K^T, row softmax, O = P V, with Q,
// K, V, S resident in shared memory -- the scores never reach HBM (nn.MultiheadAttention inside
// TransformerEncoder.forward, dygformer.py:117-143; T = 2 * num_patches tokens, hd = E / H).
// Layout in shared memory: Qt[hd][T+4] and Kt[hd][T+4] (d-major, so a thread's 4 queries / 4 keys
// of one feature are ONE 128-bit load and neighbouring threads read neighbouring keys: conflict
// free), V[T][hd] natural, S[T][T+4].  256 threads: 4 x 4 register micro-tiles of S (one per
// thread at T = 64), a warp per 8 rows for the softmax, 4 x 4 micro-tiles of O.  fp32 FMA.
// ALIAS_V (T * hd <= 8192 floats): V does not get its own region -- every thread keeps its (up to
// eight) V granules in registers from the staging loads and writes them over Qt once the scores
// are done.  71.8 KB instead of 97.4 KB at T = 64, hd = 100: three CTAs per SM instead of two, so the
// 400 (sequence, head) CTAs of the default batch are ONE wave of the 148 SMs instead of 1.35.
constexpr int kVHold = 8;
template <bool ALIAS_V>
__global__ void __launch_bounds__(256, ALIAS_V ? 3 : 2)
dyg_attn_core_kernel(const float *__restrict__ QKV, int T, int E, int H, float scale,
                     float *__restrict__ O) {
  extern __shared__ __align__(16) float sm[];
  const int hd = E / H, TS = T + 4;
  float *Qt = sm, *Kt = Qt + hd * TS;
  float *V = ALIAS_V ? Qt : Kt + hd * TS;
  float *Sc = ALIAS_V ? Kt + hd * TS : V + T * hd;
  float4 vhold[kVHold];
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x / H;
  const int h = blockIdx.x % H;
  const float *base = QKV + b * T * 3 * E + h * hd;
  // stage: 128-bit coalesced reads along d, batches of independent loads in flight (one L2
  // latency per batch, not per element); Q and K transposed on the way in
  {
    const int hd4 = hd >> 2, total = T * hd4;
    constexpr int kBatch = 4;
#pragma unroll
    for (int it = 0; it < (ALIAS_V ? kVHold / kBatch : 1); ++it) {
      for (int i0 = tid + it * kBatch * 256; i0 < total;
           i0 += ALIAS_V ? total : kBatch * 256) {  // ALIAS_V: one pass per `it` (total <= 2048)
        float4 q[kBatch], kk[kBatch], vv[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = i0 + u * 256;
          if (i < total) {
            const int t = i / hd4, d4 = i - t * hd4;
            const float4 *row = reinterpret_cast<const float4 *>(base + int64_t(t) * 3 * E) + d4;
            q[u] = __ldg(row);
            kk[u] = __ldg(row + (E >> 2));
            vv[u] = __ldg(row + (E >> 1));
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = i0 + u * 256;
          if (i < total) {
            const int t = i / hd4, d = (i - t * hd4) << 2;
            Qt[(d + 0) * TS + t] = q[u].x * scale;  // the reference scales q before the product
            Qt[(d + 1) * TS + t] = q[u].y * scale;
            Qt[(d + 2) * TS + t] = q[u].z * scale;
            Qt[(d + 3) * TS + t] = q[u].w * scale;
            Kt[(d + 0) * TS + t] = kk[u].x;
            Kt[(d + 1) * TS + t] = kk[u].y;
            Kt[(d + 2) * TS + t] = kk[u].z;
            Kt[(d + 3) * TS + t] = kk[u].w;
            if (ALIAS_V)
              vhold[it * kBatch + u] = vv[u];
            else
              *reinterpret_cast<float4 *>(V + t * hd + d) = vv[u];
          }
        }
      }
    }
  }
  __syncthreads();
  const int tq = T >> 2;  // micro-tiles per dimension
  for (int mt = tid; mt < tq * tq; mt += blockDim.x) {
    const int q0 = (mt / tq) << 2, k0 = (mt % tq) << 2;
    float a[4][4] = {};
#pragma unroll 4
    for (int d = 0; d < hd; ++d) {
      const float4 qv = *reinterpret_cast<const float4 *>(Qt + d * TS + q0);
      const float4 kv = *reinterpret_cast<const float4 *>(Kt + d * TS + k0);
      const float q[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = fmaf(q[i], kk[j], a[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4 *>(Sc + (q0 + i) * TS + k0) = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
  }
  __syncthreads();
  if (ALIAS_V) {  // every thread is past its reads of Qt: V takes the region
    const int hd4 = hd >> 2, total = T * hd4;
#pragma unroll
    for (int j = 0; j < kVHold; ++j) {
      const int i = tid + j * 256;
      if (i < total) {
        const int t = i / hd4, d = (i - t * hd4) << 2;
        *reinterpret_cast<float4 *>(V + t * hd + d) = vhold[j];
      }
    }
  }
  {  // softmax over the keys of every query row: one warp per row
    const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < T; r += nw) {
      float *row = Sc + r * TS;
      float m = -INFINITY;
      for (int c = lane; c < T; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int c = lane; c < T; c += 32) {
        const float e = expf(row[c] - m);
        row[c] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int c = lane; c < T; c += 32) row[c] *= inv;
    }
  }
  __syncthreads();
  const int td = hd >> 2;
  for (int mt = tid; mt < tq * td; mt += blockDim.x) {
    const int q0 = (mt / td) << 2, d0 = (mt % td) << 2;
    float a[4][4] = {};
#pragma unroll 4
    for (int k = 0; k < T; ++k) {
      const float4 vv = *reinterpret_cast<const float4 *>(V + k * hd + d0);
      const float v[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Sc[(q0 + i) * TS + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = fmaf(p, v[j], a[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4 *>(O + (b * T + q0 + i) * E + h * hd + d0) =
          make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
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

// ---- backward pieces (tgm_dyg_backward) -------------------------------------------------------------
// C[M,N] = A[M,K] B[K,N] and C[Ka,Nb] = A[M,Ka]^T B[M,Nb], row-major with leading dimensions
cublasStatus_t gemm_nn(cublasHandle_t h, int64_t M, int N, int K, const float *A, int lda,
                       const float *B, int ldb, float *C, int ldc) {
  const float one = 1.f, zero = 0.f;
  return cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, int(M), K, &one, B, ldb, A, lda, &zero, C, ldc);
}
cublasStatus_t gemm_tn(cublasHandle_t h, int64_t M, int Ka, int Nb, const float *A, int lda,
                       const float *B, int ldb, float *C, int ldc) {
  const float one = 1.f, zero = 0.f;
  return cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, Nb, Ka, int(M), &one, B, ldb, A, lda, &zero, C, ldc);
}

// out[c] += sum_r a[r, c]   (out zeroed by the caller); 32 columns x 8 row lanes per CTA
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ a, int64_t rows, int cols, float *__restrict__ out) {
  __shared__ float s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = int64_t(blockIdx.y) * 8 + ty; r < rows; r += int64_t(gridDim.y) * 8)
      acc += a[r * cols + c];
  s[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += s[j][tx];
    atomicAdd(out + c, t);
  }
}

__global__ void gelu_fwd_kernel(const float *__restrict__ pre, int64_t n, float *__restrict__ act) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float v = pre[i];
    act[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  }
}
// d_pre = d_act * gelu'(pre), gelu'(x) = Phi(x) + x phi(x); in place on d
__global__ void gelu_bwd_kernel(const float *__restrict__ pre, int64_t n, float *__restrict__ d) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float v = pre[i];
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    d[i] *= cdf + v * pdf;
  }
}
// LayerNorm backward, one warp per row: dx_total[r,:] += rstd * (dxh - mean(dxh) - xh * mean(dxh*xh))
// with xh, rstd recomputed from x; dgamma/dbeta accumulated per CTA in shared memory, then atomics.
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
                     const float *__restrict__ dy, int64_t rows, int cols, float eps,
                     float *__restrict__ dx_total, float *__restrict__ dgamma,
                     float *__restrict__ dbeta) {
  extern __shared__ float s_acc[];  // [2][cols]
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows;
       r += int64_t(gridDim.x) * wpb) {
    const float *xr = x + r * cols, *dr = dy + r * cols;
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
    float m1 = 0.f, m2 = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float xh = (xr[c] - mean) * rstd, dxh = dr[c] * __ldg(w + c);
      m1 += dxh;
      m2 = fmaf(dxh, xh, m2);
      atomicAdd(&s_acc[c], dr[c] * xh);
      atomicAdd(&s_acc[cols + c], dr[c]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 /= float(cols), m2 /= float(cols);
    for (int c = lane; c < cols; c += 32) {
      const float xh = (xr[c] - mean) * rstd, dxh = dr[c] * __ldg(w + c);
      dx_total[r * cols + c] += rstd * (dxh - m1 - xh * m2);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    atomicAdd(dgamma + c, s_acc[c]);
    atomicAdd(dbeta + c, s_acc[cols + c]);
  }
}

// softmax backward in place on dp: ds = scale * p * (dp - sum(dp * p)); one warp per row
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float *__restrict__ p, float *__restrict__ dp, int64_t rows, int cols,
                   float scale) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows;
       r += int64_t(gridDim.x) * wpb) {
    const float *pr = p + r * cols;
    float *dr = dp + r * cols;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot = fmaf(dr[c], pr[c], dot);
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int c = lane; c < cols; c += 32) dr[c] = scale * pr[c] * (dr[c] - dot);
  }
}

// dX[b, side*NP + p, :] = dpooled[side*B + b, :] / NP
__global__ void meanpool_bwd_kernel(const float *__restrict__ dpooled, int64_t B, int NP, int E,
                                    float *__restrict__ dX) {
  const int64_t total = B * 2 * NP * E;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % E);
    const int64_t tok = i / E, b = tok / (2 * NP);
    const int side = int((tok - b * 2 * NP) / NP);
    dX[i] = dpooled[(int64_t(side) * B + b) * E + c] / float(NP);
  }
}

// channel c of the token gradient as a dense [2B*NP, C] matrix in the (side, b, patch) row order of
// the feature tensors: dCh[(side*B + b)*NP + p, j] = dX[(b*2NP + side*NP + p)*E + c*C + j]
__global__ void gather_channel_kernel(const float *__restrict__ dX, int64_t B, int NP, int E, int C,
                                      int c, float *__restrict__ dCh) {
  const int64_t total = 2 * B * NP * C;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(i % C);
    const int64_t row = i / C, sb = row / NP, p = row - sb * NP;
    const int64_t side = sb / B, b = sb - side * B;
    dCh[i] = dX[(b * 2 * NP + side * NP + p) * E + int64_t(c) * C + j];
  }
}

// Gradients of the Time2Vec weights and of the co-occurrence MLP from the per-position feature
// gradients; one warp per sequence position (same indexing as dyg_frontend_kernel), accumulation in
// shared memory per CTA, then global atomics.  d_time may be NULL-free: both are required.
__global__ void __launch_bounds__(256)
dyg_frontend_bwd_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                        const int64_t *__restrict__ edge_time, const int32_t *__restrict__ nbrs,
                        const int64_t *__restrict__ nbr_t, int64_t B, int L, int dT, int C,
                        const float *__restrict__ tw, const float *__restrict__ tb,
                        const float *__restrict__ w1, const float *__restrict__ b1,
                        const float *__restrict__ w2, const float *__restrict__ d_time,
                        const float *__restrict__ d_cooc, float *__restrict__ g_tw,
                        float *__restrict__ g_tb, float *__restrict__ g_w1, float *__restrict__ g_b1,
                        float *__restrict__ g_w2, float *__restrict__ g_b2) {
  extern __shared__ float sm[];
  float *a_tw = sm, *a_tb = a_tw + dT, *a_w1 = a_tb + dT, *a_b1 = a_w1 + C, *a_b2 = a_b1 + C,
        *a_w2 = a_b2 + C, *scratch = a_w2 + C * C;  // scratch: [warps][2C] (hsum, dh)
  const int nacc = 2 * dT + 3 * C + C * C;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float *hs = scratch + warp * 2 * C, *dh = hs + C;
  const int k = L - 1;
  const int64_t total = 2 * B * L;
  for (int64_t i = int64_t(blockIdx.x) * wpb + warp; i < total; i += int64_t(gridDim.x) * wpb) {
    const int64_t r = i / L;
    const int l = int(i - r * L);
    const int64_t pair = r < B ? r : r - B, other = r < B ? r + B : r - B;
    const int32_t id = seq_id(src, dst, nbrs, B, k, r, l);
    const bool padded = id == TGM_PADDED_NODE_ID;
    if (!padded) {  // time channel: f = cos(dt w + b)  ->  d/dw = -sin(.) dt, d/db = -sin(.)
      const int64_t tq = edge_time[pair];
      const float dt = float(tq - (l == 0 ? tq : nbr_t[r * k + (l - 1)]));
      for (int c = lane; c < dT; c += 32) {
        const float d = -sinf(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c))) * d_time[i * dT + c];
        atomicAdd(a_tw + c, d * dt);
        atomicAdd(a_tb + c, d);
      }
    }
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
    const float *dc = d_cooc + i * C;
    for (int j = lane; j < C; j += 32) {
      const float w = __ldg(w1 + j), bb = __ldg(b1 + j);
      hs[j] = fmaxf(fmaf(w, f0, bb), 0.f) + fmaxf(fmaf(w, f1, bb), 0.f);
      float acc = 0.f;  // dh[j] = sum_c W2[c, j] dc[c]
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w2 + c * C + j), dc[c], acc);
      dh[j] = acc;
      const float m0 = fmaf(w, f0, bb) > 0.f ? 1.f : 0.f, m1 = fmaf(w, f1, bb) > 0.f ? 1.f : 0.f;
      atomicAdd(a_w1 + j, acc * (m0 * f0 + m1 * f1));
      atomicAdd(a_b1 + j, acc * (m0 + m1));
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) atomicAdd(a_b2 + c, 2.f * dc[c]);
    for (int e = lane; e < C * C; e += 32) {  // dW2[c, j] += dc[c] * hsum[j]
      const int c = e / C, j = e - c * C;
      atomicAdd(a_w2 + e, dc[c] * hs[j]);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dT; i += blockDim.x) {
    atomicAdd(g_tw + i, a_tw[i]);
    atomicAdd(g_tb + i, a_tb[i]);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(g_w1 + i, a_w1[i]);
    atomicAdd(g_b1 + i, a_b1[i]);
    atomicAdd(g_b2 + i, a_b2[i]);
  }
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) atomicAdd(g_w2 + i, a_w2[i]);
}

// out[S,N] = act(A[S,K] W[N,K]^T + b (+ residual)); residual may alias out, `tmp` is [S,N] scratch.
// Token-sized linears run on the tensor cores with the bias / residual / GELU fused into the
// epilogue (gemm_fastf32.cu: fp32-accurate 9xBF16 emulation on tcgen05, 1.6x cuBLAS's SIMT SGEMM at
// a quarter of its rounding error); small or unaligned ones use cuBLAS (true fp32) + one
// elementwise pass.
int linear(cublasHandle_t blas, int64_t S, int N, int K, const float *A, const float *W,
           const float *b, const float *residual, int gelu, float *out, float *tmp,
           cudaStream_t st) {
  // hand-written tcgen05 kernel, 3xTF32 split (tc_linear.cu): every token linear (1, the default)
  // or only the GELU-fused FFN linear with the others on the CUTLASS collective (2: the mixed
  // policy measured 3.9 % faster per forward at 12800 tokens, 0.743 vs 0.772 ms; the collective
  // alone: 0.770)
  if ((g_tc_linear == 1 || (g_tc_linear == 2 && gelu)) && S >= 2048) {
    const int rc = tc3_linear(S, N, K, A, W, b, residual, gelu, out, st);
    if (rc != 0) return rc < 0 ? rc : TGM_OK;
  }
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

int reserve_forward(tgm_dyg *m, int64_t B, cudaStream_t st) {
  const int L = m->L, E = m->E, H = m->H, T = 2 * m->NP;
  const int dims[4] = {m->dN, m->dE, m->dT, m->C};
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
  if (int rc = reserve_forward(m, B, st)) return rc;
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
  // a patch is P consecutive positions, i.e. a row of the [rows*NP, P*d] view of the channel.
  // The 8 B small products (4 channels x 2 sides x B pairs, NP x C x P d_c each) and their bias go
  // out as ONE launch of the grouped short-matrix kernel (small_gemm.cu); the library form is
  // eight strided-batched SGEMMs + a bias pass.
  int proj_rc = 0;
  if (2 * B <= 8000) {
    SmallGemmGroups pg{};
    pg.n = 8;
    for (int c = 0; c < 4; ++c) {
      const int K = P * dims[c];
      for (int side = 0; side < 2; ++side) {
        const int gi = 2 * c + side;
        pg.A[gi] = m->feat[c] + size_t(side) * B * L * dims[c];
        pg.W[gi] = m->proj_w[c], pg.bias[gi] = m->proj_b + c * C;
        pg.C[gi] = m->X + size_t(side) * NP * E + size_t(c) * C;
        pg.K[gi] = K, pg.lda[gi] = K, pg.ldw[gi] = K, pg.batch[gi] = int(B);
        pg.strideA[gi] = int64_t(NP) * K, pg.strideC[gi] = int64_t(T) * E;
      }
    }
    proj_rc = small_gemm_groups(pg, NP, C, E, 0, st);
    if (proj_rc < 0) return proj_rc;
  }
  if (proj_rc == 0) {
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
  }

  const float scale = 1.0f / sqrtf(float(hd));
  for (const DygLayerDev &ly : m->lay) {  // TransformerEncoder.forward (:117-143)
    layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(m->X, ly.ln0_w, ly.ln0_b, tokens, E,
                                                             m->eps, m->Xn);
    TGM_LAUNCH_CHECK();
    if (int rc = linear(m->blas, tokens, 3 * E, E, m->Xn, ly.in_w, ly.in_b, nullptr, 0, m->QKV,
                        nullptr, st))
      return rc;
    const bool alias_v = int64_t(T) * hd <= int64_t(kVHold) * 256 * 4;
    const size_t attn_smem =
        (size_t(2) * hd * (T + 4) + (alias_v ? 0 : size_t(T) * hd) + size_t(T) * (T + 4)) * 4;
    if (g_dyg_fused_attn && T % 4 == 0 && hd % 4 == 0 && E % 4 == 0 && attn_smem <= 200 * 1024) {
      // QK^T, softmax and PV of every (sequence, head) in one kernel, scores on chip
      if (alias_v) {
        if (attn_smem > 48 * 1024)
          TGM_CUDA(cudaFuncSetAttribute(dyg_attn_core_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_smem)));
        dyg_attn_core_kernel<true><<<int(B * H), 256, attn_smem, st>>>(m->QKV, T, E, H, scale, m->O);
      } else {
        if (attn_smem > 48 * 1024)
          TGM_CUDA(cudaFuncSetAttribute(dyg_attn_core_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_smem)));
        dyg_attn_core_kernel<false><<<int(B * H), 256, attn_smem, st>>>(m->QKV, T, E, H, scale, m->O);
      }
      TGM_LAUNCH_CHECK();
    } else {
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
    }
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

// ---- training support ----------------------------------------------------------------------------
namespace {

int copy_param(float *dst, const float *src, size_t n, cudaStream_t st) {
  TGM_REQUIRE(src != nullptr, "tgm_dyg_set_params: NULL parameter");
  if (n) TGM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDefault, st));
  return TGM_OK;
}

int reserve_backward(tgm_dyg *m, int64_t B, cudaStream_t st) {
  if (B <= m->bcap && int(m->saved.size()) == m->layers) return TGM_OK;
  TGM_CUDA(cudaStreamSynchronize(st));
  m->free_backward();
  m->saved.assign(m->layers, tgm_dyg::Saved{});
  const int L = m->L, E = m->E, H = m->H, T = 2 * m->NP, C = m->C;
  const size_t cb = size_t(B + B / 4 + 1), ct = cb * T;
  const int dmax = std::max(m->dT, C);
  for (tgm_dyg::Saved &sv : m->saved) {
    TGM_CUDA(cudaMalloc(&sv.Xin, ct * E * 4));
    TGM_CUDA(cudaMalloc(&sv.QKV, ct * 3 * E * 4));
    TGM_CUDA(cudaMalloc(&sv.Pm, cb * H * T * T * 4));
    TGM_CUDA(cudaMalloc(&sv.O, ct * E * 4));
    TGM_CUDA(cudaMalloc(&sv.X1, ct * E * 4));
    TGM_CUDA(cudaMalloc(&sv.F1pre, ct * 4 * E * 4));
  }
  TGM_CUDA(cudaMalloc(&m->dX, ct * E * 4));
  TGM_CUDA(cudaMalloc(&m->dXn, ct * E * 4));
  TGM_CUDA(cudaMalloc(&m->dQKV, ct * 3 * E * 4));
  TGM_CUDA(cudaMalloc(&m->dP, cb * H * T * T * 4));
  TGM_CUDA(cudaMalloc(&m->dO, ct * E * 4));
  TGM_CUDA(cudaMalloc(&m->dF1, ct * 4 * E * 4));
  TGM_CUDA(cudaMalloc(&m->G1, ct * 4 * E * 4));
  TGM_CUDA(cudaMalloc(&m->dCh, 2 * cb * m->NP * C * 4));
  TGM_CUDA(cudaMalloc(&m->dFeat, 2 * cb * L * dmax * 4 * 2));  // time | co-occurrence
  TGM_CUDA(cudaMalloc(&m->Gst, 2 * cb * m->out * 4));
  TGM_CUDA(cudaMalloc(&m->dPool, 2 * cb * E * 4));
  m->bcap = int64_t(cb);
  return TGM_OK;
}

int colsum(const float *a, int64_t rows, int cols, float *out, cudaStream_t st) {
  TGM_CUDA(cudaMemsetAsync(out, 0, size_t(cols) * sizeof(float), st));
  if (rows == 0) return TGM_OK;
  const dim3 grid((cols + 31) / 32, unsigned(std::min<int64_t>((rows + 63) / 64, 256)));
  colsum_kernel<<<grid, 256, 0, st>>>(a, rows, cols, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_dyg_set_params(tgm_dyg *m, const tgm_dyg_params *p, tgm_stream stream) {
  TGM_REQUIRE(m != nullptr && p != nullptr, "tgm_dyg_set_params: NULL argument");
  TGM_REQUIRE(p->node_dim == m->dN && p->edge_dim == m->dE && p->time_dim == m->dT &&
                  p->channel_dim == m->C && p->out_dim == m->out && p->patch_size == m->P &&
                  p->num_layers == m->layers && p->num_heads == m->H && p->seq_len == m->L,
              "tgm_dyg_set_params: shapes differ from the handle's");
  TGM_REQUIRE(m->layers == 0 || p->layers != nullptr, "tgm_dyg_set_params: layers is NULL");
  DeviceGuard g(m->device);
  cudaStream_t st = as_stream(stream);
  const size_t C = m->C, E = m->E, P = m->P;
  const size_t dims[4] = {size_t(m->dN), size_t(m->dE), size_t(m->dT), C};
  int rc = copy_param(m->tw, p->t2v_w, m->dT, st);
  if (!rc) rc = copy_param(m->tb, p->t2v_b, m->dT, st);
  if (!rc) rc = copy_param(m->c_w1, p->cooc_w1, C, st);
  if (!rc) rc = copy_param(m->c_b1, p->cooc_b1, C, st);
  if (!rc) rc = copy_param(m->c_w2, p->cooc_w2, C * C, st);
  if (!rc) rc = copy_param(m->c_b2, p->cooc_b2, C, st);
  for (int c = 0; c < 4 && !rc; ++c) {
    rc = copy_param(m->proj_w[c], p->proj_w[c], C * P * dims[c], st);
    if (!rc) rc = copy_param(m->proj_b + c * C, p->proj_b[c], C, st);
  }
  if (!rc) rc = copy_param(m->out_w, p->out_w, size_t(m->out) * E, st);
  if (!rc) rc = copy_param(m->out_b, p->out_b, m->out, st);
  for (int l = 0; l < m->layers && !rc; ++l) {
    const tgm_dyg_layer &s = p->layers[l];
    DygLayerDev &d = m->lay[l];
    rc = copy_param(d.in_w, s.in_proj_w, 3 * E * E, st);
    if (!rc) rc = copy_param(d.in_b, s.in_proj_b, 3 * E, st);
    if (!rc) rc = copy_param(d.out_w, s.out_proj_w, E * E, st);
    if (!rc) rc = copy_param(d.out_b, s.out_proj_b, E, st);
    if (!rc) rc = copy_param(d.f1_w, s.ffn1_w, 4 * E * E, st);
    if (!rc) rc = copy_param(d.f1_b, s.ffn1_b, 4 * E, st);
    if (!rc) rc = copy_param(d.f2_w, s.ffn2_w, 4 * E * E, st);
    if (!rc) rc = copy_param(d.f2_b, s.ffn2_b, E, st);
    if (!rc) rc = copy_param(d.ln0_w, s.ln0_w, E, st);
    if (!rc) rc = copy_param(d.ln0_b, s.ln0_b, E, st);
    if (!rc) rc = copy_param(d.ln1_w, s.ln1_w, E, st);
    if (!rc) rc = copy_param(d.ln1_b, s.ln1_b, E, st);
  }
  return rc;
}

extern "C" int tgm_dyg_backward(tgm_dyg *m, const float *node_x, int64_t num_nodes,
                                const int32_t *src, const int32_t *dst, const int64_t *edge_time,
                                const int32_t *nbrs, const int64_t *nbr_t, const float *nbr_x,
                                int64_t B, const float *d_src, const float *d_dst,
                                const tgm_dyg_grads *gr, tgm_stream stream) {
  TGM_REQUIRE(m != nullptr && gr != nullptr, "tgm_dyg_backward: NULL handle or gradient table");
  TGM_REQUIRE(B > 0 && num_nodes > 0, "tgm_dyg_backward: bad sizes");
  TGM_REQUIRE(node_x && src && dst && edge_time && nbrs && nbr_t && nbr_x && d_src && d_dst,
              "tgm_dyg_backward: NULL array argument");
  TGM_REQUIRE(m->layers == 0 || gr->layers != nullptr, "tgm_dyg_backward: gradient layers is NULL");
  DeviceGuard g(m->device);
  cudaStream_t st = as_stream(stream);
  const int L = m->L, NP = m->NP, P = m->P, C = m->C, E = m->E, H = m->H, hd = E / H, T = 2 * NP;
  const int dims[4] = {m->dN, m->dE, m->dT, C};
  const int64_t tokens = B * T, rowsF = 2 * B * L, rowsP = 2 * B * NP;
  TGM_REQUIRE(tokens * 4 * E < (int64_t(1) << 31), "tgm_dyg_backward: batch too large");
  if (int rc = reserve_forward(m, B, st)) return rc;
  if (int rc = reserve_backward(m, B, st)) return rc;
  DYG_BLAS(cublasSetStream(m->blas, st));
  const float one = 1.f, zero = 0.f;
  const float scale = 1.0f / sqrtf(float(hd));
  const size_t tokE = size_t(tokens) * E * sizeof(float);

  // ---- forward, keeping what the chain rule needs (layer inputs, QKV, probabilities, O, the
  //      post-attention stream, the pre-GELU activations); normalised inputs are recomputed ----
  {
    const size_t smem = size_t(8) * C * sizeof(float);
    TGM_REQUIRE(smem <= 48 * 1024, "tgm_dyg_backward: channel_dim too large");
    dyg_frontend_kernel<<<grid_for(rowsF, 8, 8), 256, smem, st>>>(
        node_x, num_nodes, src, dst, edge_time, nbrs, nbr_t, nbr_x, B, L, m->dN, m->dE, m->dT, C,
        m->tw, m->tb, m->c_w1, m->c_b1, m->c_w2, m->c_b2, m->feat[0], m->feat[1], m->feat[2],
        m->feat[3]);
    TGM_LAUNCH_CHECK();
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
    for (int l = 0; l < m->layers; ++l) {
      const DygLayerDev &ly = m->lay[l];
      tgm_dyg::Saved &sv = m->saved[l];
      TGM_CUDA(cudaMemcpyAsync(sv.Xin, m->X, tokE, cudaMemcpyDeviceToDevice, st));
      layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(m->X, ly.ln0_w, ly.ln0_b, tokens, E,
                                                               m->eps, m->Xn);
      TGM_LAUNCH_CHECK();
      if (int rc = linear(m->blas, tokens, 3 * E, E, m->Xn, ly.in_w, ly.in_b, nullptr, 0, sv.QKV,
                          nullptr, st))
        return rc;
      for (int h = 0; h < H; ++h)
        DYG_BLAS(cublasSgemmStridedBatched(
            m->blas, CUBLAS_OP_T, CUBLAS_OP_N, T, T, hd, &one, sv.QKV + E + h * hd, 3 * E,
            int64_t(T) * 3 * E, sv.QKV + h * hd, 3 * E, int64_t(T) * 3 * E, &zero,
            sv.Pm + size_t(h) * T * T, T, int64_t(H) * T * T, int(B)));
      softmax_rows_kernel<<<grid_for(B * H * T, 8, 8), 256, 0, st>>>(sv.Pm, B * H * T, T, scale);
      TGM_LAUNCH_CHECK();
      for (int h = 0; h < H; ++h)
        DYG_BLAS(cublasSgemmStridedBatched(
            m->blas, CUBLAS_OP_N, CUBLAS_OP_N, hd, T, T, &one, sv.QKV + 2 * E + h * hd, 3 * E,
            int64_t(T) * 3 * E, sv.Pm + size_t(h) * T * T, T, int64_t(H) * T * T, &zero,
            sv.O + h * hd, E, int64_t(T) * E, int(B)));
      if (int rc = linear(m->blas, tokens, E, E, sv.O, ly.out_w, ly.out_b, m->X, 0, m->X, m->tmp, st))
        return rc;
      TGM_CUDA(cudaMemcpyAsync(sv.X1, m->X, tokE, cudaMemcpyDeviceToDevice, st));
      layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(m->X, ly.ln1_w, ly.ln1_b, tokens, E,
                                                               m->eps, m->Xn);
      TGM_LAUNCH_CHECK();
      if (int rc = linear(m->blas, tokens, 4 * E, E, m->Xn, ly.f1_w, ly.f1_b, nullptr, 0, sv.F1pre,
                          nullptr, st))
        return rc;
      gelu_fwd_kernel<<<grid_for(tokens * 4 * E, 256, 8), 256, 0, st>>>(sv.F1pre, tokens * 4 * E,
                                                                       m->F1);
      TGM_LAUNCH_CHECK();
      if (int rc = linear(m->blas, tokens, E, 4 * E, m->F1, ly.f2_w, ly.f2_b, m->X, 0, m->X, m->tmp,
                          st))
        return rc;
    }
    meanpool_kernel<<<grid_for(2 * B * E, 256, 8), 256, 0, st>>>(m->X, B, NP, E, m->pooled);
    TGM_LAUNCH_CHECK();
  }

  // ---- output layer and mean pooling ------------------------------------------------------------
  TGM_CUDA(cudaMemcpyAsync(m->Gst, d_src, size_t(B) * m->out * 4, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(m->Gst + size_t(B) * m->out, d_dst, size_t(B) * m->out * 4,
                           cudaMemcpyDeviceToDevice, st));
  DYG_BLAS(gemm_tn(m->blas, 2 * B, m->out, E, m->Gst, m->out, m->pooled, E, gr->out_w, E));
  if (int rc = colsum(m->Gst, 2 * B, m->out, gr->out_b, st)) return rc;
  DYG_BLAS(gemm_nn(m->blas, 2 * B, E, m->out, m->Gst, m->out, m->out_w, E, m->dPool, E));
  meanpool_bwd_kernel<<<grid_for(tokens * E, 256, 8), 256, 0, st>>>(m->dPool, B, NP, E, m->dX);
  TGM_LAUNCH_CHECK();

  // ---- transformer layers, last to first ------------------------------------------------------
  const size_t ln_smem = size_t(2) * E * sizeof(float);
  for (int l = m->layers - 1; l >= 0; --l) {
    const DygLayerDev &ly = m->lay[l];
    const tgm_dyg::Saved &sv = m->saved[l];
    const tgm_dyg_layer_grads &gl = gr->layers[l];
    // FFN2: X2 = X1 + G1 W2^T + b2
    gelu_fwd_kernel<<<grid_for(tokens * 4 * E, 256, 8), 256, 0, st>>>(sv.F1pre, tokens * 4 * E,
                                                                     m->G1);
    TGM_LAUNCH_CHECK();
    DYG_BLAS(gemm_tn(m->blas, tokens, E, 4 * E, m->dX, E, m->G1, 4 * E, gl.ffn2_w, 4 * E));
    if (int rc = colsum(m->dX, tokens, E, gl.ffn2_b, st)) return rc;
    DYG_BLAS(gemm_nn(m->blas, tokens, 4 * E, E, m->dX, E, ly.f2_w, 4 * E, m->dF1, 4 * E));
    gelu_bwd_kernel<<<grid_for(tokens * 4 * E, 256, 8), 256, 0, st>>>(sv.F1pre, tokens * 4 * E,
                                                                     m->dF1);
    TGM_LAUNCH_CHECK();
    // FFN1: F1pre = LN1(X1) W1^T + b1
    layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(sv.X1, ly.ln1_w, ly.ln1_b, tokens, E,
                                                             m->eps, m->Xn);
    TGM_LAUNCH_CHECK();
    DYG_BLAS(gemm_tn(m->blas, tokens, 4 * E, E, m->dF1, 4 * E, m->Xn, E, gl.ffn1_w, E));
    if (int rc = colsum(m->dF1, tokens, 4 * E, gl.ffn1_b, st)) return rc;
    DYG_BLAS(gemm_nn(m->blas, tokens, E, 4 * E, m->dF1, 4 * E, ly.f1_w, E, m->dXn, E));
    TGM_CUDA(cudaMemsetAsync(gl.ln1_w, 0, size_t(E) * 4, st));
    TGM_CUDA(cudaMemsetAsync(gl.ln1_b, 0, size_t(E) * 4, st));
    layernorm_bwd_kernel<<<grid_for(tokens, 8, 4), 256, ln_smem, st>>>(
        sv.X1, ly.ln1_w, m->dXn, tokens, E, m->eps, m->dX, gl.ln1_w, gl.ln1_b);  // dX is now dX1
    TGM_LAUNCH_CHECK();
    // attention output projection: X1 = Xin + O Wout^T + bout
    DYG_BLAS(gemm_tn(m->blas, tokens, E, E, m->dX, E, sv.O, E, gl.out_proj_w, E));
    if (int rc = colsum(m->dX, tokens, E, gl.out_proj_b, st)) return rc;
    DYG_BLAS(gemm_nn(m->blas, tokens, E, E, m->dX, E, ly.out_w, E, m->dO, E));
    // attention core, per head batched over the B sequences pairs
    for (int h = 0; h < H; ++h) {
      const float *V = sv.QKV + 2 * E + h * hd;
      float *dV = m->dQKV + 2 * E + h * hd;
      const float *Pm = sv.Pm + size_t(h) * T * T;
      float *dPh = m->dP + size_t(h) * T * T;
      const float *dOh = m->dO + h * hd;
      const int64_t sQ = int64_t(T) * 3 * E, sP = int64_t(H) * T * T, sO = int64_t(T) * E;
      // dP[q,k] = sum_d dO[q,d] V[k,d]
      DYG_BLAS(cublasSgemmStridedBatched(m->blas, CUBLAS_OP_T, CUBLAS_OP_N, T, T, hd, &one, V, 3 * E,
                                         sQ, dOh, E, sO, &zero, dPh, T, sP, int(B)));
      // dV[k,d] = sum_q P[q,k] dO[q,d]
      DYG_BLAS(cublasSgemmStridedBatched(m->blas, CUBLAS_OP_N, CUBLAS_OP_T, hd, T, T, &one, dOh, E,
                                         sO, Pm, T, sP, &zero, dV, 3 * E, sQ, int(B)));
    }
    softmax_bwd_kernel<<<grid_for(B * H * T, 8, 8), 256, 0, st>>>(sv.Pm, m->dP, B * H * T, T, scale);
    TGM_LAUNCH_CHECK();
    for (int h = 0; h < H; ++h) {
      const float *Q = sv.QKV + h * hd, *K = sv.QKV + E + h * hd;
      float *dQ = m->dQKV + h * hd, *dK = m->dQKV + E + h * hd;
      const float *dS = m->dP + size_t(h) * T * T;
      const int64_t sQ = int64_t(T) * 3 * E, sP = int64_t(H) * T * T;
      // dQ[q,d] = sum_k dS[q,k] K[k,d]
      DYG_BLAS(cublasSgemmStridedBatched(m->blas, CUBLAS_OP_N, CUBLAS_OP_N, hd, T, T, &one, K, 3 * E,
                                         sQ, dS, T, sP, &zero, dQ, 3 * E, sQ, int(B)));
      // dK[k,d] = sum_q dS[q,k] Q[q,d]
      DYG_BLAS(cublasSgemmStridedBatched(m->blas, CUBLAS_OP_N, CUBLAS_OP_T, hd, T, T, &one, Q, 3 * E,
                                         sQ, dS, T, sP, &zero, dK, 3 * E, sQ, int(B)));
    }
    // in_proj: QKV = LN0(Xin) Win^T + bin
    layernorm_kernel<<<grid_for(tokens, 8, 8), 256, 0, st>>>(sv.Xin, ly.ln0_w, ly.ln0_b, tokens, E,
                                                             m->eps, m->Xn);
    TGM_LAUNCH_CHECK();
    DYG_BLAS(gemm_tn(m->blas, tokens, 3 * E, E, m->dQKV, 3 * E, m->Xn, E, gl.in_proj_w, E));
    if (int rc = colsum(m->dQKV, tokens, 3 * E, gl.in_proj_b, st)) return rc;
    DYG_BLAS(gemm_nn(m->blas, tokens, E, 3 * E, m->dQKV, 3 * E, ly.in_w, E, m->dXn, E));
    TGM_CUDA(cudaMemsetAsync(gl.ln0_w, 0, size_t(E) * 4, st));
    TGM_CUDA(cudaMemsetAsync(gl.ln0_b, 0, size_t(E) * 4, st));
    layernorm_bwd_kernel<<<grid_for(tokens, 8, 4), 256, ln_smem, st>>>(
        sv.Xin, ly.ln0_w, m->dXn, tokens, E, m->eps, m->dX, gl.ln0_w, gl.ln0_b);  // dX is now dXin
    TGM_LAUNCH_CHECK();
  }

  // ---- channel projections and the two learnable feature channels -------------------------------
  const int dmax = std::max(m->dT, C);
  float *dTime = m->dFeat, *dCooc = m->dFeat + size_t(rowsF) * dmax;
  for (int c = 0; c < 4; ++c) {
    const int K = P * dims[c];
    gather_channel_kernel<<<grid_for(rowsP * C, 256, 8), 256, 0, st>>>(m->dX, B, NP, E, C, c, m->dCh);
    TGM_LAUNCH_CHECK();
    // patches of channel c: the [2B*NP, P*d_c] view of feat[c]
    DYG_BLAS(gemm_tn(m->blas, rowsP, C, K, m->dCh, C, m->feat[c], K, gr->proj_w[c], K));
    if (int rc = colsum(m->dCh, rowsP, C, gr->proj_b[c], st)) return rc;
    if (c >= 2)
      DYG_BLAS(gemm_nn(m->blas, rowsP, K, C, m->dCh, C, m->proj_w[c], K, c == 2 ? dTime : dCooc, K));
  }
  for (float *q : {gr->t2v_w, gr->t2v_b})
    TGM_CUDA(cudaMemsetAsync(q, 0, size_t(m->dT) * 4, st));
  for (float *q : {gr->cooc_w1, gr->cooc_b1, gr->cooc_b2})
    TGM_CUDA(cudaMemsetAsync(q, 0, size_t(C) * 4, st));
  TGM_CUDA(cudaMemsetAsync(gr->cooc_w2, 0, size_t(C) * C * 4, st));
  {
    const int wpb = 8;
    const size_t smem = (size_t(2) * m->dT + 3 * C + size_t(C) * C + size_t(wpb) * 2 * C) * sizeof(float);
    TGM_REQUIRE(smem <= 96 * 1024, "tgm_dyg_backward: channel_dim too large for the frontend backward");
    if (smem > 48 * 1024)
      TGM_CUDA(cudaFuncSetAttribute(dyg_frontend_bwd_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    dyg_frontend_bwd_kernel<<<grid_for(rowsF, wpb, 2), 256, smem, st>>>(
        src, dst, edge_time, nbrs, nbr_t, B, L, m->dT, C, m->tw, m->tb, m->c_w1, m->c_b1, m->c_w2,
        dTime, dCooc, gr->t2v_w, gr->t2v_b, gr->cooc_w1, gr->cooc_b1, gr->cooc_w2, gr->cooc_b2);
    TGM_LAUNCH_CHECK();
  }
  return TGM_OK;
}
