// Time-encoded single-query multi-head attention over sampled neighbours (TGAT aggregation).
//
// Replaces TemporalAttention.forward + the two Time2Vec calls feeding it (reference
// tgm-team/tgm @ 5183dc9: tgm/nn/modules/attention.py:58-128, tgm/nn/modules/time_encoding.py:
// 22-24, call site tgm/nn/encoder/tgat.py:136-147) and MergeLayer (tgat.py:11-38).
//
// The reference materialises Z = cat[nbr_node_feat, edge_feat, Time2Vec(dt)] (S,k,key_dim) and
// pushes all S*k rows through W_KV (key_dim -> 2*out).  Because there is ONE query per seed the
// contraction reassociates exactly (W_KV has no bias, attention.py:52):
//     Q_h . K_n   = Q_h . (W_K,h z_n)        = (W_K,h^T Q_h) . z_n          =: qk_h . z_n
//     sum_n a_hn V_n = sum_n a_hn W_V,h z_n  = W_V,h (sum_n a_hn z_n)       =: W_V,h u_h
// so the per-neighbour GEMM disappears: per seed two skinny GEMMs (out x key_dim) and a fused
// HBM-streaming kernel that gathers the neighbour rows, evaluates Time2Vec in registers, takes
// the masked softmax and accumulates u_h -- Z, K and V never exist in memory.  k-fold fewer
// flops than the reference; fp32 FMA throughout (TF32/BF16 would break the 1e-5 bar).
//
// Dense plain GEMMs (S rows) go through cuBLAS SGEMM (default math: true fp32).
#include <cublas_v2.h>

#include <algorithm>
#include <cmath>
#include <new>

#include "common.cuh"

using namespace tgm;

#include "attention.cuh"

namespace {

int blas_fail(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define TGM_BLAS(expr)                                          \
  do {                                                          \
    cublasStatus_t _s = (expr);                                 \
    if (_s != CUBLAS_STATUS_SUCCESS) return blas_fail(_s, #expr); \
  } while (0)

// C[S,N] = A[S,K] . W[N,K]^T   (all row-major)
cublasStatus_t gemm_nt(cublasHandle_t h, int64_t S, int N, int K, const float *A, int lda,
                       const float *W, float *C, int ldc) {
  const float one = 1.f, zero = 0.f;
  return cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, int(S), K, &one, W, K, A, lda, &zero, C, ldc);
}

// R[s] = [X[s] | 0 (pad) | Time2Vec(0)]   (attention.py:93-95, tgat.py:141)
__global__ void attn_residual_kernel(const float *__restrict__ X, const float *__restrict__ t0,
                                     const float *__restrict__ seed_tf, int64_t S, int node_dim,
                                     int pad_dim, int time_dim, float *__restrict__ R) {
  const int out = node_dim + pad_dim + time_dim;
  const int64_t total = S * out;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / out;
    const int c = int(i - s * out);
    float v = 0.f;
    if (c < node_dim) v = X[s * node_dim + c];
    else if (c >= node_dim + pad_dim) {
      const int j = c - node_dim - pad_dim;
      v = seed_tf ? seed_tf[s * time_dim + j] : __ldg(t0 + j);
    }
    R[i] = v;
  }
}

__global__ void cos_kernel(const float *__restrict__ b, int n, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cosf(b[i]);
}

// ---- the fused neighbour pass ------------------------------------------------------------------
// One CTA per seed.  z_n = [nbr_feat[s,n,:] | edge_feat[s,n,:] | cos(fma(float(tq - t_n), w, b))]
// is built once in shared memory (k x key_dim floats), logits_hn = (qk_h . z_n) * hd^-0.5 with
// masked slots set to -1e10 (attention.py:110-113), softmax over n, u_h = sum_n a_hn z_n.
// Padded slots are real inputs of the reference computation (their value rows are averaged when a
// seed has no valid neighbour at all), so they are assembled like any other slot; the caller
// passes the same rows the reference would gather (node row N-1 for id -1, zero edge features,
// time 0).
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
  const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Shared-memory row layout: z row n starts at z + n * zs + pad with pad = (4 - node_dim % 4) % 4 and
// zs = roundup4(pad + key), so the edge-feature part of every row is 16-byte aligned and is filled
// with 128-bit loads/stores.
struct AttnLayout {
  int pad, zs, a_len;
  size_t bytes;
};
static inline AttnLayout attn_layout(int k, int node_dim, int key, int H, bool qk_in_smem) {
  AttnLayout L;
  L.pad = (4 - node_dim % 4) % 4;
  L.zs = (L.pad + key + 3) & ~3;
  L.a_len = ((H + 1) / 2) * k * 2;  // probabilities of a head pair interleaved: [pair][n][2]
  L.bytes = (size_t(k) * L.zs + (qk_in_smem ? size_t(H) * key : 0) + L.a_len) * sizeof(float);
  return L;
}

// KT > 0: key <= 32 * KT and each lane keeps its KT columns of qk (two heads) in registers;
// KT == 0: any key, qk staged in shared memory.
template <int KT>
__global__ void __launch_bounds__(kAttnThreads, 8)
attn_neighbor_kernel(const float *__restrict__ nbr_feat, const float *__restrict__ edge_feat,
                     const int64_t *__restrict__ seed_t, const int64_t *__restrict__ nbr_t,
                     const int32_t *__restrict__ nbr_id, const float *__restrict__ tw,
                     const float *__restrict__ tb, const float *__restrict__ nbr_tf,
                     const float *__restrict__ QK, int64_t S, int k, int node_dim, int edge_dim,
                     int time_dim, int H, float scale, int vec_node, int vec_edge,
                     float *__restrict__ U, const int32_t *__restrict__ edge_rows) {
  extern __shared__ __align__(16) float smem[];
  const int key = node_dim + edge_dim + time_dim;
  const int pad = (4 - node_dim % 4) % 4, zs = (pad + key + 3) & ~3;
  float *z = smem + pad;                            // row n: z + n * zs, [key] floats
  float *qk = smem + k * zs;                        // [H][key]   (KT == 0 only)
  float *a = qk + (KT == 0 ? H * key : 0);          // [ceil(H/2)][k][2] logits -> probabilities
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kAttnThreads >> 5;
  const int td_full = time_dim & ~31, td_rem = time_dim - td_full;  // whole warp tiles + leftover
  const int feat = node_dim + edge_dim;
  // Time2Vec weights of this lane's columns (first four tiles) live in registers
  float wr[4], br[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = lane + 32 * t;
    wr[t] = c < td_full ? __ldg(tw + c) : 0.f;
    br[t] = c < td_full ? __ldg(tb + c) : 0.f;
  }
  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    const int64_t tq = nbr_tf ? 0 : seed_t[s];
    const float *nf = nbr_feat + s * int64_t(k) * node_dim;
    // edge features: the (S, k, edge_dim) block the sampler materialised, or -- edge_rows given --
    // row edge_rows[s, n] of the feature TABLE `edge_feat` (the store's edge_x; -1 = padding ->
    // zeros): the sampled rows are then read once, here, and never written anywhere
    const float *ef = edge_rows ? edge_feat : edge_feat + s * int64_t(k) * edge_dim;
    const float *qks = QK + s * int64_t(H) * key;
    // feature rows go global -> shared with 16-byte asynchronous copies (no register staging, all
    // of a warp's rows in flight at once); the cosines below overlap their latency.  One warp per
    // neighbour row, lanes along the feature dimension (coalesced); short runtime-bounded loops
    // are kept rolled (the unrolled forms cost ~25 instructions per copy in bounds logic).
    if (!vec_node) {  // narrow / unaligned node features (TGAT layer 1: one column): flat pass
#pragma unroll 1
      for (int i = tid; i < k * node_dim; i += kAttnThreads) {
        const int n = i / node_dim;
        z[n * zs + (i - n * node_dim)] = __ldg(nf + i);
      }
    }
#pragma unroll 1
    for (int n = warp; n < k; n += nwarp) {
      float *zn = z + n * zs;
      if (vec_node) {
        const float *src = nf + n * node_dim + 4 * lane;
        float *dst = zn + 4 * lane;
#pragma unroll 1
        for (int c = lane; c < (node_dim >> 2); c += 32, src += 128, dst += 128) cp_async16(dst, src);
      }
      const int64_t erow = edge_rows ? int64_t(__ldg(edge_rows + s * k + n)) : int64_t(n);
      if (erow < 0) {
#pragma unroll 1
        for (int c = lane; c < edge_dim; c += 32) zn[node_dim + c] = 0.f;
      } else if (vec_edge) {
        const float *src = ef + erow * edge_dim + 4 * lane;
        float *dst = zn + node_dim + 4 * lane;
#pragma unroll 1
        for (int c = lane; c < (edge_dim >> 2); c += 32, src += 128, dst += 128) cp_async16(dst, src);
      } else {
        const float *efr = ef + erow * edge_dim;
#pragma unroll 1
        for (int c = lane; c < edge_dim; c += 32) zn[node_dim + c] = __ldg(efr + c);
      }
    }
    cp_async_commit();
    // this lane's qk columns of the first head pair: requested now, used after the barrier
    float q0[KT > 0 ? KT : 1], q1[KT > 0 ? KT : 1];
    if (KT > 0) {
#pragma unroll
      for (int t = 0; t < KT; ++t) {
        const int j = lane + 32 * t;
        q0[t] = j < key ? __ldg(qks + j) : 0.f;
        q1[t] = (H > 1 && j < key) ? __ldg(qks + key + j) : 0.f;
      }
    }
    if (KT == 0)
      for (int i = tid; i < H * key; i += kAttnThreads) qk[i] = __ldg(qks + i);
    // leftover time columns (time_dim % 32) of all k rows as flat (row, column) pairs, so the
    // per-row pass below runs whole warp tiles only (100 columns: 63 warp-cosines, not 80)
    if (!nbr_tf && td_rem) {
      for (int p = tid; p < k * td_rem; p += kAttnThreads) {
        const int n = p / td_rem, c = td_full + (p - n * td_rem);
        const float dt = float(tq - nbr_t[s * k + n]);
        z[n * zs + feat + c] = t2v_cos(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c)));
      }
    }
    for (int n = warp; n < k; n += nwarp) {
      float *zn = z + n * zs;
      float *zt = zn + feat;
      if (nbr_tf) {  // caller-provided time features (the plain attention.py:58 signature)
        const float *tf = nbr_tf + (s * int64_t(k) + n) * time_dim;
        for (int c = lane; c < time_dim; c += 32) zt[c] = __ldg(tf + c);
      } else {
        const float dt = float(tq - nbr_t[s * k + n]);  // int64 difference, then .float() (:23)
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (32 * t < td_full) zt[lane + 32 * t] = t2v_cos(__fmaf_rn(dt, wr[t], br[t]));
        for (int c = 128 + lane; c < td_full; c += 32)
          zt[c] = t2v_cos(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c)));
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // logits: the same warp-per-row split, two heads share each z load
    for (int h0 = 0; h0 < H; h0 += 2) {
      const bool two = h0 + 1 < H;
      float *ap = a + (h0 >> 1) * k * 2;
      if (KT > 0) {
        if (h0 > 0) {
#pragma unroll
          for (int t = 0; t < KT; ++t) {
            const int j = lane + 32 * t;
            q0[t] = j < key ? __ldg(qks + h0 * key + j) : 0.f;
            q1[t] = (two && j < key) ? __ldg(qks + (h0 + 1) * key + j) : 0.f;
          }
        }
        for (int n = warp; n < k; n += nwarp) {
          const float *zn = z + n * zs;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int t = 0; t < KT; ++t) {
            const int j = lane + 32 * t;
            const float v = j < key ? zn[j] : 0.f;
            acc0 = fmaf(q0[t], v, acc0);
            acc1 = fmaf(q1[t], v, acc1);
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
          }
          if (lane == 0) {
            const bool valid = nbr_id[s * k + n] != TGM_PADDED_NODE_ID;
            ap[2 * n] = valid ? acc0 * scale : -1e10f;
            ap[2 * n + 1] = valid ? acc1 * scale : -1e10f;
          }
        }
      } else {
        const float *q0 = qk + h0 * key, *q1 = q0 + (two ? key : 0);
        for (int n = warp; n < k; n += nwarp) {
          const float *zn = z + n * zs;
          float acc0 = 0.f, acc1 = 0.f;
          for (int j = lane; j < key; j += 32) {
            const float v = zn[j];
            acc0 = fmaf(q0[j], v, acc0);
            acc1 = fmaf(q1[j], v, acc1);
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
          }
          if (lane == 0) {
            const bool valid = nbr_id[s * k + n] != TGM_PADDED_NODE_ID;
            ap[2 * n] = valid ? acc0 * scale : -1e10f;
            ap[2 * n + 1] = valid ? acc1 * scale : -1e10f;
          }
        }
      }
    }
    __syncthreads();
    // softmax over the k slots of each head: one warp per head
    for (int h = warp; h < H; h += nwarp) {
      float *ah = a + (h >> 1) * k * 2 + (h & 1);  // stride 2
      float m = -INFINITY;
      for (int n = lane; n < k; n += 32) m = fmaxf(m, ah[2 * n]);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int n = lane; n < k; n += 32) {
        const float e = expf(ah[2 * n] - m);
        ah[2 * n] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int n = lane; n < k; n += 32) ah[2 * n] *= inv;
    }
    __syncthreads();
    // u_h[j] = sum_n a_hn z_n[j]: one thread per column j, two heads per pass (one 64-bit
    // broadcast load brings both probabilities)
    float *u = U + s * int64_t(H) * key;
    for (int h0 = 0; h0 < H; h0 += 2) {
      const bool two = h0 + 1 < H;
      const float2 *ap = reinterpret_cast<const float2 *>(a + (h0 >> 1) * k * 2);
      for (int j = tid; j < key; j += kAttnThreads) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
        for (int n = 0; n < k; ++n) {
          const float v = z[n * zs + j];
          const float2 av = ap[n];
          acc0 = fmaf(av.x, v, acc0);
          acc1 = fmaf(av.y, v, acc1);
        }
        u[h0 * key + j] = acc0;
        if (two) u[(h0 + 1) * key + j] = acc1;
      }
    }
    __syncthreads();
  }
}

template <int KT>
int launch_attn_neighbor(const tgm_attn *a, const float *nbr_node_feat, const float *edge_feat,
                         const int64_t *seed_t, const int64_t *nbr_t, const int32_t *nbr_id,
                         const float *nbr_tf, int64_t S, int k, cudaStream_t st,
                         const int32_t *edge_rows) {
  const AttnLayout L = attn_layout(k, a->node_dim, a->key, a->H, KT == 0);
  TGM_REQUIRE(L.bytes <= 200 * 1024, "tgm_attn_forward: k * key_dim too large for shared memory");
  if (L.bytes > 48 * 1024)
    TGM_CUDA(cudaFuncSetAttribute(attn_neighbor_kernel<KT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.bytes)));
  const int ctas_per_sm =
      int(std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (L.bytes + 1024))));
  const int vec_node = a->node_dim % 4 == 0 && aligned16(nbr_node_feat);
  const int vec_edge = a->edge_dim % 4 == 0 && aligned16(edge_feat);
  attn_neighbor_kernel<KT><<<grid_for(S, 1, ctas_per_sm), kAttnThreads, L.bytes, st>>>(
      nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, a->tw, a->tb, nbr_tf, a->QK, S, k,
      a->node_dim, a->edge_dim, a->time_dim, a->H, 1.0f / sqrtf(float(a->hd)), vec_node, vec_edge,
      a->U, edge_rows);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

// out = LayerNorm(Y + b_O + R) (attention.py:124-127); one warp per row, two-pass variance
__global__ void __launch_bounds__(256)
attn_epilogue_kernel(const float *__restrict__ Y, const float *__restrict__ bo,
                     const float *__restrict__ R, const float *__restrict__ lnw,
                     const float *__restrict__ lnb, int64_t S, int out, float eps,
                     float *__restrict__ dst) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t s = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); s < S;
       s += int64_t(gridDim.x) * wpb) {
    const float *y = Y + s * out, *r = R + s * out;
    float sum = 0.f;
    for (int c = lane; c < out; c += 32) sum += y[c] + __ldg(bo + c) + r[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / float(out);
    float var = 0.f;
    for (int c = lane; c < out; c += 32) {
      const float d = y[c] + __ldg(bo + c) + r[c] - mean;
      var = fmaf(d, d, var);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / float(out) + eps);
    for (int c = lane; c < out; c += 32) {
      const float v = y[c] + __ldg(bo + c) + r[c];
      dst[s * out + c] = (v - mean) * rstd * __ldg(lnw + c) + __ldg(lnb + c);
    }
  }
}

// ---- MergeLayer pieces (tgat.py:34-38) ----------------------------------------------------------
// out[s] = [a[s] | b[s] | 0 ...] with row pitch d >= da + db
__global__ void concat2_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t S,
                               int da, int db, int d, float *__restrict__ out) {
  const int64_t total = S * d;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t s = i / d;
    const int c = int(i - s * d);
    out[i] = c < da ? a[s * da + c] : c < da + db ? b[s * db + (c - da)] : 0.f;
  }
}
__global__ void bias_act_kernel(float *__restrict__ x, const float *__restrict__ b, int64_t S,
                                int d, int relu) {
  const int64_t total = S * d;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    float v = x[i] + __ldg(b + int(i % d));
    x[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// Row gather with torch negative-index semantics: out[i,:] = table[ids[i] < 0 ? ids[i]+N : ids[i]]
__global__ void gather_rows_kernel(const float *__restrict__ table, int64_t N, int dim,
                                   const int32_t *__restrict__ ids, int64_t n,
                                   float *__restrict__ out) {
  const int64_t total = n * dim;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int c = int(i - r * dim);
    int64_t v = ids[r];
    if (v < 0) v += N;
    out[i] = (v >= 0 && v < N) ? __ldg(table + v * dim + c) : 0.f;
  }
}

int dev_copy(float **dst, const float *src, size_t n) {
  TGM_CUDA(cudaMalloc(dst, (n ? n : 1) * sizeof(float)));
  if (n) TGM_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyDefault));
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_attn_create(tgm_attn **out, int32_t n_heads, int32_t node_dim, int32_t edge_dim,
                               int32_t time_dim, const float *W_Q, const float *W_KV,
                               const float *W_O, const float *b_O, const float *ln_w,
                               const float *ln_b, float ln_eps, const float *t2v_w,
                               const float *t2v_b, int device) {
  TGM_REQUIRE(out != nullptr, "tgm_attn_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(n_heads > 0 && node_dim > 0 && edge_dim > 0 && time_dim > 0,
              "tgm_attn_create: n_heads,node_dim,edge_dim,time_dim must be > 0");  // attention.py:38-39
  TGM_REQUIRE(W_Q && W_KV && W_O && b_O && ln_w && ln_b && t2v_w && t2v_b,
              "tgm_attn_create: NULL parameter");
  TGM_REQUIRE(device >= 0, "tgm_attn_create: a CUDA device is required (no CPU fallback)");
  DeviceGuard g(device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_attn_create: cannot select device");
  tgm_attn *a = new (std::nothrow) tgm_attn();
  if (!a) return fail(TGM_ERR_OOM, "tgm_attn_create: host allocation failed");
  a->device = device;
  a->H = n_heads, a->node_dim = node_dim, a->edge_dim = edge_dim, a->time_dim = time_dim;
  int od = node_dim + time_dim;  // attention.py:41-45
  if (od % n_heads) a->pad_dim = n_heads - od % n_heads;
  od += a->pad_dim;
  a->out_dim = od, a->hd = od / n_heads, a->key = node_dim + edge_dim + time_dim;
  a->eps = ln_eps;
  const size_t o = size_t(od), kd = size_t(a->key), td = size_t(time_dim);
  int rc = dev_copy(&a->Wq, W_Q, o * o);
  if (!rc) rc = dev_copy(&a->Wkv, W_KV, 2 * o * kd);
  if (!rc) rc = dev_copy(&a->Wo, W_O, o * o);
  if (!rc) rc = dev_copy(&a->bo, b_O, o);
  if (!rc) rc = dev_copy(&a->lnw, ln_w, o);
  if (!rc) rc = dev_copy(&a->lnb, ln_b, o);
  if (!rc) rc = dev_copy(&a->tw, t2v_w, td);
  if (!rc) rc = dev_copy(&a->tb, t2v_b, td);
  if (!rc) rc = dev_copy(&a->t0, t2v_b, td);
  if (!rc) {
    cos_kernel<<<(time_dim + 127) / 128, 128>>>(a->tb, time_dim, a->t0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "Time2Vec(0)", __FILE__, __LINE__);
  }
  if (!rc) {
    cublasStatus_t s = cublasCreate(&a->blas);
    if (s != CUBLAS_STATUS_SUCCESS) rc = blas_fail(s, "cublasCreate");
    else cublasSetMathMode(a->blas, CUBLAS_PEDANTIC_MATH);  // true fp32, no TF32 down-conversion
  }
  if (!rc) rc = attn_fold_alloc(a);
  if (!rc) rc = attn_fold_refresh(a, nullptr);
  if (!rc) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "weight folding", __FILE__, __LINE__);
  }
  if (rc) {
    delete a;
    return rc;
  }
  *out = a;
  return TGM_OK;
}

extern "C" int tgm_attn_set_params(tgm_attn *a, const float *W_Q, const float *W_KV,
                                   const float *W_O, const float *b_O, const float *ln_w,
                                   const float *ln_b, const float *t2v_w, const float *t2v_b,
                                   tgm_stream stream) {
  TGM_REQUIRE(a != nullptr, "tgm_attn_set_params: handle is NULL");
  TGM_REQUIRE(W_Q && W_KV && W_O && b_O && ln_w && ln_b && t2v_w && t2v_b,
              "tgm_attn_set_params: NULL parameter");
  DeviceGuard g(a->device);
  cudaStream_t st = as_stream(stream);
  const size_t o = size_t(a->out_dim), kd = size_t(a->key), td = size_t(a->time_dim);
  TGM_CUDA(cudaMemcpyAsync(a->Wq, W_Q, o * o * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->Wkv, W_KV, 2 * o * kd * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->Wo, W_O, o * o * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->bo, b_O, o * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->lnw, ln_w, o * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->lnb, ln_b, o * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->tw, t2v_w, td * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(a->tb, t2v_b, td * 4, cudaMemcpyDefault, st));
  cos_kernel<<<(a->time_dim + 127) / 128, 128, 0, st>>>(a->tb, a->time_dim, a->t0);
  TGM_LAUNCH_CHECK();
  return attn_fold_refresh(a, st);
}

extern "C" void tgm_attn_destroy(tgm_attn *a) { delete a; }

extern "C" int tgm_attn_out_dim(const tgm_attn *a) { return a ? a->out_dim : TGM_ERR_INVALID; }

int attn_workspace(tgm_attn *a, int64_t S, cudaStream_t st) {
  if (S <= a->cap) return TGM_OK;
  const int od = a->out_dim, key = a->key, H = a->H;
  TGM_CUDA(cudaStreamSynchronize(st));
  for (float **p : {&a->R, &a->Q, &a->QK, &a->U, &a->O, &a->Y}) {
    cudaFree(*p);
    *p = nullptr;
  }
  a->cap = 0;
  const size_t rows = size_t(S + S / 4);
  TGM_CUDA(cudaMalloc(&a->R, rows * od * 4));
  TGM_CUDA(cudaMalloc(&a->Q, rows * od * 4));
  TGM_CUDA(cudaMalloc(&a->QK, rows * H * key * 4));
  TGM_CUDA(cudaMalloc(&a->U, rows * a->Kp * 4));  // Kp >= H key, Np >= od: both chains fit
  TGM_CUDA(cudaMalloc(&a->O, rows * od * 4));
  TGM_CUDA(cudaMalloc(&a->Y, rows * a->Np * 4));
  a->cap = int64_t(rows);
  return TGM_OK;
}

int attn_forward_impl(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                             const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                             const float *seed_tf, const float *nbr_tf, const int32_t *nbr_id,
                             int64_t S, int32_t k, float *out, tgm_stream stream,
                             const int32_t *edge_rows, bool keep_intermediates) {
  TGM_REQUIRE(a != nullptr, "tgm_attn_forward: handle is NULL");
  TGM_REQUIRE(S >= 0 && k >= 1, "tgm_attn_forward: bad sizes");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(node_x && nbr_node_feat && edge_feat && nbr_id && out,
              "tgm_attn_forward: NULL array argument");
  TGM_REQUIRE((seed_t && nbr_t) || (seed_tf && nbr_tf),
              "tgm_attn_forward: need either timestamps or time features");
  TGM_REQUIRE(S < (int64_t(1) << 31), "tgm_attn_forward: S must be < 2^31");
  DeviceGuard g(a->device);
  cudaStream_t st = as_stream(stream);
  const int od = a->out_dim, key = a->key, H = a->H, hd = a->hd;
  if (int rc = attn_workspace(a, S, st)) return rc;
  if (!keep_intermediates && !seed_tf && attn_folded_covers(a, k))
    return attn_forward_folded(a, node_x, nbr_node_feat,
                               single_hop(edge_feat, edge_rows, seed_t, nbr_t, nbr_id, S), S, k,
                               LnTarget{out, a->out_dim, nullptr, 0}, st);
  TGM_BLAS(cublasSetStream(a->blas, st));
  const float one = 1.f, zero = 0.f;
  // R = [X | pad | Time2Vec(0)],  Q = R W_Q^T
  attn_residual_kernel<<<grid_for(S * od, 256, 8), 256, 0, st>>>(
      node_x, a->t0, seed_tf, S, a->node_dim, a->pad_dim, a->time_dim, a->R);
  TGM_LAUNCH_CHECK();
  TGM_BLAS(gemm_nt(a->blas, S, od, od, a->R, od, a->Wq, a->Q, od));
  // qk[s,h,:] = W_K,h^T Q[s,h,:]   (W_K = rows [0,out) of W_KV)
  TGM_BLAS(cublasSgemmStridedBatched(a->blas, CUBLAS_OP_N, CUBLAS_OP_N, key, int(S), hd, &one,
                                     a->Wkv, key, int64_t(hd) * key, a->Q, od, hd, &zero, a->QK,
                                     H * key, key, H));
  // fused gather + Time2Vec + masked softmax + weighted sum
  {
    const int kt = (key + 31) / 32;
    int rc;
    if (kt <= 4) rc = launch_attn_neighbor<4>(a, nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, nbr_tf, S, k, st, edge_rows);
    else if (kt <= 9) rc = launch_attn_neighbor<9>(a, nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, nbr_tf, S, k, st, edge_rows);
    else if (kt <= 14) rc = launch_attn_neighbor<14>(a, nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, nbr_tf, S, k, st, edge_rows);
    else rc = launch_attn_neighbor<0>(a, nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, nbr_tf, S, k, st, edge_rows);
    if (rc) return rc;
  }
  // O[s,h,:] = W_V,h u[s,h,:]   (W_V = rows [out, 2 out) of W_KV)
  TGM_BLAS(cublasSgemmStridedBatched(a->blas, CUBLAS_OP_T, CUBLAS_OP_N, hd, int(S), key, &one,
                                     a->Wkv + size_t(od) * key, key, int64_t(hd) * key, a->U,
                                     H * key, key, &zero, a->O, od, hd, H));
  // Y = O W_O^T ; out = LayerNorm(Y + b_O + R)
  TGM_BLAS(gemm_nt(a->blas, S, od, od, a->O, od, a->Wo, a->Y, od));
  attn_epilogue_kernel<<<grid_for(S, 8, 8), 256, 0, st>>>(a->Y, a->bo, a->R, a->lnw, a->lnb, S, od,
                                                          a->eps, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_attn_forward(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                                const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                                const int32_t *nbr_id, int64_t S, int32_t k, float *out,
                                tgm_stream stream) {
  TGM_REQUIRE(seed_t && nbr_t, "tgm_attn_forward: NULL array argument");
  return attn_forward_impl(a, node_x, nbr_node_feat, edge_feat, seed_t, nbr_t, nullptr, nullptr,
                           nbr_id, S, k, out, stream);
}

extern "C" int tgm_attn_forward_rows(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                                     const float *edge_table, const int32_t *edge_rows,
                                     const int64_t *seed_t, const int64_t *nbr_t,
                                     const int32_t *nbr_id, int64_t S, int32_t k, float *out,
                                     tgm_stream stream) {
  TGM_REQUIRE(seed_t && nbr_t && edge_rows, "tgm_attn_forward_rows: NULL array argument");
  return attn_forward_impl(a, node_x, nbr_node_feat, edge_table, seed_t, nbr_t, nullptr, nullptr,
                           nbr_id, S, k, out, stream, edge_rows);
}

extern "C" int tgm_attn_folded_covers(const tgm_attn *a, int32_t k) {
  return a ? int(attn_folded_covers(a, k)) : TGM_ERR_INVALID;
}

extern "C" int tgm_attn_forward_segments(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                                         const float *const *edge_feat_segs, const int64_t *seg_rows,
                                         int32_t n_segs, const int64_t *seed_t, const int64_t *nbr_t,
                                         const int32_t *nbr_id, int64_t S, int32_t k, float *out,
                                         tgm_stream stream) {
  TGM_REQUIRE(a != nullptr, "tgm_attn_forward_segments: handle is NULL");
  TGM_REQUIRE(S >= 0 && k >= 1, "tgm_attn_forward_segments: bad sizes");
  TGM_REQUIRE(n_segs >= 1 && n_segs <= 4 && edge_feat_segs && seg_rows,
              "tgm_attn_forward_segments: 1..4 edge-feature segments");
  int64_t total = 0;
  for (int i = 0; i < n_segs; ++i) {
    TGM_REQUIRE(seg_rows[i] >= 0 && (edge_feat_segs[i] || seg_rows[i] == 0),
                "tgm_attn_forward_segments: bad segment");
    total += seg_rows[i];
  }
  TGM_REQUIRE(total == S, "tgm_attn_forward_segments: segment rows must add up to S");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(node_x && nbr_node_feat && seed_t && nbr_t && nbr_id && out,
              "tgm_attn_forward_segments: NULL array argument");
  TGM_REQUIRE(S < (int64_t(1) << 31), "tgm_attn_forward_segments: S must be < 2^31");
  TGM_REQUIRE(attn_folded_covers(a, k),
              "tgm_attn_forward_segments: shape outside the folded chain (see tgm_attn_folded_covers)");
  DeviceGuard g(a->device);
  cudaStream_t st = as_stream(stream);
  if (int rc = attn_workspace(a, S, st)) return rc;
  HopSegs hops{};
  hops.n = n_segs;
  int64_t first = 0;
  for (int i = 0; i < 4; ++i) {
    if (i < n_segs) {
      hops.nid[i] = nbr_id + first * k, hops.nt[i] = nbr_t + first * k, hops.st[i] = seed_t + first;
      hops.ef[i] = edge_feat_segs[i];
      first += seg_rows[i];
    }
    hops.end[i] = first;
  }
  return attn_forward_folded(a, node_x, nbr_node_feat, hops, S, k,
                             LnTarget{out, a->out_dim, nullptr, 0}, st);
}

extern "C" int tgm_attn_forward_feats(tgm_attn *a, const float *node_x, const float *time_feat,
                                      const float *edge_feat, const float *nbr_node_feat,
                                      const float *nbr_time_feat, const int32_t *nbr_id, int64_t S,
                                      int32_t k, float *out, tgm_stream stream) {
  TGM_REQUIRE(time_feat && nbr_time_feat, "tgm_attn_forward_feats: NULL array argument");
  return attn_forward_impl(a, node_x, nbr_node_feat, edge_feat, nullptr, nullptr, time_feat,
                           nbr_time_feat, nbr_id, S, k, out, stream);
}

extern "C" int tgm_mlp2_create(tgm_mlp2 **out, int32_t in1, int32_t in2, int32_t hidden,
                               int32_t out_dim, const float *W1, const float *b1, const float *W2,
                               const float *b2, int device) {
  TGM_REQUIRE(out != nullptr, "tgm_mlp2_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(in1 > 0 && in2 >= 0 && hidden > 0 && out_dim > 0, "tgm_mlp2_create: bad sizes");
  TGM_REQUIRE(W1 && b1 && W2 && b2, "tgm_mlp2_create: NULL parameter");
  TGM_REQUIRE(device >= 0, "tgm_mlp2_create: a CUDA device is required (no CPU fallback)");
  DeviceGuard g(device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_mlp2_create: cannot select device");
  tgm_mlp2 *m = new (std::nothrow) tgm_mlp2();
  if (!m) return fail(TGM_ERR_OOM, "tgm_mlp2_create: host allocation failed");
  m->device = device, m->in1 = in1, m->in2 = in2, m->hidden = hidden, m->out = out_dim;
  m->inp = (in1 + in2 + 3) & ~3;
  int rc = TGM_OK;
  {
    cudaError_t e = cudaMalloc(&m->W1, size_t(hidden) * m->inp * 4);
    if (e == cudaSuccess) e = cudaMemset(m->W1, 0, size_t(hidden) * m->inp * 4);
    if (e == cudaSuccess)
      e = cudaMemcpy2D(m->W1, size_t(m->inp) * 4, W1, size_t(in1 + in2) * 4, size_t(in1 + in2) * 4,
                       hidden, cudaMemcpyDefault);
    if (e != cudaSuccess) rc = cuda_fail(e, "tgm_mlp2_create: W1", __FILE__, __LINE__);
  }
  if (!rc) rc = dev_copy(&m->b1, b1, hidden);
  if (!rc) rc = dev_copy(&m->W2, W2, size_t(out_dim) * hidden);
  if (!rc) rc = dev_copy(&m->b2, b2, out_dim);
  if (!rc) {
    cublasStatus_t s = cublasCreate(&m->blas);
    if (s != CUBLAS_STATUS_SUCCESS) rc = blas_fail(s, "cublasCreate");
    else cublasSetMathMode(m->blas, CUBLAS_PEDANTIC_MATH);
  }
  if (rc) {
    delete m;
    return rc;
  }
  *out = m;
  return TGM_OK;
}

extern "C" void tgm_mlp2_destroy(tgm_mlp2 *m) { delete m; }

int mlp2_workspace(tgm_mlp2 *m, int64_t S, cudaStream_t st) {
  if (S <= m->cap) return TGM_OK;
  TGM_CUDA(cudaStreamSynchronize(st));
  cudaFree(m->cat), cudaFree(m->h);
  m->cat = m->h = nullptr;
  m->cap = 0;
  const size_t rows = size_t(S + S / 4);
  TGM_CUDA(cudaMalloc(&m->cat, rows * m->inp * 4));
  TGM_CUDA(cudaMalloc(&m->h, rows * m->hidden * 4));
  m->cap = int64_t(rows);
  return TGM_OK;
}

int mlp2_forward_cat(tgm_mlp2 *m, int64_t S, float *out, cudaStream_t st) {
  // bias and ReLU ride in the product's epilogue (tensor-core kernel) or one pass after it (cuBLAS)
  if (int rc = dense_linear(m->blas, S, m->hidden, m->inp, m->cat, m->W1, m->b1, 2, m->h, st)) return rc;
  return dense_linear(m->blas, S, m->out, m->hidden, m->h, m->W2, m->b2, 0, out, st);
}

extern "C" int tgm_mlp2_forward(tgm_mlp2 *m, const float *x1, const float *x2, int64_t S,
                                float *out, tgm_stream stream) {
  TGM_REQUIRE(m != nullptr, "tgm_mlp2_forward: handle is NULL");
  TGM_REQUIRE(S >= 0, "tgm_mlp2_forward: S must be >= 0");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(x1 && (x2 || m->in2 == 0) && out, "tgm_mlp2_forward: NULL array argument");
  DeviceGuard g(m->device);
  cudaStream_t st = as_stream(stream);
  if (int rc = mlp2_workspace(m, S, st)) return rc;
  concat2_kernel<<<grid_for(S * m->inp, 256, 8), 256, 0, st>>>(x1, x2, S, m->in1, m->in2, m->inp,
                                                             m->cat);
  TGM_LAUNCH_CHECK();
  return mlp2_forward_cat(m, S, out, st);
}

extern "C" int tgm_gather_rows(const float *table, int64_t num_rows, int32_t dim,
                               const int32_t *ids, int64_t n, float *out, tgm_stream stream) {
  TGM_REQUIRE(num_rows >= 0 && dim >= 1 && n >= 0, "tgm_gather_rows: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(table && ids && out, "tgm_gather_rows: NULL array argument");
  gather_rows_kernel<<<grid_for(n * dim, 256, 8), 256, 0, as_stream(stream)>>>(table, num_rows, dim,
                                                                               ids, n, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
