// Backward of the time-encoded single-query attention (training through TGAT's aggregation).
//
// The reference trains through TemporalAttention with autograd (tgm/nn/modules/attention.py:58-128
// inside examples/linkproppred/tgat.py:83-86).  Here the backward follows the same reassociated
// dataflow as the forward (attention.cu): with z_n = [nbr_feat | edge_feat | Time2Vec(dt_n)],
//   l_hn = scale * (qk_h . z_n)  (masked slots: the constant -1e10, no gradient),
//   a_h = softmax_n(l_h),  u_h = sum_n a_hn z_n,
// the per-seed part is  da_hn = dU_h . z_n,  dl_hn = a_hn (da_hn - sum_m a_hm da_hm),
//   dqk_h = scale * sum_{valid n} dl_hn z_n,   dz_n = sum_h a_hn dU_h + [valid] scale dl_hn qk_h,
// and dz splits into the neighbour-feature gradient, the (optional) edge-feature gradient and the
// Time2Vec parameter gradients (-sin(arg) dz).  Everything around it is LayerNorm backward plus
// plain SGEMMs (cuBLAS, true fp32).  The forward intermediates are recomputed into the handle's
// workspace instead of being saved, so the ABI needs no opaque "saved state".
#include <algorithm>
#include <cmath>

#include "attention.cuh"

using namespace tgm;

namespace {

int blas_fail_b(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define BWD_BLAS(expr)                                             \
  do {                                                             \
    cublasStatus_t _s = (expr);                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) return blas_fail_b(_s, #expr); \
  } while (0)

// row-major C[M,N] = alpha * op(A)[M,K] . op(B)[K,N] + beta * C
cublasStatus_t gemm_rm(cublasHandle_t h, bool ta, bool tb, int64_t M, int64_t N, int64_t K,
                       const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                       float beta) {
  const float one = 1.f;
  return cublasSgemm(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, int(N),
                     int(M), int(K), &one, B, ldb, A, lda, &beta, C, ldc);
}

// LayerNorm backward per row (attention.py:127): v = Y + b_O + R, out = xhat * gamma + beta.
// Writes dV (= dY = residual gradient) and accumulates dgamma, dbeta, db_O.
__global__ void __launch_bounds__(256)
attn_ln_bwd_kernel(const float *__restrict__ Y, const float *__restrict__ bo,
                   const float *__restrict__ R, const float *__restrict__ lnw,
                   const float *__restrict__ dOut, int64_t S, int out, float eps,
                   float *__restrict__ dV, float *__restrict__ dlnw, float *__restrict__ dlnb,
                   float *__restrict__ dbo) {
  extern __shared__ float s_acc[];  // [3][out] per-CTA partial sums of dgamma, dbeta, db_O
  for (int i = threadIdx.x; i < 3 * out; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t s = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); s < S;
       s += int64_t(gridDim.x) * wpb) {
    const float *y = Y + s * out, *r = R + s * out, *g = dOut + s * out;
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
    float m1 = 0.f, m2 = 0.f;  // mean(g*gamma), mean(g*gamma*xhat)
    for (int c = lane; c < out; c += 32) {
      const float xhat = (y[c] + __ldg(bo + c) + r[c] - mean) * rstd;
      const float gg = g[c] * __ldg(lnw + c);
      m1 += gg;
      m2 = fmaf(gg, xhat, m2);
      atomicAdd(&s_acc[c], g[c] * xhat);
      atomicAdd(&s_acc[out + c], g[c]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 /= float(out), m2 /= float(out);
    for (int c = lane; c < out; c += 32) {
      const float xhat = (y[c] + __ldg(bo + c) + r[c] - mean) * rstd;
      const float dv = rstd * (g[c] * __ldg(lnw + c) - m1 - xhat * m2);
      dV[s * out + c] = dv;
      atomicAdd(&s_acc[2 * out + c], dv);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < out; c += blockDim.x) {
    atomicAdd(dlnw + c, s_acc[c]);
    atomicAdd(dlnb + c, s_acc[out + c]);
    atomicAdd(dbo + c, s_acc[2 * out + c]);
  }
}

constexpr int kBwdThreads = 128;

// Per-seed backward of the neighbour pass (one CTA per seed, same smem staging as the forward).
__global__ void __launch_bounds__(kBwdThreads)
attn_neighbor_bwd_kernel(const float *__restrict__ nbr_feat, const float *__restrict__ edge_feat,
                         const int64_t *__restrict__ seed_t, const int64_t *__restrict__ nbr_t,
                         const int32_t *__restrict__ nbr_id, const float *__restrict__ tw,
                         const float *__restrict__ tb, const float *__restrict__ QK,
                         const float *__restrict__ dU, int64_t S, int k, int node_dim,
                         int edge_dim, int time_dim, int H, float scale,
                         float *__restrict__ dQK, float *__restrict__ d_nbr_feat,
                         float *__restrict__ d_edge_feat, float *__restrict__ dtw,
                         float *__restrict__ dtb) {
  extern __shared__ float smem[];
  const int key = node_dim + edge_dim + time_dim;
  float *z = smem;               // [k][key]
  float *qk = z + k * key;       // [H][key]
  float *du = qk + H * key;      // [H][key]
  float *a = du + H * key;       // [H][k] probabilities
  float *dl = a + H * k;         // [H][k] da, then dl
  float *sw = dl + H * k;        // [time_dim] partial dw
  float *sb = sw + time_dim;     // [time_dim] partial db
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kBwdThreads >> 5;
  for (int i = tid; i < 2 * time_dim; i += kBwdThreads) sw[i] = 0.f;
  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    const int64_t tq = seed_t[s];
    const float *nf = nbr_feat + s * int64_t(k) * node_dim;
    const float *ef = edge_feat + s * int64_t(k) * edge_dim;
    for (int n = warp; n < k; n += nwarp) {
      float *zn = z + n * key;
      for (int c = lane; c < node_dim; c += 32) zn[c] = __ldg(nf + n * node_dim + c);
      for (int c = lane; c < edge_dim; c += 32) zn[node_dim + c] = __ldg(ef + n * edge_dim + c);
      const float dt = float(tq - nbr_t[s * k + n]);
      for (int c = lane; c < time_dim; c += 32)
        zn[node_dim + edge_dim + c] = cosf(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c)));
    }
    for (int i = tid; i < H * key; i += kBwdThreads) {
      qk[i] = QK[s * int64_t(H) * key + i];
      du[i] = dU[s * int64_t(H) * key + i];
    }
    __syncthreads();
    // logits and da_hn = dU_h . z_n : one warp per (h, n)
    for (int p = warp; p < H * k; p += nwarp) {
      const int h = p / k, n = p - h * k;
      const float *zz = z + n * key, *qq = qk + h * key, *dd = du + h * key;
      float acc = 0.f, acc2 = 0.f;
      for (int j = lane; j < key; j += 32) {
        acc = fmaf(qq[j], zz[j], acc);
        acc2 = fmaf(dd[j], zz[j], acc2);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
      }
      if (lane == 0) {
        a[p] = nbr_id[s * k + n] != TGM_PADDED_NODE_ID ? acc * scale : -1e10f;
        dl[p] = acc2;
      }
    }
    __syncthreads();
    // softmax and its backward: dl_hn = a_hn (da_hn - sum_m a_hm da_hm); masked slots get 0
    for (int h = warp; h < H; h += nwarp) {
      float m = -INFINITY;
      for (int n = lane; n < k; n += 32) m = fmaxf(m, a[h * k + n]);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int n = lane; n < k; n += 32) {
        const float e = expf(a[h * k + n] - m);
        a[h * k + n] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      float dot = 0.f;
      for (int n = lane; n < k; n += 32) {
        const float p = a[h * k + n] * inv;
        a[h * k + n] = p;
        dot = fmaf(p, dl[h * k + n], dot);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      for (int n = lane; n < k; n += 32) {
        const bool valid = nbr_id[s * k + n] != TGM_PADDED_NODE_ID;
        dl[h * k + n] = valid ? a[h * k + n] * (dl[h * k + n] - dot) * scale : 0.f;
      }
    }
    __syncthreads();
    // dqk_h[j] = sum_n dl_hn z_n[j]   (dl already carries the scale)
    for (int i = tid; i < H * key; i += kBwdThreads) {
      const int h = i / key, j = i - h * key;
      float acc = 0.f;
      for (int n = 0; n < k; ++n) acc = fmaf(dl[h * k + n], z[n * key + j], acc);
      dQK[s * int64_t(H) * key + i] = acc;
    }
    // dz_n[j] = sum_h a_hn dU_h[j] + dl_hn qk_h[j], routed to its three destinations
    for (int n = warp; n < k; n += nwarp) {
      const float dt = float(tq - nbr_t[s * k + n]);
      for (int j = lane; j < key; j += 32) {
        float dz = 0.f;
        for (int h = 0; h < H; ++h)
          dz = fmaf(a[h * k + n], du[h * key + j], fmaf(dl[h * k + n], qk[h * key + j], dz));
        if (j < node_dim) {
          if (d_nbr_feat) d_nbr_feat[(s * k + n) * int64_t(node_dim) + j] = dz;
        } else if (j < node_dim + edge_dim) {
          if (d_edge_feat) d_edge_feat[(s * k + n) * int64_t(edge_dim) + (j - node_dim)] = dz;
        } else {
          const int c = j - node_dim - edge_dim;
          const float darg = -sinf(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c))) * dz;
          atomicAdd(&sw[c], dt * darg);
          atomicAdd(&sb[c], darg);
        }
      }
    }
    __syncthreads();
  }
  for (int c = tid; c < time_dim; c += kBwdThreads) {
    atomicAdd(dtw + c, sw[c]);
    atomicAdd(dtb + c, sb[c]);
  }
}

// dX = dR[:, :node_dim];  Time2Vec(0) = cos(b): db_c += -sin(b_c) * sum_s dR[s, off + c]
// (column sums reduced per CTA in shared memory, then one atomic per column and CTA)
__global__ void __launch_bounds__(256)
attn_residual_bwd_kernel(const float *__restrict__ dR, const float *__restrict__ tb, int64_t S,
                         int node_dim, int pad_dim, int time_dim, float *__restrict__ dX,
                         float *__restrict__ dtb) {
  extern __shared__ float s_col[];  // [time_dim]
  for (int c = threadIdx.x; c < time_dim; c += blockDim.x) s_col[c] = 0.f;
  __syncthreads();
  const int out = node_dim + pad_dim + time_dim, off = node_dim + pad_dim;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t s = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); s < S;
       s += int64_t(gridDim.x) * wpb) {
    const float *g = dR + s * out;
    if (dX)
      for (int c = lane; c < node_dim; c += 32) dX[s * node_dim + c] = g[c];
    for (int c = lane; c < time_dim; c += 32) atomicAdd(&s_col[c], g[off + c]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < time_dim; c += blockDim.x)
    atomicAdd(dtb + c, -sinf(__ldg(tb + c)) * s_col[c]);
}

}  // namespace

extern "C" int tgm_attn_backward(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                                 const float *edge_feat, const int64_t *seed_t,
                                 const int64_t *nbr_t, const int32_t *nbr_id, int64_t S, int32_t k,
                                 const float *d_out, float *d_node_x, float *d_nbr_node_feat,
                                 float *d_edge_feat, float *dW_Q, float *dW_KV, float *dW_O,
                                 float *db_O, float *dln_w, float *dln_b, float *dt2v_w,
                                 float *dt2v_b, tgm_stream stream) {
  TGM_REQUIRE(a != nullptr, "tgm_attn_backward: handle is NULL");
  TGM_REQUIRE(S >= 0 && k >= 1, "tgm_attn_backward: bad sizes");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(d_out && dW_Q && dW_KV && dW_O && db_O && dln_w && dln_b && dt2v_w && dt2v_b,
              "tgm_attn_backward: NULL gradient argument");
  DeviceGuard g(a->device);
  cudaStream_t st = as_stream(stream);
  const int od = a->out_dim, key = a->key, H = a->H, hd = a->hd;
  if (S > a->bcap) {
    TGM_CUDA(cudaStreamSynchronize(st));
    for (float **p : {&a->dV, &a->dO, &a->dU, &a->dQK, &a->dQ, &a->dR, &a->fwd_out}) {
      cudaFree(*p);
      *p = nullptr;
    }
    a->bcap = 0;
    const size_t rows = size_t(S + S / 4);
    for (float **p : {&a->dV, &a->dO, &a->dQ, &a->dR, &a->fwd_out})
      TGM_CUDA(cudaMalloc(p, rows * od * 4));
    TGM_CUDA(cudaMalloc(&a->dU, rows * H * key * 4));
    TGM_CUDA(cudaMalloc(&a->dQK, rows * H * key * 4));
    a->bcap = int64_t(rows);
  }
  // recompute the forward intermediates (R, Q, QK, U, O, Y) into the workspace
  int rc = attn_forward_impl(a, node_x, nbr_node_feat, edge_feat, seed_t, nbr_t, nullptr, nullptr,
                             nbr_id, S, k, a->fwd_out, stream, nullptr, /*keep_intermediates=*/true);
  if (rc != TGM_OK) return rc;
  BWD_BLAS(cublasSetStream(a->blas, st));
  // LayerNorm backward: dV = d(Y + b_O + R)
  attn_ln_bwd_kernel<<<grid_for(S, 8, 4), 256, size_t(3) * od * sizeof(float), st>>>(
      a->Y, a->bo, a->R, a->lnw, d_out, S, od, a->eps, a->dV, dln_w, dln_b, db_O);
  TGM_LAUNCH_CHECK();
  // Y = O W_O^T
  BWD_BLAS(gemm_rm(a->blas, false, false, S, od, od, a->dV, od, a->Wo, od, a->dO, od, 0.f));
  BWD_BLAS(gemm_rm(a->blas, true, false, od, od, S, a->dV, od, a->O, od, dW_O, od, 1.f));
  const float *Wk = a->Wkv, *Wv = a->Wkv + size_t(od) * key;
  float *dWk = dW_KV, *dWv = dW_KV + size_t(od) * key;
  for (int h = 0; h < H; ++h) {  // O_h = U_h W_V,h^T
    BWD_BLAS(gemm_rm(a->blas, false, false, S, key, hd, a->dO + h * hd, od,
                     Wv + size_t(h) * hd * key, key, a->dU + h * key, H * key, 0.f));
    BWD_BLAS(gemm_rm(a->blas, true, false, hd, key, S, a->dO + h * hd, od, a->U + h * key, H * key,
                     dWv + size_t(h) * hd * key, key, 1.f));
  }
  const size_t smem = (size_t(k) * key + 2 * size_t(H) * key + 2 * size_t(H) * k +
                       2 * size_t(a->time_dim)) * sizeof(float);
  TGM_REQUIRE(smem <= 200 * 1024, "tgm_attn_backward: k * key_dim too large for shared memory");
  if (smem > 48 * 1024)
    TGM_CUDA(cudaFuncSetAttribute(attn_neighbor_bwd_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024))));
  attn_neighbor_bwd_kernel<<<grid_for(S, 1, per_sm), kBwdThreads, smem, st>>>(
      nbr_node_feat, edge_feat, seed_t, nbr_t, nbr_id, a->tw, a->tb, a->QK, a->dU, S, k,
      a->node_dim, a->edge_dim, a->time_dim, H, 1.0f / sqrtf(float(hd)), a->dQK, d_nbr_node_feat,
      d_edge_feat, dt2v_w, dt2v_b);
  TGM_LAUNCH_CHECK();
  for (int h = 0; h < H; ++h) {  // qk_h = Q_h W_K,h
    BWD_BLAS(gemm_rm(a->blas, false, true, S, hd, key, a->dQK + h * key, H * key,
                     Wk + size_t(h) * hd * key, key, a->dQ + h * hd, od, 0.f));
    BWD_BLAS(gemm_rm(a->blas, true, false, hd, key, S, a->Q + h * hd, od, a->dQK + h * key, H * key,
                     dWk + size_t(h) * hd * key, key, 1.f));
  }
  // Q = R W_Q^T ; the residual adds dV to dR
  TGM_CUDA(cudaMemcpyAsync(a->dR, a->dV, size_t(S) * od * 4, cudaMemcpyDeviceToDevice, st));
  BWD_BLAS(gemm_rm(a->blas, false, false, S, od, od, a->dQ, od, a->Wq, od, a->dR, od, 1.f));
  BWD_BLAS(gemm_rm(a->blas, true, false, od, od, S, a->dQ, od, a->R, od, dW_Q, od, 1.f));
  attn_residual_bwd_kernel<<<grid_for(S, 8, 4), 256, size_t(a->time_dim) * sizeof(float), st>>>(
      a->dR, a->tb, S, a->node_dim, a->pad_dim, a->time_dim, d_node_x, dt2v_b);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
