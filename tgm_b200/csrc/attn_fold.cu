// The inference chain of TemporalAttention with the projection weights folded.
//
// attention.cu reassociates the reference contraction (tgm/nn/modules/attention.py:93-128) so the
// per-neighbour K/V projections disappear.  With no gradient to keep, two more steps fold away:
//
//   qk_h = W_K,h^T (W_Q,h r)          r = [x | 0 | Time2Vec(0)]  (attention.py:93-95, tgat.py:141)
//        = (W_K,h^T W_Q,h)[:, :node_dim] x + (W_K,h^T W_Q,h)[:, time part] cos(b)
//        =: Wqx_h x + cqk_h                                   -- K = node_dim, not out_dim, and no
//                                                                R / Q matrices in memory
//   y    = W_O cat_h(W_V,h u_h) = sum_h (W_O[:, h] W_V,h) u_h =: Wov u        -- one product, no O
//
// The folded matrices are a few hundred kB, computed in float64 and rounded once, every time the
// parameters change (create / set_params).  A call is then
//   [node_dim > 1: QK = X Wqx^T]  ->  attn_warp_kernel  ->  Y = U Wov^T  ->  LayerNorm epilogue
// 3-4 launches against 7, and the neighbour kernel is one WARP per seed: every lane owns columns
// lane, lane+32, ... of the key vector in registers, rows stream through once (next row
// prefetched), the softmax is the running-maximum form, and nothing goes through shared memory or
// a barrier.  The backward pass keeps the unfolded chain (it needs Q, O, ...).
#include <cublas_v2.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

using namespace tgm;

#include "attention.cuh"

namespace tgm {
int g_attn_folded = 1;  // tgm_set_option("attn_folded", 0|1): 0 keeps inference on the unfolded chain
}

namespace {

int blas_fail2(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define TGM_BLAS2(expr)                                            \
  do {                                                             \
    cublasStatus_t _s = (expr);                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) return blas_fail2(_s, #expr); \
  } while (0)

// ---- weight folding (float64 accumulation, one rounding) ----------------------------------------
// qt0[r] = sum_c W_Q[r, nd + pad + c] cos(b_c)
__global__ void fold_qt0_kernel(const float *__restrict__ Wq, const float *__restrict__ t0, int od,
                                int toff, int td, float *__restrict__ qt0_hi,
                                float *__restrict__ qt0_lo) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= od) return;
  double acc = 0.0;
  for (int c = 0; c < td; ++c) acc += double(Wq[size_t(r) * od + toff + c]) * double(t0[c]);
  const float hi = float(acc);
  qt0_hi[r] = hi;
  qt0_lo[r] = float(acc - double(hi));
}

// column i < nd: Wqx[n, i] = sum_d W_K[(h, d), j] W_Q[(h, d), i]  ([H key, nd]: the W of an
// "A W^T" product; with node_dim == 1, the in-kernel case, lanes read consecutive n);
// i == nd: cqk[n] with qt0
__global__ void fold_qk_kernel(const float *__restrict__ Wq, const float *__restrict__ Wk,
                               const float *__restrict__ qt0_hi, const float *__restrict__ qt0_lo,
                               int H, int hd, int od, int key, int nd, float *__restrict__ Wqx,
                               float *__restrict__ cqk) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(H) * key * (nd + 1);
  if (idx >= total) return;
  const int n = int(idx / (nd + 1)), i = int(idx - int64_t(n) * (nd + 1));
  const int h = n / key, j = n - h * key;
  double acc = 0.0;
  for (int d = 0; d < hd; ++d) {
    const int r = h * hd + d;
    const double wk = double(Wk[size_t(r) * key + j]);
    const double q = i < nd ? double(Wq[size_t(r) * od + i]) : double(qt0_hi[r]) + double(qt0_lo[r]);
    acc += wk * q;
  }
  if (i < nd) Wqx[size_t(n) * nd + i] = float(acc);
  else cqk[n] = float(acc);
}

// Wov[o, h * key + j] = sum_d W_O[o, h hd + d] W_V[(h, d), j]; rows >= od and columns >= H key: 0
__global__ void fold_ov_kernel(const float *__restrict__ Wo, const float *__restrict__ Wv, int H,
                               int hd, int od, int key, int Np, int Kp, float *__restrict__ Wov) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(Np) * Kp) return;
  const int o = int(idx / Kp), n = int(idx - int64_t(o) * Kp);
  double acc = 0.0;
  if (o < od && n < H * key) {
    const int h = n / key, j = n - h * key;
    for (int d = 0; d < hd; ++d)
      acc += double(Wo[size_t(o) * od + h * hd + d]) * double(Wv[size_t(h * hd + d) * key + j]);
  }
  Wov[idx] = float(acc);
}

// ---- the neighbour pass, one warp per seed ------------------------------------------------------
constexpr int kWarpThreads = 128;
constexpr int kTT = 4;  // time columns per lane (time_dim <= 128)

// sums over the warp of four values at once: two exchange steps leave one value per quarter-warp
// group (bits 4 and 3 of the lane pick which), three butterfly steps finish them, four broadcasts
// hand them out
__device__ __forceinline__ void warp_sum4(float &a, float &b, float &c, float &d, int lane) {
  const bool h16 = lane & 16, h8 = lane & 8;
  // lanes with bit 4 clear keep (a, b), the others (c, d)
  float u = (h16 ? c : a) + __shfl_xor_sync(0xffffffffu, h16 ? a : c, 16);
  float v = (h16 ? d : b) + __shfl_xor_sync(0xffffffffu, h16 ? b : d, 16);
  // bit 3 clear keeps the first of the pair
  float w = (h8 ? v : u) + __shfl_xor_sync(0xffffffffu, h8 ? u : v, 8);
#pragma unroll
  for (int o = 4; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  a = __shfl_sync(0xffffffffu, w, 0);
  b = __shfl_sync(0xffffffffu, w, 8);
  c = __shfl_sync(0xffffffffu, w, 16);
  d = __shfl_sync(0xffffffffu, w, 24);
}

// INLINE_Q: node_dim == 1 (TGAT layer 1 on a featureless graph): qk = x Wqx + cqk built here;
// otherwise the x part arrives as QK = X Wqx from a plain product.
template <int TN, int TE, int MINB, bool INLINE_Q>
__global__ void __launch_bounds__(kWarpThreads, MINB)
attn_warp_kernel(const float *__restrict__ X, const float *__restrict__ nbr_feat,
                 const HopSegs segs, const float *__restrict__ tw,
                 const float *__restrict__ tb, const float *__restrict__ QK,
                 const float *__restrict__ Wqx, const float *__restrict__ cqk, int64_t S, int k,
                 int nd, int ed, int td, int H, float scale, int Kp, float *__restrict__ U) {
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * (kWarpThreads / 32) + (threadIdx.x >> 5);
  if (s >= S) return;
  const int key = nd + ed + td;
  const bool two = H > 1;
  const int64_t base = s * k;

  // this seed's hop segment (warp-uniform), then the per-slot scalars: lane n holds slot n (k <= 32)
  int seg = 0;
  int64_t first = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i + 1 < segs.n && s >= segs.end[i]) seg = i + 1, first = segs.end[i];
  const int64_t lbase = (s - first) * k;  // slot 0 of this seed inside its segment
  const int64_t tq = segs.st[seg][s - first];
  const bool lazy = segs.table != nullptr;
  int my_id = TGM_PADDED_NODE_ID, my_er = -1;
  float my_dt = 0.f;
  if (lane < k) {
    my_id = __ldg(segs.nid[seg] + lbase + lane);
    my_dt = float(tq - __ldg(segs.nt[seg] + lbase + lane));  // int64 difference, then .float() (:23)
    my_er = lazy ? __ldg(segs.er[seg] + lbase + lane) : lane;
  }
  const float *ef = lazy ? segs.table : segs.ef[seg] + lbase * ed;
  const float *nf = nbr_feat + base * nd;

  // qk of both heads, this lane's columns (column c of a segment at offset `off`: index off + c of
  // head 0, key + off + c of head 1)
  float qn0[TN], qn1[TN], qe0[TE], qe1[TE], qt0[kTT], qt1[kTT];
  const float *cq = cqk + lane;
  const float *xq = INLINE_Q ? Wqx + lane : QK + s * int64_t(H) * key + lane;
  const float x0 = INLINE_Q ? __ldg(X + s) : 1.f;
  auto qk_at = [&](int n) -> float {  // n: index minus the lane
    return INLINE_Q ? fmaf(x0, __ldg(xq + n), __ldg(cq + n)) : __ldg(xq + n) + __ldg(cq + n);
  };
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    const bool in = lane + 32 * t < nd;
    qn0[t] = in ? qk_at(32 * t) : 0.f;
    qn1[t] = (two && in) ? qk_at(key + 32 * t) : 0.f;
  }
#pragma unroll
  for (int t = 0; t < TE; ++t) {
    const bool in = lane + 32 * t < ed;
    qe0[t] = in ? qk_at(nd + 32 * t) : 0.f;
    qe1[t] = (two && in) ? qk_at(key + nd + 32 * t) : 0.f;
  }
  float wr[kTT], br[kTT];
  bool small = true;  // every Time2Vec argument of this seed inside the cosine's reduced-range path
  {
    float dmax = fabsf(my_dt);
#pragma unroll
    for (int o = 16; o; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
#pragma unroll
    for (int t = 0; t < kTT; ++t) {
      const bool in = lane + 32 * t < td;
      wr[t] = in ? __ldg(tw + lane + 32 * t) : 0.f;
      br[t] = in ? __ldg(tb + lane + 32 * t) : 0.f;
      qt0[t] = in ? qk_at(nd + ed + 32 * t) : 0.f;
      qt1[t] = (two && in) ? qk_at(key + nd + ed + 32 * t) : 0.f;
      small = small && (fmaf(dmax, fabsf(wr[t]), fabsf(br[t])) < 4.0e6f);
    }
    small = __all_sync(0xffffffffu, small);
  }

  float an0[TN], an1[TN], ae0[TE], ae1[TE], at0[kTT], at1[kTT];
#pragma unroll
  for (int t = 0; t < TN; ++t) an0[t] = an1[t] = 0.f;
#pragma unroll
  for (int t = 0; t < TE; ++t) ae0[t] = ae1[t] = 0.f;
#pragma unroll
  for (int t = 0; t < kTT; ++t) at0[t] = at1[t] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // rows go two at a time: the loads of both are issued, the eight cosine tiles are evaluated
  // while they fly, four dot products are reduced together
#pragma unroll 1
  for (int n = 0; n < k; n += 2) {
    const bool hasB = n + 1 < k;
    const int nB = hasB ? n + 1 : n;
    const int erA = __shfl_sync(0xffffffffu, my_er, n), erB = __shfl_sync(0xffffffffu, my_er, nB);
    float znA[TN], znB[TN], zeA[TE], zeB[TE];
    {
      const float *rA = nf + n * nd, *rB = nf + nB * nd;
#pragma unroll
      for (int t = 0; t < TN; ++t) {
        const int c = lane + 32 * t;
        znA[t] = c < nd ? __ldg(rA + c) : 0.f;
        znB[t] = c < nd ? __ldg(rB + c) : 0.f;
      }
      const float *eA = ef + int64_t(erA < 0 ? 0 : erA) * ed, *eB = ef + int64_t(erB < 0 ? 0 : erB) * ed;
#pragma unroll
      for (int t = 0; t < TE; ++t) {
        const int c = lane + 32 * t;
        zeA[t] = (erA >= 0 && c < ed) ? __ldg(eA + c) : 0.f;
        zeB[t] = (erB >= 0 && c < ed) ? __ldg(eB + c) : 0.f;
      }
    }
    const float dtA = __shfl_sync(0xffffffffu, my_dt, n), dtB = __shfl_sync(0xffffffffu, my_dt, nB);
    const bool vA = __shfl_sync(0xffffffffu, my_id, n) != TGM_PADDED_NODE_ID;
    const bool vB = __shfl_sync(0xffffffffu, my_id, nB) != TGM_PADDED_NODE_ID;
    float ztA[kTT], ztB[kTT];
    if (small) {
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        const bool on = 32 * t < td;
        ztA[t] = on ? t2v_cos_fast(__fmaf_rn(dtA, wr[t], br[t])) : 0.f;
        ztB[t] = on ? t2v_cos_fast(__fmaf_rn(dtB, wr[t], br[t])) : 0.f;
      }
    } else {
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        const bool on = 32 * t < td;
        ztA[t] = on ? t2v_cos(__fmaf_rn(dtA, wr[t], br[t])) : 0.f;
        ztB[t] = on ? t2v_cos(__fmaf_rn(dtB, wr[t], br[t])) : 0.f;
      }
    }
    float pA0 = 0.f, pA1 = 0.f, pB0 = 0.f, pB1 = 0.f;
#pragma unroll
    for (int t = 0; t < kTT; ++t) {
      pA0 = fmaf(qt0[t], ztA[t], pA0), pA1 = fmaf(qt1[t], ztA[t], pA1);
      pB0 = fmaf(qt0[t], ztB[t], pB0), pB1 = fmaf(qt1[t], ztB[t], pB1);
    }
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      pA0 = fmaf(qn0[t], znA[t], pA0), pA1 = fmaf(qn1[t], znA[t], pA1);
      pB0 = fmaf(qn0[t], znB[t], pB0), pB1 = fmaf(qn1[t], znB[t], pB1);
    }
#pragma unroll
    for (int t = 0; t < TE; ++t) {
      pA0 = fmaf(qe0[t], zeA[t], pA0), pA1 = fmaf(qe1[t], zeA[t], pA1);
      pB0 = fmaf(qe0[t], zeB[t], pB0), pB1 = fmaf(qe1[t], zeB[t], pB1);
    }
    warp_sum4(pA0, pA1, pB0, pB1, lane);
    // masked slots take -1e10 after the scaling (attention.py:110-113); a missing second row takes
    // -inf (weight exactly 0).  Running-maximum softmax: the accumulators are rescaled only when a
    // pair raises the maximum (warp-uniform branch)
    {
      const float lgA = vA ? pA0 * scale : -1e10f;
      const float lgB = hasB ? (vB ? pB0 * scale : -1e10f) : -INFINITY;
      const float mx = fmaxf(lgA, lgB);
      if (mx > m0) {
        const float c = expf(m0 - mx);
        l0 *= c;
#pragma unroll
        for (int t = 0; t < TN; ++t) an0[t] *= c;
#pragma unroll
        for (int t = 0; t < TE; ++t) ae0[t] *= c;
#pragma unroll
        for (int t = 0; t < kTT; ++t) at0[t] *= c;
        m0 = mx;
      }
      const float eA = expf(lgA - m0), eB = expf(lgB - m0);
      l0 += eA + eB;
#pragma unroll
      for (int t = 0; t < TN; ++t) an0[t] = fmaf(eB, znB[t], fmaf(eA, znA[t], an0[t]));
#pragma unroll
      for (int t = 0; t < TE; ++t) ae0[t] = fmaf(eB, zeB[t], fmaf(eA, zeA[t], ae0[t]));
#pragma unroll
      for (int t = 0; t < kTT; ++t) at0[t] = fmaf(eB, ztB[t], fmaf(eA, ztA[t], at0[t]));
    }
    if (two) {
      const float lgA = vA ? pA1 * scale : -1e10f;
      const float lgB = hasB ? (vB ? pB1 * scale : -1e10f) : -INFINITY;
      const float mx = fmaxf(lgA, lgB);
      if (mx > m1) {
        const float c = expf(m1 - mx);
        l1 *= c;
#pragma unroll
        for (int t = 0; t < TN; ++t) an1[t] *= c;
#pragma unroll
        for (int t = 0; t < TE; ++t) ae1[t] *= c;
#pragma unroll
        for (int t = 0; t < kTT; ++t) at1[t] *= c;
        m1 = mx;
      }
      const float eA = expf(lgA - m1), eB = expf(lgB - m1);
      l1 += eA + eB;
#pragma unroll
      for (int t = 0; t < TN; ++t) an1[t] = fmaf(eB, znB[t], fmaf(eA, znA[t], an1[t]));
#pragma unroll
      for (int t = 0; t < TE; ++t) ae1[t] = fmaf(eB, zeB[t], fmaf(eA, zeA[t], ae1[t]));
#pragma unroll
      for (int t = 0; t < kTT; ++t) at1[t] = fmaf(eB, ztB[t], fmaf(eA, ztA[t], at1[t]));
    }
  }

  // u_h = acc_h / l_h, row pitch Kp (columns [H key, Kp) are zero for the padded product)
  float *u = U + s * int64_t(Kp);
  const float i0 = 1.f / l0, i1 = two ? 1.f / l1 : 0.f;
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    const int c = lane + 32 * t;
    if (c < nd) {
      u[c] = an0[t] * i0;
      if (two) u[key + c] = an1[t] * i1;
    }
  }
#pragma unroll
  for (int t = 0; t < TE; ++t) {
    const int c = lane + 32 * t;
    if (c < ed) {
      u[nd + c] = ae0[t] * i0;
      if (two) u[key + nd + c] = ae1[t] * i1;
    }
  }
#pragma unroll
  for (int t = 0; t < kTT; ++t) {
    const int c = lane + 32 * t;
    if (c < td) {
      u[nd + ed + c] = at0[t] * i0;
      if (two) u[key + nd + ed + c] = at1[t] * i1;
    }
  }
  if (lane < Kp - H * key) u[H * key + lane] = 0.f;
}

// out = LayerNorm(Y[:, :od] + b_O + [x | 0 | Time2Vec(0)]) (attention.py:124-127); one warp per
// row, the row kept in registers between the passes (out_dim <= 384), two-pass variance
constexpr int kLnTiles = 12;
__global__ void __launch_bounds__(256)
attn_ln_kernel(const float *__restrict__ Y, int ldy, const float *__restrict__ bo,
               const float *__restrict__ X, const float *__restrict__ t0,
               const float *__restrict__ lnw, const float *__restrict__ lnb, int64_t S, int od,
               int nd, int toff, float eps, float *__restrict__ dst, int pitch,
               const float *__restrict__ x2, int nd2) {
  const int lane = threadIdx.x & 31;
  const int64_t s = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const float *y = Y + s * ldy;
  float v[kLnTiles];
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < kLnTiles; ++t) {
    const int c = lane + 32 * t;
    float r = 0.f;
    if (c < od) {
      r = y[c] + __ldg(bo + c);
      if (c < nd) r += __ldg(X + s * nd + c);
      else if (c >= toff) r += __ldg(t0 + c - toff);
    }
    v[t] = r;
    sum += r;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / float(od);
  float var = 0.f;
#pragma unroll
  for (int t = 0; t < kLnTiles; ++t) {
    const float d = lane + 32 * t < od ? v[t] - mean : 0.f;
    var = fmaf(d, d, var);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / float(od) + eps);
#pragma unroll
  for (int t = 0; t < kLnTiles; ++t) {
    const int c = lane + 32 * t;
    if (c < od) dst[s * pitch + c] = (v[t] - mean) * rstd * __ldg(lnw + c) + __ldg(lnb + c);
  }
  // the rest of a wider destination row: the merge layer's second input, then zeros
  for (int c = od + lane; c < pitch; c += 32)
    dst[s * pitch + c] = (x2 && c - od < nd2) ? __ldg(x2 + s * nd2 + (c - od)) : 0.f;
}

__global__ void bias_act2_kernel(float *__restrict__ x, const float *__restrict__ b, int64_t S,
                                 int d, int relu) {
  const int64_t total = S * d;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float v = x[i] + __ldg(b + int(i % d));
    x[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

template <int TN, int TE, int MINB>
int launch_warp(const tgm_attn *a, const float *X, const float *nbr_feat, const HopSegs &hops,
                const float *QK, int64_t S, int k, cudaStream_t st) {
  const int wpb = kWarpThreads / 32;
  const unsigned grid = unsigned((S + wpb - 1) / wpb);
  const float scale = 1.0f / sqrtf(float(a->hd));
  if (QK)
    attn_warp_kernel<TN, TE, MINB, false><<<grid, kWarpThreads, 0, st>>>(
        X, nbr_feat, hops, a->tw, a->tb, QK, a->Wqx, a->cqk,
        S, k, a->node_dim, a->edge_dim, a->time_dim, a->H, scale, a->Kp, a->U);
  else
    attn_warp_kernel<TN, TE, MINB, true><<<grid, kWarpThreads, 0, st>>>(
        X, nbr_feat, hops, a->tw, a->tb, QK, a->Wqx, a->cqk,
        S, k, a->node_dim, a->edge_dim, a->time_dim, a->H, scale, a->Kp, a->U);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

}  // namespace

// which engine for C[S, N] = A[S, K] W[N, K]^T: measured on the B200 (profiles/r2_tc_linear_timings.txt)
//   * no activation, >= 2048 rows: the CUTLASS FastF32 collective (31 vs 42 us at 12600x104x548)
//   * short matrices (<= 4096 rows) WITH a bias / activation: the 32x32-tile SIMT kernel of
//     small_gemm.cu, one launch with the epilogue fused (600x172x444 + ReLU: 12 us against
//     cuBLAS split-K + reduce + bias pass 23 us; 600x172x172: 9 vs 18 on the tensor-core kernel,
//     whose 128-row tiles leave most SMs idle there).  Without an epilogue cuBLAS split-K keeps the
//     short products (600x888x172: 11 vs 18 us; 600x272x888: 32 vs 36)
//   * otherwise the hand-written tcgen05 kernel with bias / ReLU in its epilogue
constexpr int64_t kSmallRows = 4096;
static bool tc3_suits(int64_t S, int K) {
  (void)K;
  return g_tc_linear != 0 && S > kSmallRows;
}

int dense_linear(cublasHandle_t blas, int64_t S, int N, int K, const float *A, const float *W,
                 const float *bias, int act, float *out, cudaStream_t st) {
  if (act == 0 && S >= 2048 && g_gemm_fastf32) {
    const int rc = fastf32_linear(S, N, K, A, W, bias, nullptr, 0, out, st);
    if (rc != 0) return rc < 0 ? rc : TGM_OK;
  }
  if (S <= kSmallRows) {
    const int rc = small_gemm_nt(S, N, K, A, K, 0, W, K, 0, bias, act, out, N, 0, 1, st);
    if (rc != 0) return rc < 0 ? rc : TGM_OK;
  }
  if (tc3_suits(S, K)) {
    const int rc = tc3_linear(S, N, K, A, W, bias, nullptr, act, out, st);
    if (rc != 0) return rc < 0 ? rc : TGM_OK;
  }
  const float one = 1.f, zero = 0.f;
  TGM_BLAS2(cublasSetStream(blas, st));
  TGM_BLAS2(cublasSgemm(blas, CUBLAS_OP_T, CUBLAS_OP_N, N, int(S), K, &one, W, K, A, K, &zero, out,
                        N));
  bias_act2_kernel<<<grid_for(S * N, 256, 8), 256, 0, st>>>(out, bias, S, N, act == 2);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

int attn_fold_alloc(tgm_attn *a) {
  a->Np = (a->out_dim + 3) & ~3;
  a->Kp = (a->H * a->key + 3) & ~3;
  const size_t hk = size_t(a->H) * a->key;
  TGM_CUDA(cudaMalloc(&a->Wqx, hk * a->node_dim * 4));
  TGM_CUDA(cudaMalloc(&a->cqk, hk * 4));
  TGM_CUDA(cudaMalloc(&a->Wov, size_t(a->Np) * a->Kp * 4));
  TGM_CUDA(cudaMalloc(&a->zeros, size_t(a->Np) * 4));
  TGM_CUDA(cudaMalloc(&a->qt0, size_t(a->out_dim) * 2 * 4));
  TGM_CUDA(cudaMemset(a->zeros, 0, size_t(a->Np) * 4));
  return TGM_OK;
}

int attn_fold_refresh(tgm_attn *a, cudaStream_t st) {
  const int od = a->out_dim, key = a->key, H = a->H, hd = a->hd, nd = a->node_dim;
  fold_qt0_kernel<<<(od + 127) / 128, 128, 0, st>>>(a->Wq, a->t0, od, nd + a->pad_dim, a->time_dim,
                                                    a->qt0, a->qt0 + od);
  TGM_LAUNCH_CHECK();
  const int64_t nq = int64_t(H) * key * (nd + 1);
  fold_qk_kernel<<<unsigned((nq + 127) / 128), 128, 0, st>>>(a->Wq, a->Wkv, a->qt0, a->qt0 + od, H,
                                                             hd, od, key, nd, a->Wqx, a->cqk);
  TGM_LAUNCH_CHECK();
  const int64_t no = int64_t(a->Np) * a->Kp;
  fold_ov_kernel<<<unsigned((no + 127) / 128), 128, 0, st>>>(
      a->Wo, a->Wkv + size_t(od) * key, H, hd, od, key, a->Np, a->Kp, a->Wov);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

bool attn_folded_covers(const tgm_attn *a, int k) {
  return g_attn_folded && k <= 32 && a->H <= 2 && a->node_dim <= 192 && a->edge_dim <= 192 &&
         a->time_dim <= 32 * kTT && a->out_dim <= 32 * kLnTiles;
}

HopSegs single_hop(const float *edge_feat, const int32_t *edge_rows, const int64_t *seed_t,
                   const int64_t *nbr_t, const int32_t *nbr_id, int64_t S) {
  HopSegs h{};
  h.n = 1;
  h.nid[0] = nbr_id, h.nt[0] = nbr_t, h.st[0] = seed_t, h.end[0] = S;
  if (edge_rows) h.table = edge_feat, h.er[0] = edge_rows;
  else h.ef[0] = edge_feat;
  for (int i = 1; i < 4; ++i) h.end[i] = S;
  return h;
}

int attn_forward_folded(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                        const HopSegs &hops, int64_t S, int32_t k, const LnTarget &target,
                        cudaStream_t st) {
  const int od = a->out_dim, key = a->key, H = a->H, nd = a->node_dim;
  const float *QK = nullptr;
  if (nd > 1) {  // qk's x part is a plain product; a single node column folds into the kernel
    const float one = 1.f, zero = 0.f;
    TGM_BLAS2(cublasSetStream(a->blas, st));
    TGM_BLAS2(cublasSgemm(a->blas, CUBLAS_OP_T, CUBLAS_OP_N, H * key, int(S), nd, &one, a->Wqx, nd,
                          node_x, nd, &zero, a->QK, H * key));
    QK = a->QK;
  }
  int rc;
  const bool wide_n = nd > 32, wide_e = a->edge_dim > 64;
  if (!wide_n && !wide_e)
    rc = launch_warp<1, 2, 5>(a, node_x, nbr_node_feat, hops, QK, S, k, st);
  else if (!wide_n)  // (5 CTAs/SM at 96 registers measured the same: 289.4 vs 289.8 us per forward)
    rc = launch_warp<1, 6, 4>(a, node_x, nbr_node_feat, hops, QK, S, k, st);
  else if (!wide_e)
    rc = launch_warp<6, 2, 4>(a, node_x, nbr_node_feat, hops, QK, S, k, st);
  else
    rc = launch_warp<6, 6, 3>(a, node_x, nbr_node_feat, hops, QK, S, k, st);
  if (rc) return rc;
  // Y = U Wov^T (bias and residual join in the LayerNorm pass)
  rc = 0;
  if (S >= 2048 && g_gemm_fastf32)
    rc = fastf32_linear(S, a->Np, a->Kp, a->U, a->Wov, a->zeros, nullptr, 0, a->Y, st);
  if (rc == 0 && tc3_suits(S, a->Kp))
    rc = tc3_linear(S, a->Np, a->Kp, a->U, a->Wov, a->zeros, nullptr, 0, a->Y, st);
  if (rc < 0) return rc;
  if (rc == 0) {
    const float one = 1.f, zero = 0.f;
    TGM_BLAS2(cublasSetStream(a->blas, st));
    TGM_BLAS2(cublasSgemm(a->blas, CUBLAS_OP_T, CUBLAS_OP_N, a->Np, int(S), a->Kp, &one, a->Wov,
                          a->Kp, a->U, a->Kp, &zero, a->Y, a->Np));
  }
  const int wpb = 8;
  attn_ln_kernel<<<unsigned((S + wpb - 1) / wpb), 32 * wpb, 0, st>>>(
      a->Y, a->Np, a->bo, node_x, a->t0, a->lnw, a->lnb, S, od, nd, nd + a->pad_dim, a->eps,
      target.dst, target.pitch, target.x2, target.nd2);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
