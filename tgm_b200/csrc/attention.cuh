// Internal definition of the attention handle (shared by attention.cu and attention_bwd.cu).
#pragma once

#include <cublas_v2.h>

#include "common.cuh"

struct tgm_attn {
  int device = -1;
  int H = 0, node_dim = 0, edge_dim = 0, time_dim = 0, pad_dim = 0, out_dim = 0, hd = 0, key = 0;
  float eps = 1e-5f;
  // parameters (device copies)
  float *Wq = nullptr, *Wkv = nullptr, *Wo = nullptr, *bo = nullptr, *lnw = nullptr, *lnb = nullptr;
  float *tw = nullptr, *tb = nullptr;  // Time2Vec weight [time_dim], bias [time_dim]
  float *t0 = nullptr;                 // Time2Vec(0) = cos(b)   [time_dim]
  // folded weights of the inference path (attn_fold.cu), rebuilt whenever the parameters change:
  //   Wqx [H*key, node_dim]  = (W_K,h^T W_Q,h)[:, :node_dim]      qk = Wqx x + cqk
  //   cqk [H*key]            = (W_K,h^T W_Q,h)[:, time part] Time2Vec(0)
  //   Wov [Np, Kp]           = W_O[:, h] W_V,h   (zero padded to multiples of 4)
  float *Wqx = nullptr, *cqk = nullptr, *Wov = nullptr, *zeros = nullptr, *qt0 = nullptr;
  int Np = 0, Kp = 0;
  cublasHandle_t blas = nullptr;
  // workspace, grown on demand (rows = seeds); U rows hold Kp floats, Y rows Np
  int64_t cap = 0;
  float *R = nullptr, *Q = nullptr, *QK = nullptr, *U = nullptr, *O = nullptr, *Y = nullptr;
  // backward workspace (rows = seeds)
  int64_t bcap = 0;
  float *dV = nullptr, *dO = nullptr, *dU = nullptr, *dQK = nullptr, *dQ = nullptr, *dR = nullptr,
        *fwd_out = nullptr;
  ~tgm_attn() {
    if (device >= 0) {
      tgm::DeviceGuard g(device);
      for (float *p : {Wq, Wkv, Wo, bo, lnw, lnb, tw, tb, t0, Wqx, cqk, Wov, zeros, qt0, R, Q, QK, U,
                       O, Y, dV, dO, dU, dQK, dQ, dR, fwd_out})
        cudaFree(p);
      if (blas) cublasDestroy(blas);
    }
  }
};


// forward pass into the handle's workspace (attention.cu).  keep_intermediates: run the unfolded
// chain and leave R, Q, QK, U, O, Y valid (what the backward pass reads); otherwise the folded
// inference chain (attn_fold.cu) serves the call whenever its kernel covers the shape.
int attn_forward_impl(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                      const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                      const float *seed_tf, const float *nbr_tf, const int32_t *nbr_id, int64_t S,
                      int32_t k, float *out, tgm_stream stream, const int32_t *edge_rows = nullptr,
                      bool keep_intermediates = false);

// Per-hop inputs of one folded attention call over several hops of a TGAT layer: segment i covers
// seeds [end[i-1], end[i]).  nid / nt are (rows_i, k), st (rows_i); the edge features are either a
// dense (rows_i, k, edge_dim) block per segment (ef) or -- table != nullptr -- row ids per slot
// (er, -1 = zeros) into the shared feature table.
struct HopSegs {
  const int32_t *nid[4];
  const int64_t *nt[4];
  const int64_t *st[4];
  const float *ef[4];
  const int32_t *er[4];
  int64_t end[4];
  const float *table;
  int n;
};
// where the LayerNorm pass writes: `dst` rows of `pitch` floats; optionally the merge layer's
// second input x2 [S, nd2] is copied behind the od output columns and the rest of the row zeroed
// (the row is then MergeLayer's concatenated, padded input)
struct LnTarget {
  float *dst;
  int pitch;
  const float *x2;
  int nd2;
};

struct tgm_mlp2 {
  int device = -1;
  int in1 = 0, in2 = 0, hidden = 0, out = 0;
  int inp = 0;  // in1 + in2 rounded up to a multiple of 4: row pitch of `cat` and of W1 (zero padded)
  float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr;
  cublasHandle_t blas = nullptr;
  int64_t cap = 0;
  float *cat = nullptr, *h = nullptr;
  ~tgm_mlp2() {
    if (device >= 0) {
      tgm::DeviceGuard g(device);
      for (float *p : {W1, b1, W2, b2, cat, h}) cudaFree(p);
      if (blas) cublasDestroy(blas);
    }
  }
};
int mlp2_workspace(tgm_mlp2 *m, int64_t S, cudaStream_t st);
// fc2(relu(fc1(m->cat[:S]))) -> out (the caller has filled m->cat rows of m->inp floats)
int mlp2_forward_cat(tgm_mlp2 *m, int64_t S, float *out, cudaStream_t st);

// attn_fold.cu
int attn_fold_alloc(tgm_attn *a);                       // once, at create
int attn_fold_refresh(tgm_attn *a, cudaStream_t st);    // after every parameter change
bool attn_folded_covers(const tgm_attn *a, int k);
int attn_workspace(tgm_attn *a, int64_t S, cudaStream_t st);
int attn_forward_folded(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                        const HopSegs &hops, int64_t S, int32_t k, const LnTarget &target,
                        cudaStream_t st);
HopSegs single_hop(const float *edge_feat, const int32_t *edge_rows, const int64_t *seed_t,
                   const int64_t *nbr_t, const int32_t *nbr_id, int64_t S);
// C[S, N] = act(A[S, K] W[N, K]^T + bias): the hand-written tensor-core kernel when the shape
// allows, cuBLAS SGEMM + one elementwise pass otherwise.  act: 0 none, 2 ReLU.
int dense_linear(cublasHandle_t blas, int64_t S, int N, int K, const float *A, const float *W,
                 const float *bias, int act, float *out, cudaStream_t st);
