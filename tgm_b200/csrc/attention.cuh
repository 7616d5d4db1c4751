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
  cublasHandle_t blas = nullptr;
  // workspace, grown on demand (rows = seeds)
  int64_t cap = 0;
  float *R = nullptr, *Q = nullptr, *QK = nullptr, *U = nullptr, *O = nullptr, *Y = nullptr;
  // backward workspace (rows = seeds)
  int64_t bcap = 0;
  float *dV = nullptr, *dO = nullptr, *dU = nullptr, *dQK = nullptr, *dQ = nullptr, *dR = nullptr,
        *fwd_out = nullptr;
  ~tgm_attn() {
    if (device >= 0) {
      tgm::DeviceGuard g(device);
      for (float *p : {Wq, Wkv, Wo, bo, lnw, lnb, tw, tb, t0, R, Q, QK, U, O, Y, dV, dO, dU, dQK, dQ,
                       dR, fwd_out})
        cudaFree(p);
      if (blas) cublasDestroy(blas);
    }
  }
};


// forward pass into the handle's workspace (attention.cu); leaves R, Q, QK, U, O, Y valid
int attn_forward_impl(tgm_attn *a, const float *node_x, const float *nbr_node_feat,
                      const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                      const float *seed_tf, const float *nbr_tf, const int32_t *nbr_id, int64_t S,
                      int32_t k, float *out, tgm_stream stream, const int32_t *edge_rows = nullptr);
