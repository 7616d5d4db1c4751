// Internal definition of the edge-store handle (shared by store.cu and csr.cu).
#pragma once

#include <vector>

#include "common.cuh"

struct tgm_store {
  int64_t E = 0;
  int32_t D = 0;
  int32_t num_nodes = 0;
  int device = -1;
  bool owns_device = false;
  // device slabs, time-sorted (structure of arrays: a slice is a pointer offset)
  const int32_t *src = nullptr;
  const int32_t *dst = nullptr;
  const int64_t *t = nullptr;
  const float *x = nullptr;  // [E, D] row-major, NULL when D == 0
  // host mirror of t for O(log E) slice bounds without touching the device: owned (uploaded
  // stores), borrowed from the caller, or absent (adopted device stream: bounds search on device)
  const int64_t *t_host = nullptr;
  std::vector<int64_t> t_owned;
  int64_t *scratch = nullptr;  // 2 device words: order check / device bounds
  ~tgm_store();
};

// Read-only view of the per-node adjacency handle (defined in csr.cu) for other translation units.
struct tgm_csr;
namespace tgm {
struct CsrView {
  int device;
  const Entry *entries;      // grouped by node, per node ordered (batch, time, side, edge)
  const int64_t *rowptr;     // [N + 1]
  const float *x;            // the store's feature rows [E, D] (NULL when D == 0)
  int32_t N, D;
  int64_t bs, e_start, Ew;
};
CsrView csr_view(const tgm_csr *c);
}  // namespace tgm
