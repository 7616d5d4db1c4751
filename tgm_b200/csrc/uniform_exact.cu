// Uniform full-history sampler, reference-exact mode.
//
// DGStorageArrayBackend.get_nbrs (reference tgm-team/tgm @ 5183dc9, tgm/core/_storage/backends/
// array_backend.py:108-171) keeps `random.sample(candidates, k)` -- CPython's global Mersenne
// Twister -- for every unique seed node with more than k candidates (:147-153).  The draw depends
// only on the candidate COUNT and on k, so it can be made on the host with the very same call and
// applied on the device: tgm_csr_candidate_counts returns the counts, the caller draws
// random.sample(range(count), k) per unique node in ascending node order, and
// tgm_csr_gather_picks gathers the chosen candidate ordinals.  Candidates of a node are its
// entries with e_lo <= edge < e_hi in the (edge, side) order of the batch_size-1 adjacency, which
// is the order the reference appends them in (:132-137).
#include "store.cuh"

using namespace tgm;

namespace {

// first index in [lo, hi) whose entry belongs to an edge >= cut (edge-major order)
__device__ __forceinline__ int64_t lower_bound_edge(const Entry *__restrict__ entries, int64_t lo,
                                                    int64_t hi, int64_t cut) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(entries[mid].eid) < cut) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
candidate_counts_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                        int32_t N, const int32_t *__restrict__ seeds, int64_t S, int64_t e_lo,
                        int64_t e_hi, int64_t *__restrict__ counts) {
  for (int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; s < S;
       s += int64_t(gridDim.x) * blockDim.x) {
    const int32_t v = seeds[s];
    int64_t cnt = 0;
    if (v >= 0 && v < N) {
      const int64_t r0 = rowptr[v], r1 = rowptr[v + 1];
      const int64_t lo = lower_bound_edge(entries, r0, r1, e_lo);
      cnt = lower_bound_edge(entries, lo, r1, e_hi) - lo;
    }
    counts[s] = cnt;
  }
}

// One warp per seed: column c takes candidate ordinal picks[s, c] (-1 or out of range = padding).
__global__ void __launch_bounds__(256)
gather_picks_kernel(const Entry *__restrict__ entries, const int64_t *__restrict__ rowptr,
                    const float *__restrict__ x, int32_t N, int D,
                    const int32_t *__restrict__ seeds, int64_t S, int64_t e_lo, int64_t e_hi, int k,
                    const int32_t *__restrict__ picks, int32_t *__restrict__ out_nid,
                    int64_t *__restrict__ out_t, float *__restrict__ out_x) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t s = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); s < S;
       s += int64_t(gridDim.x) * wpb) {
    const int32_t v = seeds[s];
    int64_t lo = 0, cnt = 0;
    if (v >= 0 && v < N) {
      const int64_t r0 = rowptr[v], r1 = rowptr[v + 1];
      lo = lower_bound_edge(entries, r0, r1, e_lo);
      cnt = lower_bound_edge(entries, lo, r1, e_hi) - lo;
    }
    for (int c = 0; c < k; ++c) {  // columns one after the other: lanes stream the feature row
      const int32_t p = picks[s * k + c];
      const bool ok = p >= 0 && int64_t(p) < cnt;
      Entry en;
      en.nbr = TGM_PADDED_NODE_ID, en.eid = 0, en.t = 0;
      if (ok) en = entries[lo + p];
      if (lane == 0) {
        out_nid[s * k + c] = en.nbr;
        out_t[s * k + c] = en.t;
      }
      if (D > 0) {
        float *o = out_x + (s * k + c) * int64_t(D);
        const float *row = x + int64_t(en.eid) * D;
        for (int d = lane; d < D; d += 32) o[d] = ok ? __ldg(row + d) : 0.f;
      }
    }
  }
}

int check_args(const CsrView &c, const char *who, int64_t S, int64_t e_lo, int64_t e_hi) {
  if (!(c.bs == 1 && c.e_start == 0))
    return fail(TGM_ERR_INVALID, std::string(who) +
                                     ": the adjacency must be built with batch_size 1, e_start 0");
  if (S < 0) return fail(TGM_ERR_INVALID, std::string(who) + ": S must be >= 0");
  if (!(0 <= e_lo && e_lo <= e_hi && e_hi <= c.Ew))
    return fail(TGM_ERR_INVALID, std::string(who) + ": [e_lo, e_hi) outside the store");
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_csr_candidate_counts(const tgm_csr *csr, const int32_t *seeds, int64_t S,
                                        int64_t e_lo, int64_t e_hi, int64_t *out_counts,
                                        tgm_stream stream) {
  TGM_REQUIRE(csr != nullptr, "tgm_csr_candidate_counts: csr is NULL");
  const CsrView c = csr_view(csr);
  int rc = check_args(c, "tgm_csr_candidate_counts", S, e_lo, e_hi);
  if (rc != TGM_OK) return rc;
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && out_counts, "tgm_csr_candidate_counts: NULL array argument");
  DeviceGuard g(c.device);
  candidate_counts_kernel<<<grid_for(S, 256, 8), 256, 0, as_stream(stream)>>>(
      c.entries, c.rowptr, c.N, seeds, S, e_lo, e_hi, out_counts);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_csr_gather_picks(const tgm_csr *csr, const int32_t *seeds, int64_t S,
                                    int64_t e_lo, int64_t e_hi, int32_t k, const int32_t *picks,
                                    int32_t *out_nid, int64_t *out_t, float *out_x,
                                    tgm_stream stream) {
  TGM_REQUIRE(csr != nullptr, "tgm_csr_gather_picks: csr is NULL");
  const CsrView c = csr_view(csr);
  int rc = check_args(c, "tgm_csr_gather_picks", S, e_lo, e_hi);
  if (rc != TGM_OK) return rc;
  TGM_REQUIRE(k >= 1, "tgm_csr_gather_picks: k must be >= 1");
  if (S == 0) return TGM_OK;
  TGM_REQUIRE(seeds && picks && out_nid && out_t, "tgm_csr_gather_picks: NULL array argument");
  TGM_REQUIRE(c.D == 0 || out_x != nullptr, "tgm_csr_gather_picks: out_x is NULL but D > 0");
  DeviceGuard g(c.device);
  gather_picks_kernel<<<grid_for(S, 8, 8), 256, 0, as_stream(stream)>>>(
      c.entries, c.rowptr, c.x, c.N, c.D, seeds, S, e_lo, e_hi, k, picks, out_nid, out_t, out_x);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
