// TGN node memory: the per-batch state machine of TGNMemory with IdentityMessage + LastAggregator.
//
// Replaces (reference tgm-team/tgm @ 5183dc9, tgm/nn/encoder/tgn.py): state :128-133, reset_state
// :149-152/:180-185, forward :157-163, update_state :165-178, _get_updated_memory :192-216,
// _update_msg_store :218-229 (a Python dict with one entry per node, rebuilt by a Python loop per
// unique node per batch), _compute_msg :231-243, train(False) flush :245-251.
//
// Reduction that makes this a fixed-size device state.  A node's message store holds the events
// of the LAST batch it appeared in (as source / as destination).  LastAggregator (:43-56) keeps,
// per node, only the message with the largest float(t) -- the first one in list order on ties,
// source-store messages listed before destination-store messages -- and last_update is the
// largest int t over the node's stored events (:215).  So per node and per store it suffices to
// keep ONE event {other endpoint, t, raw message row} plus the store's largest t:
//     ev_other int32[N] (-1 = empty) | ev_t int64[N] | ev_tmax int64[N] | ev_raw f32[N,D]   (x2)
// Within one batch the winner among a node's events is the one with the largest float(t), the
// earliest position on ties (the reference orders them with an UNSTABLE sort, tgn.py:226, so its
// own choice among exact ties is implementation-defined; parity is claimed for streams where a
// node has no two equal-float(t) events in the same role within a batch).
//
// memory_updater is a GRUCell: two cuBLAS SGEMMs (true fp32) + a fused gate kernel.
#include <cublas_v2.h>

#include <new>

#include "bwd_common.cuh"
#include "common.cuh"

using namespace tgm;

struct tgm_tgn {
  int device = -1;
  int32_t N = 0, D = 0, M = 0, TD = 0, in = 0;
  float *memory = nullptr;       // [N, M]
  int64_t *last_update = nullptr;  // [N]
  struct Store {
    int32_t *other = nullptr;
    int64_t *t = nullptr, *tmax = nullptr;
    float *raw = nullptr;
  } st[2];  // 0: node was the source, 1: node was the destination
  float *Wih = nullptr, *Whh = nullptr, *bih = nullptr, *bhh = nullptr, *tw = nullptr, *tb = nullptr;
  cublasHandle_t blas = nullptr;
  int64_t cap = 0;
  int32_t *rows = nullptr;  // [cap] node id of each workspace row
  float *X = nullptr, *H = nullptr, *GI = nullptr, *GH = nullptr, *newmem = nullptr;
  int64_t *newlu = nullptr;
  // MeanAggregator mode (tgm_tgn_set_aggregator): a node's store is "the events of the last batch
  // it appeared in", so the batches pushed since the last reset/flush are kept in an append-only
  // log and every node remembers where its last batch sits: bstart (-1 = none) / blen, per role
  int aggr = 0;  // 0 = LastAggregator, 1 = MeanAggregator
  int32_t *log_src = nullptr, *log_dst = nullptr;
  int64_t *log_t = nullptr;
  float *log_raw = nullptr;
  int64_t log_cap = 0, log_used = 0;
  int64_t *bstart[2] = {nullptr, nullptr};
  int32_t *blen[2] = {nullptr, nullptr};
  float *S = nullptr;  // [cap, 2*TD] per row: mean of sin(arg) | mean of sin(arg) * dt (backward)
  ~tgm_tgn() {
    if (device >= 0) {
      DeviceGuard g(device);
      cudaFree(log_src), cudaFree(log_dst), cudaFree(log_t), cudaFree(log_raw), cudaFree(S);
      for (int r = 0; r < 2; ++r) cudaFree(bstart[r]), cudaFree(blen[r]);
      cudaFree(memory), cudaFree(last_update);
      for (auto &s : st) cudaFree(s.other), cudaFree(s.t), cudaFree(s.tmax), cudaFree(s.raw);
      for (float *p : {Wih, Whh, bih, bhh, tw, tb, X, H, GI, GH, newmem}) cudaFree(p);
      cudaFree(rows), cudaFree(newlu);
      if (blas) cublasDestroy(blas);
    }
  }
};

namespace {

int blas_fail2(cublasStatus_t s, const char *what) {
  return fail(TGM_ERR_CUDA, std::string("cuBLAS error ") + std::to_string(int(s)) + " in " + what);
}
#define TGN_BLAS(expr)                                             \
  do {                                                             \
    cublasStatus_t _s = (expr);                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) return blas_fail2(_s, #expr); \
  } while (0)

template <typename T>
__global__ void cast_ids_kernel(const T *__restrict__ in, int64_t n, int32_t *__restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x)
    out[i] = int32_t(in[i]);
}

// rows = [src | dst]
__global__ void concat_ids_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b,
                                  int64_t n, int32_t *__restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * n;
       i += int64_t(gridDim.x) * blockDim.x)
    out[i] = i < n ? a[i] : b[i - n];
}

__global__ void iota_kernel(int32_t *out, int64_t n) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x)
    out[i] = int32_t(i);
}

// One warp per workspace row: pick the node's last message, build the GRU input
// X = [mem[v] | mem[other] | raw | Time2Vec(t - last_update[v])] (zeros without a message),
// H = mem[v], and the node's new last_update (tgn.py:192-216, :231-243).
__global__ void __launch_bounds__(256)
tgn_message_kernel(const float *__restrict__ memory, const int64_t *__restrict__ last_update,
                   const int32_t *__restrict__ o_s, const int64_t *__restrict__ t_s,
                   const int64_t *__restrict__ m_s, const float *__restrict__ r_s,
                   const int32_t *__restrict__ o_d, const int64_t *__restrict__ t_d,
                   const int64_t *__restrict__ m_d, const float *__restrict__ r_d,
                   const float *__restrict__ tw, const float *__restrict__ tb, int32_t N, int M,
                   int D, int TD, const int32_t *__restrict__ rows, int64_t n,
                   float *__restrict__ X, float *__restrict__ H, int64_t *__restrict__ newlu) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int in = 2 * M + D + TD;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < n;
       r += int64_t(gridDim.x) * wpb) {
    const int32_t v = rows[r];
    float *x = X + r * in, *h = H + r * M;
    if (v < 0 || v >= N) {  // not a node: inert row
      for (int c = lane; c < in; c += 32) x[c] = 0.f;
      for (int c = lane; c < M; c += 32) h[c] = 0.f;
      if (lane == 0) newlu[r] = 0;
      continue;
    }
    const int32_t os = o_s[v], od = o_d[v];
    const bool has_s = os >= 0, has_d = od >= 0;
    // LastAggregator: largest float(t); source-store messages come first in the list, so the
    // destination-store message wins only when strictly later (tgn.py:50-53)
    const bool use_d = has_d && (!has_s || float(t_d[v]) > float(t_s[v]));
    const bool any = has_s || has_d;
    const int32_t other = use_d ? od : os;
    const int64_t te = use_d ? t_d[v] : (has_s ? t_s[v] : 0);
    const float *raw = (use_d ? r_d : r_s) + int64_t(v) * D;
    const float *mv = memory + int64_t(v) * M;
    for (int c = lane; c < M; c += 32) {
      const float m = mv[c];
      h[c] = m;
      x[c] = any ? m : 0.f;
    }
    if (any) {
      const float *mo = memory + int64_t(other) * M;
      for (int c = lane; c < M; c += 32) x[M + c] = mo[c];
      for (int c = lane; c < D; c += 32) x[2 * M + c] = raw[c];
      const float dt = float(te - last_update[v]);  // t_rel.to(float32) (:239-240)
      for (int c = lane; c < TD; c += 32)
        x[2 * M + D + c] = cosf(__fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c)));
    } else {
      for (int c = M + lane; c < in; c += 32) x[c] = 0.f;
    }
    if (lane == 0) {
      int64_t lu = 0;  // scatter(..., reduce='max') leaves nodes without messages at 0 (:215)
      if (has_s) lu = m_s[v];
      if (has_d) lu = has_s ? (m_d[v] > lu ? m_d[v] : lu) : m_d[v];
      newlu[r] = lu;
    }
  }
}

// GRUCell gates: r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n),
// h' = (1 - z) * n + z * h, with the biases added here.
__global__ void tgn_gru_kernel(const float *__restrict__ GI, const float *__restrict__ GH,
                               const float *__restrict__ bih, const float *__restrict__ bhh,
                               const float *__restrict__ H, int64_t n, int M,
                               float *__restrict__ out) {
  const int64_t total = n * M;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int c = int(i - r * M);
    const float *gi = GI + r * 3 * M, *gh = GH + r * 3 * M;
    const float ir = gi[c] + __ldg(bih + c), hr = gh[c] + __ldg(bhh + c);
    const float iz = gi[M + c] + __ldg(bih + M + c), hz = gh[M + c] + __ldg(bhh + M + c);
    const float in_ = gi[2 * M + c] + __ldg(bih + 2 * M + c);
    const float hn = gh[2 * M + c] + __ldg(bhh + 2 * M + c);
    const float rg = 1.f / (1.f + expf(-(ir + hr)));
    const float zg = 1.f / (1.f + expf(-(iz + hz)));
    const float ng = tanhf(in_ + rg * hn);
    out[i] = (1.f - zg) * ng + zg * H[i];
  }
}

__global__ void tgn_scatter_kernel(const int32_t *__restrict__ rows, int64_t n, int32_t N, int M,
                                   const float *__restrict__ newmem,
                                   const int64_t *__restrict__ newlu, float *__restrict__ memory,
                                   int64_t *__restrict__ last_update) {
  const int64_t total = n * M;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int c = int(i - r * M);
    const int32_t v = rows[r];
    if (v < 0 || v >= N) continue;
    memory[int64_t(v) * M + c] = newmem[i];  // duplicate rows carry identical values
    if (c == 0) last_update[v] = newlu[r];
  }
}

__global__ void tgn_gather_kernel(const int32_t *__restrict__ rows, int64_t n, int32_t N, int M,
                                  const float *__restrict__ memory,
                                  const int64_t *__restrict__ last_update,
                                  float *__restrict__ out_mem, int64_t *__restrict__ out_lu) {
  const int64_t total = n * M;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int c = int(i - r * M);
    const int32_t v = rows[r];
    const bool ok = v >= 0 && v < N;
    out_mem[i] = ok ? memory[int64_t(v) * M + c] : 0.f;
    if (c == 0) out_lu[r] = ok ? last_update[v] : 0;
  }
}

// _update_msg_store for one role: among the batch events of node v = key[p] the winner is the one
// with the largest float(t), earliest position on ties; it replaces v's stored event.
constexpr int kStoreThreads = 256;
__global__ void __launch_bounds__(kStoreThreads)
tgn_store_kernel(const int32_t *__restrict__ key, const int32_t *__restrict__ other,
                 const int64_t *__restrict__ t, const float *__restrict__ raw, int64_t Eb, int32_t N,
                 int D, int32_t *__restrict__ ev_other, int64_t *__restrict__ ev_t,
                 int64_t *__restrict__ ev_tmax, float *__restrict__ ev_raw,
                 int32_t *__restrict__ winner_of) {
  __shared__ int32_t s_key[kStoreThreads];
  __shared__ int64_t s_t[kStoreThreads];
  const int64_t p = int64_t(blockIdx.x) * kStoreThreads + threadIdx.x;
  const int32_t v = p < Eb ? key[p] : -2;
  const int64_t tp = p < Eb ? t[p] : 0;
  const float fp = float(tp);
  bool win = p < Eb && v >= 0 && v < N;
  int64_t tmax = tp;
  for (int64_t base = 0; base < Eb; base += kStoreThreads) {
    const int64_t j = base + threadIdx.x;
    s_key[threadIdx.x] = j < Eb ? key[j] : -3;
    s_t[threadIdx.x] = j < Eb ? t[j] : 0;
    __syncthreads();
    const int lim = int(Eb - base < kStoreThreads ? Eb - base : kStoreThreads);
    for (int u = 0; u < lim; ++u) {
      if (s_key[u] == v) {
        const int64_t tu = s_t[u];
        const float fu = float(tu);
        if (fu > fp || (fu == fp && base + u < p)) win = false;
        if (tu > tmax) tmax = tu;
      }
    }
    __syncthreads();
  }
  if (p < Eb) winner_of[p] = win ? v : -1;
  if (win) {
    ev_other[v] = other[p];
    ev_t[v] = tp;
    ev_tmax[v] = tmax;
  }
  (void)raw, (void)ev_raw, (void)D;
}

// copy the winners' raw message rows (a warp per batch event)
__global__ void __launch_bounds__(256)
tgn_store_raw_kernel(const int32_t *__restrict__ winner_of, const float *__restrict__ raw,
                     int64_t Eb, int D, float *__restrict__ ev_raw) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t p = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); p < Eb;
       p += int64_t(gridDim.x) * wpb) {
    const int32_t v = winner_of[p];
    if (v < 0) continue;
    for (int c = lane; c < D; c += 32) ev_raw[int64_t(v) * D + c] = raw[p * D + c];
  }
}

// ---- MeanAggregator (tgn.py:59-63) ------------------------------------------------------------------
// One warp per workspace row.  The node's messages are the events of its last batch as source
// (role 0) and as destination (role 1) in which it is the key endpoint; the GRU input is the mean
// of [mem[v] | mem[other] | raw | Time2Vec(t - last_update[v])] over them (zeros without any),
// accumulated in list order (source store first) and divided by the count, as scatter(mean) does.
// Lanes own columns, so the read-modify-write of the X row needs no synchronisation.
// S (nullable): per row the means of sin(arg) and sin(arg) * dt, all the backward needs of the
// individual messages (d/dw cos(w dt + b) = -sin(.) dt, d/db = -sin(.)).
__global__ void __launch_bounds__(256)
tgn_message_mean_kernel(const float *__restrict__ memory, const int64_t *__restrict__ last_update,
                        const int32_t *__restrict__ log_src, const int32_t *__restrict__ log_dst,
                        const int64_t *__restrict__ log_t, const float *__restrict__ log_raw,
                        const int64_t *__restrict__ bstart_s, const int32_t *__restrict__ blen_s,
                        const int64_t *__restrict__ bstart_d, const int32_t *__restrict__ blen_d,
                        const float *__restrict__ tw, const float *__restrict__ tb, int32_t N, int M,
                        int D, int TD, const int32_t *__restrict__ rows, int64_t n,
                        float *__restrict__ X, float *__restrict__ H, int64_t *__restrict__ newlu,
                        float *__restrict__ S) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int in = 2 * M + D + TD;
  for (int64_t r = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); r < n;
       r += int64_t(gridDim.x) * wpb) {
    const int32_t v = rows[r];
    float *x = X + r * in, *h = H + r * M;
    float *s = S ? S + r * 2 * TD : nullptr;
    for (int c = lane; c < in; c += 32) x[c] = 0.f;
    if (s) for (int c = lane; c < 2 * TD; c += 32) s[c] = 0.f;
    if (v < 0 || v >= N) {  // not a node: inert row
      for (int c = lane; c < M; c += 32) h[c] = 0.f;
      if (lane == 0) newlu[r] = 0;
      continue;
    }
    const float *mv = memory + int64_t(v) * M;
    for (int c = lane; c < M; c += 32) h[c] = mv[c];
    const int64_t lu_v = last_update[v];
    int count = 0;
    int64_t tmax = 0;
    for (int role = 0; role < 2; ++role) {
      const int64_t st = role ? bstart_d[v] : bstart_s[v];
      if (st < 0) continue;
      const int len = role ? blen_d[v] : blen_s[v];
      const int32_t *key = role ? log_dst : log_src, *oth = role ? log_src : log_dst;
      for (int base = 0; base < len; base += 32) {
        const int p = base + lane;
        unsigned m = __ballot_sync(0xffffffffu, p < len && key[st + p] == v);
        while (m) {
          const int64_t idx = st + base + (__ffs(m) - 1);
          m &= m - 1;
          const int32_t other = oth[idx];
          const int64_t te = log_t[idx];
          const float dt = float(te - lu_v);  // t_rel.to(float32) (:239-240)
          tmax = count == 0 ? te : (te > tmax ? te : tmax);
          ++count;
          const float *mo = memory + int64_t(other) * M;
          for (int c = lane; c < M; c += 32) {
            x[c] += mv[c];
            x[M + c] += mo[c];
          }
          for (int c = lane; c < D; c += 32) x[2 * M + c] += log_raw[idx * D + c];
          for (int c = lane; c < TD; c += 32) {
            const float arg = __fmaf_rn(dt, __ldg(tw + c), __ldg(tb + c));
            x[2 * M + D + c] += cosf(arg);
            if (s) {
              const float sn = sinf(arg);
              s[c] += sn;
              s[TD + c] += sn * dt;
            }
          }
        }
      }
    }
    if (count > 0) {
      const float cnt = float(count);
      for (int c = lane; c < in; c += 32) x[c] = x[c] / cnt;
      if (s) for (int c = lane; c < 2 * TD; c += 32) s[c] = s[c] / cnt;
    }
    if (lane == 0) newlu[r] = count > 0 ? tmax : 0;  // scatter(max) leaves nodes without messages at 0
  }
}

// every endpoint of the batch now points at it: duplicates write identical values
__global__ void tgn_log_mark_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                    int64_t Eb, int32_t N, int64_t off,
                                    int64_t *__restrict__ bstart_s, int32_t *__restrict__ blen_s,
                                    int64_t *__restrict__ bstart_d, int32_t *__restrict__ blen_d) {
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < Eb;
       p += int64_t(gridDim.x) * blockDim.x) {
    const int32_t u = src[p], w = dst[p];
    if (u >= 0 && u < N) bstart_s[u] = off, blen_s[u] = int32_t(Eb);
    if (w >= 0 && w < N) bstart_d[w] = off, blen_d[w] = int32_t(Eb);
  }
}

// Time2Vec gradients from the per-row sums of the mean kernel:
//   gw[c] -= sum_r S[r, TD + c] * d_enc[r, c];   gb[c] -= sum_r S[r, c] * d_enc[r, c]
__global__ void __launch_bounds__(128)
t2v_grad_sums_kernel(const float *__restrict__ S, const float *__restrict__ d_enc, int64_t n, int TD,
                     float *__restrict__ gw, float *__restrict__ gb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= TD) return;
  float aw = 0.f, ab = 0.f;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
    const float g = d_enc[r * TD + c];
    aw = __fmaf_rn(-S[r * 2 * TD + TD + c], g, aw);
    ab = __fmaf_rn(-S[r * 2 * TD + c], g, ab);
  }
  atomicAdd(gw + c, aw);
  atomicAdd(gb + c, ab);
}

int dev_copy2(float **dst, const float *src, size_t n) {
  TGM_CUDA(cudaMalloc(dst, (n ? n : 1) * sizeof(float)));
  if (n) TGM_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyDefault));
  return TGM_OK;
}

int ensure_rows(tgm_tgn *h, int64_t n, cudaStream_t st) {
  if (n <= h->cap) return TGM_OK;
  TGM_CUDA(cudaStreamSynchronize(st));
  for (float **p : {&h->X, &h->H, &h->GI, &h->GH, &h->newmem, &h->S}) {
    cudaFree(*p);
    *p = nullptr;
  }
  cudaFree(h->rows), cudaFree(h->newlu);
  h->rows = nullptr, h->newlu = nullptr, h->cap = 0;
  const size_t cap = size_t(n + n / 4 + 64);
  if (h->aggr == 1) TGM_CUDA(cudaMalloc(&h->S, cap * 2 * h->TD * 4));
  TGM_CUDA(cudaMalloc(&h->rows, cap * 4));
  TGM_CUDA(cudaMalloc(&h->newlu, cap * 8));
  TGM_CUDA(cudaMalloc(&h->X, cap * h->in * 4));
  TGM_CUDA(cudaMalloc(&h->H, cap * h->M * 4));
  TGM_CUDA(cudaMalloc(&h->GI, cap * 3 * h->M * 4));
  TGM_CUDA(cudaMalloc(&h->GH, cap * 3 * h->M * 4));
  TGM_CUDA(cudaMalloc(&h->newmem, cap * h->M * 4));
  h->cap = int64_t(cap);
  return TGM_OK;
}

// _get_updated_memory for the node ids in h->rows[0..n): results in h->newmem / h->newlu
int compute_rows(tgm_tgn *h, int64_t n, cudaStream_t st) {
  const int M = h->M, in = h->in;
  if (h->aggr == 1) {
    tgn_message_mean_kernel<<<grid_for(n, 8, 8), 256, 0, st>>>(
        h->memory, h->last_update, h->log_src, h->log_dst, h->log_t, h->log_raw, h->bstart[0],
        h->blen[0], h->bstart[1], h->blen[1], h->tw, h->tb, h->N, M, h->D, h->TD, h->rows, n, h->X,
        h->H, h->newlu, h->S);
  } else {
    tgn_message_kernel<<<grid_for(n, 8, 8), 256, 0, st>>>(
        h->memory, h->last_update, h->st[0].other, h->st[0].t, h->st[0].tmax, h->st[0].raw,
        h->st[1].other, h->st[1].t, h->st[1].tmax, h->st[1].raw, h->tw, h->tb, h->N, M, h->D, h->TD,
        h->rows, n, h->X, h->H, h->newlu);
  }
  TGM_LAUNCH_CHECK();
  TGN_BLAS(cublasSetStream(h->blas, st));
  const float one = 1.f, zero = 0.f;
  // GI[n,3M] = X[n,in] W_ih[3M,in]^T ; GH[n,3M] = H[n,M] W_hh[3M,M]^T   (row-major)
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, 3 * M, int(n), in, &one, h->Wih, in, h->X,
                       in, &zero, h->GI, 3 * M));
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, 3 * M, int(n), M, &one, h->Whh, M, h->H, M,
                       &zero, h->GH, 3 * M));
  tgn_gru_kernel<<<grid_for(n * M, 256, 8), 256, 0, st>>>(h->GI, h->GH, h->bih, h->bhh, h->H, n, M,
                                                          h->newmem);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

int write_rows(tgm_tgn *h, int64_t n, cudaStream_t st) {
  tgn_scatter_kernel<<<grid_for(n * h->M, 256, 8), 256, 0, st>>>(h->rows, n, h->N, h->M, h->newmem,
                                                                 h->newlu, h->memory,
                                                                 h->last_update);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

int clear_stores(tgm_tgn *h, cudaStream_t st) {
  if (h->aggr == 1) {  // forget every node's last batch; the log starts over
    for (int r = 0; r < 2; ++r) TGM_CUDA(cudaMemsetAsync(h->bstart[r], 0xFF, size_t(h->N) * 8, st));
    h->log_used = 0;
  }
  for (auto &s : h->st) {
    TGM_CUDA(cudaMemsetAsync(s.other, 0xFF, size_t(h->N) * 4, st));  // -1 = empty
    TGM_CUDA(cudaMemsetAsync(s.t, 0, size_t(h->N) * 8, st));
    TGM_CUDA(cudaMemsetAsync(s.tmax, 0, size_t(h->N) * 8, st));
  }
  return TGM_OK;
}

int push_log(tgm_tgn *h, const int32_t *src, const int32_t *dst, const int64_t *t,
             const float *raw, int64_t Eb, cudaStream_t st);

int push_stores(tgm_tgn *h, const int32_t *src, const int32_t *dst, const int64_t *t,
                const float *raw, int64_t Eb, cudaStream_t st) {
  if (h->aggr == 1) return push_log(h, src, dst, t, raw, Eb, st);
  // winner_of scratch: reuse the row-id buffer tail (rows has >= 2*Eb entries here)
  int32_t *winner = h->rows;
  const int grid = int((Eb + kStoreThreads - 1) / kStoreThreads);
  for (int role = 0; role < 2; ++role) {
    const int32_t *key = role ? dst : src, *other = role ? src : dst;
    auto &s = h->st[role];
    tgn_store_kernel<<<grid, kStoreThreads, 0, st>>>(key, other, t, raw, Eb, h->N, h->D, s.other,
                                                     s.t, s.tmax, s.raw, winner);
    TGM_LAUNCH_CHECK();
    if (h->D > 0) {
      tgn_store_raw_kernel<<<grid_for(Eb, 8, 8), 256, 0, st>>>(winner, raw, Eb, h->D, s.raw);
      TGM_LAUNCH_CHECK();
    }
  }
  return TGM_OK;
}

// MeanAggregator: append the batch to the log (grown by doubling: the one synchronising path) and
// point its endpoints at it.
int push_log(tgm_tgn *h, const int32_t *src, const int32_t *dst, const int64_t *t,
             const float *raw, int64_t Eb, cudaStream_t st) {
  const size_t D = size_t(h->D);
  if (h->log_used + Eb > h->log_cap) {
    TGM_CUDA(cudaStreamSynchronize(st));
    int64_t cap = h->log_cap > 0 ? h->log_cap : (int64_t(1) << 16);
    while (cap < h->log_used + Eb) cap *= 2;
    int32_t *ns = nullptr, *nd = nullptr;
    int64_t *nt = nullptr;
    float *nr = nullptr;
    TGM_CUDA(cudaMalloc(&ns, size_t(cap) * 4));
    TGM_CUDA(cudaMalloc(&nd, size_t(cap) * 4));
    TGM_CUDA(cudaMalloc(&nt, size_t(cap) * 8));
    TGM_CUDA(cudaMalloc(&nr, (size_t(cap) * D ? size_t(cap) * D : 1) * 4));
    const size_t used = size_t(h->log_used);
    if (used) {
      TGM_CUDA(cudaMemcpy(ns, h->log_src, used * 4, cudaMemcpyDeviceToDevice));
      TGM_CUDA(cudaMemcpy(nd, h->log_dst, used * 4, cudaMemcpyDeviceToDevice));
      TGM_CUDA(cudaMemcpy(nt, h->log_t, used * 8, cudaMemcpyDeviceToDevice));
      if (D) TGM_CUDA(cudaMemcpy(nr, h->log_raw, used * D * 4, cudaMemcpyDeviceToDevice));
    }
    cudaFree(h->log_src), cudaFree(h->log_dst), cudaFree(h->log_t), cudaFree(h->log_raw);
    h->log_src = ns, h->log_dst = nd, h->log_t = nt, h->log_raw = nr, h->log_cap = cap;
  }
  const int64_t off = h->log_used;
  const size_t n = size_t(Eb);
  TGM_CUDA(cudaMemcpyAsync(h->log_src + off, src, n * 4, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(h->log_dst + off, dst, n * 4, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(h->log_t + off, t, n * 8, cudaMemcpyDeviceToDevice, st));
  if (D) TGM_CUDA(cudaMemcpyAsync(h->log_raw + size_t(off) * D, raw, n * D * 4,
                                  cudaMemcpyDeviceToDevice, st));
  tgn_log_mark_kernel<<<grid_for(Eb, 256, 8), 256, 0, st>>>(src, dst, Eb, h->N, off, h->bstart[0],
                                                            h->blen[0], h->bstart[1], h->blen[1]);
  TGM_LAUNCH_CHECK();
  h->log_used += Eb;
  return TGM_OK;
}

}  // namespace

extern "C" int tgm_tgn_create(tgm_tgn **out, int32_t num_nodes, int32_t raw_msg_dim,
                              int32_t memory_dim, int32_t time_dim, const float *gru_w_ih,
                              const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                              const float *t2v_w, const float *t2v_b, int device) {
  TGM_REQUIRE(out != nullptr, "tgm_tgn_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(num_nodes > 0 && raw_msg_dim >= 0 && memory_dim > 0 && time_dim > 0,
              "tgm_tgn_create: bad sizes");
  TGM_REQUIRE(gru_w_ih && gru_w_hh && gru_b_ih && gru_b_hh && t2v_w && t2v_b,
              "tgm_tgn_create: NULL parameter");
  TGM_REQUIRE(device >= 0, "tgm_tgn_create: a CUDA device is required (no CPU fallback)");
  DeviceGuard g(device);
  if (!g.ok) return fail(TGM_ERR_CUDA, "tgm_tgn_create: cannot select device");
  tgm_tgn *h = new (std::nothrow) tgm_tgn();
  if (!h) return fail(TGM_ERR_OOM, "tgm_tgn_create: host allocation failed");
  h->device = device;
  h->N = num_nodes, h->D = raw_msg_dim, h->M = memory_dim, h->TD = time_dim;
  h->in = raw_msg_dim + 2 * memory_dim + time_dim;  // IdentityMessage.out_channels (:69)
  const size_t N = size_t(num_nodes), M = size_t(memory_dim), D = size_t(raw_msg_dim);
  cudaError_t e = cudaMalloc(&h->memory, N * M * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->last_update, N * 8);
  for (auto &s : h->st) {
    if (e == cudaSuccess) e = cudaMalloc(&s.other, N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.t, N * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.tmax, N * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.raw, (N * D ? N * D : 1) * 4);
  }
  int rc = e == cudaSuccess ? TGM_OK : cuda_fail(e, "TGN state allocation", __FILE__, __LINE__);
  if (!rc) rc = dev_copy2(&h->Wih, gru_w_ih, 3 * M * size_t(h->in));
  if (!rc) rc = dev_copy2(&h->Whh, gru_w_hh, 3 * M * M);
  if (!rc) rc = dev_copy2(&h->bih, gru_b_ih, 3 * M);
  if (!rc) rc = dev_copy2(&h->bhh, gru_b_hh, 3 * M);
  if (!rc) rc = dev_copy2(&h->tw, t2v_w, size_t(time_dim));
  if (!rc) rc = dev_copy2(&h->tb, t2v_b, size_t(time_dim));
  if (!rc) {
    cublasStatus_t s = cublasCreate(&h->blas);
    if (s != CUBLAS_STATUS_SUCCESS) rc = blas_fail2(s, "cublasCreate");
    else cublasSetMathMode(h->blas, CUBLAS_PEDANTIC_MATH);
  }
  if (!rc) rc = tgm_tgn_reset(h, nullptr);
  if (!rc) {
    e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) rc = cuda_fail(e, "TGN reset", __FILE__, __LINE__);
  }
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return TGM_OK;
}

extern "C" void tgm_tgn_destroy(tgm_tgn *h) { delete h; }

extern "C" int tgm_tgn_reset(tgm_tgn *h, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_reset: handle is NULL");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  TGM_CUDA(cudaMemsetAsync(h->memory, 0, size_t(h->N) * h->M * 4, st));
  TGM_CUDA(cudaMemsetAsync(h->last_update, 0, size_t(h->N) * 8, st));
  return clear_stores(h, st);
}

extern "C" int tgm_tgn_state(const tgm_tgn *h, float **memory, int64_t **last_update) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_state: handle is NULL");
  if (memory) *memory = h->memory;
  if (last_update) *last_update = h->last_update;
  return TGM_OK;
}

extern "C" int tgm_tgn_forward(tgm_tgn *h, const int64_t *n_id, int64_t n, int training,
                               float *out_memory, int64_t *out_last_update, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_forward: handle is NULL");
  TGM_REQUIRE(n >= 0, "tgm_tgn_forward: n must be >= 0");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(n_id && out_memory && out_last_update, "tgm_tgn_forward: NULL array argument");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  int rc = ensure_rows(h, n, st);
  if (rc) return rc;
  cast_ids_kernel<int64_t><<<grid_for(n, 256, 8), 256, 0, st>>>(n_id, n, h->rows);
  TGM_LAUNCH_CHECK();
  if (training) {  // updated memory WITHOUT writing it (tgn.py:158-159)
    rc = compute_rows(h, n, st);
    if (rc) return rc;
    TGM_CUDA(cudaMemcpyAsync(out_memory, h->newmem, size_t(n) * h->M * 4, cudaMemcpyDeviceToDevice, st));
    TGM_CUDA(cudaMemcpyAsync(out_last_update, h->newlu, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    tgn_gather_kernel<<<grid_for(n * h->M, 256, 8), 256, 0, st>>>(
        h->rows, n, h->N, h->M, h->memory, h->last_update, out_memory, out_last_update);
    TGM_LAUNCH_CHECK();
  }
  return TGM_OK;
}

extern "C" int tgm_tgn_update_state(tgm_tgn *h, const int32_t *src, const int32_t *dst,
                                    const int64_t *t, const float *raw_msg, int64_t Eb,
                                    int training, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_update_state: handle is NULL");
  TGM_REQUIRE(Eb >= 0, "tgm_tgn_update_state: Eb must be >= 0");
  if (Eb == 0) return TGM_OK;
  TGM_REQUIRE(src && dst && t && (raw_msg || h->D == 0),
              "tgm_tgn_update_state: NULL array argument");
  TGM_REQUIRE(Eb <= (int64_t(1) << 16),
              "tgm_tgn_update_state: batch too large (max 65536 events per update)");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  int rc = ensure_rows(h, 2 * Eb, st);
  if (rc) return rc;
  auto update_memory = [&]() -> int {
    // _update_memory(unique(cat[src, dst])): every row is computed from the pre-update snapshot and
    // duplicates write identical values, so no unique() is needed
    concat_ids_kernel<<<grid_for(2 * Eb, 256, 8), 256, 0, st>>>(src, dst, Eb, h->rows);
    TGM_LAUNCH_CHECK();
    int r = compute_rows(h, 2 * Eb, st);
    if (r) return r;
    return write_rows(h, 2 * Eb, st);
  };
  if (training) {  // tgn.py:170-173
    rc = update_memory();
    if (!rc) rc = push_stores(h, src, dst, t, raw_msg, Eb, st);
  } else {  // tgn.py:174-177
    rc = push_stores(h, src, dst, t, raw_msg, Eb, st);
    if (!rc) rc = update_memory();
  }
  return rc;
}

extern "C" int tgm_tgn_flush(tgm_tgn *h, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_flush: handle is NULL");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  int rc = ensure_rows(h, h->N, st);
  if (rc) return rc;
  iota_kernel<<<grid_for(h->N, 256, 8), 256, 0, st>>>(h->rows, h->N);
  TGM_LAUNCH_CHECK();
  rc = compute_rows(h, h->N, st);
  if (!rc) rc = write_rows(h, h->N, st);
  if (!rc) rc = clear_stores(h, st);
  return rc;
}

// ---- training: parameter refresh, forward with saved rows, backward ------------------------------
// The reference runs loss.backward() AFTER memory.update_state() (examples/linkproppred/tgn.py:
// 111-118): autograd holds the tensors of memory(n_id)'s graph while the state moves on.  The
// device equivalent: tgm_tgn_forward_saved hands the caller the GRU inputs of its rows (X, H and
// {t - last_update, has-a-message}); tgm_tgn_backward is a pure function of those rows, the
// current parameters and d_memory.  memory[...] and the raw messages carry no gradient in the
// reference (buffers, tgn.py:128-133, detached every step :154-155), so the parameters that
// receive one are the GRU cell's and, through the message's time-encoding columns, Time2Vec's.
namespace {

__global__ void tgn_aux_kernel(const int64_t *__restrict__ last_update,
                               const int32_t *__restrict__ o_s, const int64_t *__restrict__ t_s,
                               const int32_t *__restrict__ o_d, const int64_t *__restrict__ t_d,
                               int32_t N, const int32_t *__restrict__ rows, int64_t n,
                               float *__restrict__ aux) {
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n;
       r += int64_t(gridDim.x) * blockDim.x) {
    const int32_t v = rows[r];
    float dt = 0.f, has = 0.f;
    if (v >= 0 && v < N) {
      // the same pick as tgn_message_kernel
      const bool has_s = o_s[v] >= 0, has_d = o_d[v] >= 0;
      const bool use_d = has_d && (!has_s || float(t_d[v]) > float(t_s[v]));
      if (has_s || has_d) {
        const int64_t te = use_d ? t_d[v] : t_s[v];
        dt = float(te - last_update[v]);
        has = 1.f;
      }
    }
    aux[2 * r] = dt;
    aux[2 * r + 1] = has;
  }
}

// In place: GI/GH hold the gate pre-activations without bias on entry, d(gi)/d(gh) on exit.
//   h' = (1 - z) n + z h;  d n = d h' (1 - z);  d z = d h' (h - n);  d pre_n = d n (1 - n^2)
//   d pre_r = d pre_n gh_n r (1 - r);  d pre_z = d z z (1 - z);  d gh_n = d pre_n r
__global__ void tgn_gru_bwd_kernel(float *__restrict__ GI, float *__restrict__ GH,
                                   const float *__restrict__ bih, const float *__restrict__ bhh,
                                   const float *__restrict__ H, const float *__restrict__ d_out,
                                   int64_t n, int M) {
  const int64_t total = n * M;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    const int c = int(i - r * M);
    float *gi = GI + r * 3 * M, *gh = GH + r * 3 * M;
    const float ir = gi[c] + __ldg(bih + c), hr = gh[c] + __ldg(bhh + c);
    const float iz = gi[M + c] + __ldg(bih + M + c), hz = gh[M + c] + __ldg(bhh + M + c);
    const float in_ = gi[2 * M + c] + __ldg(bih + 2 * M + c);
    const float hn = gh[2 * M + c] + __ldg(bhh + 2 * M + c);
    const float rg = 1.f / (1.f + expf(-(ir + hr)));
    const float zg = 1.f / (1.f + expf(-(iz + hz)));
    const float ng = tanhf(in_ + rg * hn);
    const float go = d_out[i];
    const float d_pn = go * (1.f - zg) * (1.f - ng * ng);
    const float d_pz = go * (H[i] - ng) * zg * (1.f - zg);
    const float d_pr = d_pn * hn * rg * (1.f - rg);
    gi[c] = d_pr, gi[M + c] = d_pz, gi[2 * M + c] = d_pn;
    gh[c] = d_pr, gh[M + c] = d_pz, gh[2 * M + c] = d_pn * rg;
  }
}

}  // namespace

extern "C" int tgm_tgn_set_params(tgm_tgn *h, const float *gru_w_ih, const float *gru_w_hh,
                                  const float *gru_b_ih, const float *gru_b_hh, const float *t2v_w,
                                  const float *t2v_b, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_set_params: handle is NULL");
  TGM_REQUIRE(gru_w_ih && gru_w_hh && gru_b_ih && gru_b_hh && t2v_w && t2v_b,
              "tgm_tgn_set_params: NULL parameter");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  const size_t M = size_t(h->M);
  TGM_CUDA(cudaMemcpyAsync(h->Wih, gru_w_ih, 3 * M * size_t(h->in) * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(h->Whh, gru_w_hh, 3 * M * M * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(h->bih, gru_b_ih, 3 * M * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(h->bhh, gru_b_hh, 3 * M * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(h->tw, t2v_w, size_t(h->TD) * 4, cudaMemcpyDefault, st));
  TGM_CUDA(cudaMemcpyAsync(h->tb, t2v_b, size_t(h->TD) * 4, cudaMemcpyDefault, st));
  return TGM_OK;
}

extern "C" int tgm_tgn_forward_saved(tgm_tgn *h, const int64_t *n_id, int64_t n,
                                     float *out_memory, int64_t *out_last_update, float *saved_x,
                                     float *saved_h, float *saved_aux, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_forward_saved: handle is NULL");
  TGM_REQUIRE(n >= 0, "tgm_tgn_forward_saved: n must be >= 0");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(n_id && out_memory && out_last_update && saved_x && saved_h && saved_aux,
              "tgm_tgn_forward_saved: NULL array argument");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  int rc = ensure_rows(h, n, st);
  if (rc) return rc;
  cast_ids_kernel<int64_t><<<grid_for(n, 256, 8), 256, 0, st>>>(n_id, n, h->rows);
  TGM_LAUNCH_CHECK();
  rc = compute_rows(h, n, st);
  if (rc) return rc;
  const size_t rows = size_t(n);
  TGM_CUDA(cudaMemcpyAsync(out_memory, h->newmem, rows * h->M * 4, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(out_last_update, h->newlu, rows * 8, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(saved_x, h->X, rows * h->in * 4, cudaMemcpyDeviceToDevice, st));
  TGM_CUDA(cudaMemcpyAsync(saved_h, h->H, rows * h->M * 4, cudaMemcpyDeviceToDevice, st));
  if (h->aggr == 1) {  // MeanAggregator: the per-row sin sums written by the message kernel
    TGM_CUDA(cudaMemcpyAsync(saved_aux, h->S, rows * 2 * h->TD * 4, cudaMemcpyDeviceToDevice, st));
    return TGM_OK;
  }
  tgn_aux_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(h->last_update, h->st[0].other, h->st[0].t,
                                                      h->st[1].other, h->st[1].t, h->N, h->rows, n,
                                                      saved_aux);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_tgn_backward(tgm_tgn *h, const float *saved_x, const float *saved_h,
                                const float *saved_aux, int64_t n, const float *d_memory,
                                float *g_w_ih, float *g_w_hh, float *g_b_ih, float *g_b_hh,
                                float *g_t2v_w, float *g_t2v_b, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_backward: handle is NULL");
  TGM_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "tgm_tgn_backward: bad n");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(saved_x && saved_h && saved_aux && d_memory, "tgm_tgn_backward: NULL input");
  TGM_REQUIRE(g_w_ih && g_w_hh && g_b_ih && g_b_hh && g_t2v_w && g_t2v_b,
              "tgm_tgn_backward: NULL gradient buffer");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  int rc = ensure_rows(h, n, st);
  if (rc) return rc;
  const int M = h->M, in = h->in, TD = h->TD, M3 = 3 * h->M;
  const float one = 1.f, zero = 0.f;
  TGN_BLAS(cublasSetStream(h->blas, st));
  // recompute the gate pre-activations from the saved rows (as compute_rows)
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, M3, int(n), in, &one, h->Wih, in, saved_x,
                       in, &zero, h->GI, M3));
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, M3, int(n), M, &one, h->Whh, M, saved_h, M,
                       &zero, h->GH, M3));
  tgn_gru_bwd_kernel<<<grid_for(n * M, 256, 8), 256, 0, st>>>(h->GI, h->GH, h->bih, h->bhh, saved_h,
                                                              d_memory, n, M);
  TGM_LAUNCH_CHECK();
  // g_w_ih[3M,in] += dGI^T X ; g_w_hh[3M,M] += dGH^T H   (row-major)
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, in, M3, int(n), &one, saved_x, in, h->GI,
                       M3, &one, g_w_ih, in));
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, M, M3, int(n), &one, saved_h, M, h->GH, M3,
                       &one, g_w_hh, M));
  colsum_add_kernel<<<colsum_grid(n, M3), 128, 0, st>>>(h->GI, n, M3, M3, g_b_ih);
  TGM_LAUNCH_CHECK();
  colsum_add_kernel<<<colsum_grid(n, M3), 128, 0, st>>>(h->GH, n, M3, M3, g_b_hh);
  TGM_LAUNCH_CHECK();
  // d(time encoding)[n,TD] = dGI[n,3M] W_ih[:, 2M+D:]   (into the X workspace), then Time2Vec
  TGN_BLAS(cublasSgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, TD, int(n), M3, &one,
                       h->Wih + (2 * M + h->D), in, h->GI, M3, &zero, h->X, TD));
  if (h->aggr == 1)
    t2v_grad_sums_kernel<<<colsum_grid(n, TD), 128, 0, st>>>(saved_aux, h->X, n, TD, g_t2v_w, g_t2v_b);
  else
    t2v_grad_kernel<<<colsum_grid(n, TD), 128, 0, st>>>(saved_aux, saved_aux + 1, 2, h->X, TD, n, TD,
                                                        h->tw, h->tb, g_t2v_w, g_t2v_b);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_tgn_set_aggregator(tgm_tgn *h, int kind, int64_t log_capacity, tgm_stream stream) {
  TGM_REQUIRE(h != nullptr, "tgm_tgn_set_aggregator: handle is NULL");
  TGM_REQUIRE(kind == TGM_TGN_AGGR_LAST || kind == TGM_TGN_AGGR_MEAN,
              "tgm_tgn_set_aggregator: kind must be TGM_TGN_AGGR_LAST or TGM_TGN_AGGR_MEAN");
  TGM_REQUIRE(log_capacity >= 0, "tgm_tgn_set_aggregator: log_capacity must be >= 0");
  DeviceGuard g(h->device);
  cudaStream_t st = as_stream(stream);
  TGM_CUDA(cudaStreamSynchronize(st));
  if (kind == TGM_TGN_AGGR_MEAN && h->bstart[0] == nullptr) {
    for (int r = 0; r < 2; ++r) {
      TGM_CUDA(cudaMalloc(&h->bstart[r], size_t(h->N) * 8));
      TGM_CUDA(cudaMalloc(&h->blen[r], size_t(h->N) * 4));
      TGM_CUDA(cudaMemset(h->blen[r], 0, size_t(h->N) * 4));
    }
  }
  if (kind == TGM_TGN_AGGR_MEAN && log_capacity > h->log_cap) {
    cudaFree(h->log_src), cudaFree(h->log_dst), cudaFree(h->log_t), cudaFree(h->log_raw);
    h->log_src = h->log_dst = nullptr, h->log_t = nullptr, h->log_raw = nullptr, h->log_cap = 0;
    const size_t cap = size_t(log_capacity), D = size_t(h->D);
    TGM_CUDA(cudaMalloc(&h->log_src, cap * 4));
    TGM_CUDA(cudaMalloc(&h->log_dst, cap * 4));
    TGM_CUDA(cudaMalloc(&h->log_t, cap * 8));
    TGM_CUDA(cudaMalloc(&h->log_raw, (cap * D ? cap * D : 1) * 4));
    h->log_cap = log_capacity;
  }
  // the row workspaces are re-created on the next call (the mean mode adds one)
  for (float **p : {&h->X, &h->H, &h->GI, &h->GH, &h->newmem, &h->S}) cudaFree(*p), *p = nullptr;
  cudaFree(h->rows), cudaFree(h->newlu);
  h->rows = nullptr, h->newlu = nullptr, h->cap = 0;
  h->aggr = kind;
  return tgm_tgn_reset(h, stream);  // memory, last_update and the message stores start empty
}

extern "C" int tgm_tgn_saved_aux_width(const tgm_tgn *h) {
  if (h == nullptr) return fail(TGM_ERR_INVALID, "tgm_tgn_saved_aux_width: handle is NULL");
  return h->aggr == 1 ? 2 * h->TD : 2;
}
