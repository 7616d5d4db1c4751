/*
 * CPU ORACLE (test infrastructure, NOT product code): plain-C restatement of the reference's
 * recency neighbor sampler, used (a) by tests/ as the full-size checker, (b) by bench.py as the
 * timed CPU baseline.  Nothing under tgm_b200/ links or loads this file.
 *
 * Reference = tgm-team/tgm @ 5183dc9, tgm/hooks/neighbors/recency.py:
 *   state            :93-97, :410-416   ids int32[N,B], times int64[N,B], feats f32[N,B,D],
 *                                       write_pos int32[N]
 *   reset_state      :111-117
 *   query            :239-321           (_get_recency_neighbors)
 *   update           :323-399           (_update; ideal semantics of the sort key, i.e. the
 *                                        reference's result whenever N*(t_max+1) < 2^31)
 *   hook call        :119-171           all hops, then the push
 *
 * Parity pinning: tests/test_oracle_golden.py runs this file against every fixture under
 * tests/golden/ (outputs of the unmodified reference) and the known answers of the
 * reference's own unit tests.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC) -> oracle/_build/librecency_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PAD (-1) /* tgm/constants.py:3 */

typedef struct {
  int32_t N, B, D;
  int32_t *ids;   /* [N,B] */
  int64_t *times; /* [N,B] */
  float *feats;   /* [N,B,D] */
  int32_t *wpos;  /* [N] */
  /* update scratch */
  int64_t cap;
  int32_t *perm, *tmp;
} ring_t;

void ring_reset(ring_t *r) { /* recency.py:111-117 */
  size_t nb = (size_t)r->N * r->B;
  for (size_t i = 0; i < nb; ++i) r->ids[i] = PAD;
  memset(r->times, 0, nb * sizeof(int64_t));
  if (r->D) memset(r->feats, 0, nb * r->D * sizeof(float));
  memset(r->wpos, 0, (size_t)r->N * sizeof(int32_t));
}

ring_t *ring_create(int32_t N, int32_t B, int32_t D) {
  ring_t *r = (ring_t *)calloc(1, sizeof(ring_t));
  if (!r) return NULL;
  r->N = N, r->B = B, r->D = D;
  size_t nb = (size_t)N * B;
  r->ids = (int32_t *)malloc(nb * sizeof(int32_t));
  r->times = (int64_t *)malloc(nb * sizeof(int64_t));
  r->feats = (float *)malloc((nb * (D ? D : 1)) * sizeof(float));
  r->wpos = (int32_t *)malloc((size_t)N * sizeof(int32_t));
  if (!r->ids || !r->times || !r->feats || !r->wpos) return NULL;
  ring_reset(r);
  return r;
}

void ring_destroy(ring_t *r) {
  if (!r) return;
  free(r->ids), free(r->times), free(r->feats), free(r->wpos), free(r->perm), free(r->tmp);
  free(r);
}

void ring_state(ring_t *r, int32_t **ids, int64_t **times, float **feats, int32_t **wpos) {
  *ids = r->ids, *times = r->times, *feats = r->feats, *wpos = r->wpos;
}

/* recency.py:239-321.  Unrolled column j (oldest..newest) is ring slot (wp - (B - j)) mod B
 * (:263-264); `last` = right-most column with id != -1 and time < tq (:267-281); the output is
 * the k columns ending at `last`, right-aligned, the rest padding (:287-319). */
static void query_one(const ring_t *r, int32_t seed, int64_t tq, int k, int32_t *o_nid,
                      int64_t *o_t, float *o_x) {
  const int B = r->B, D = r->D;
  int64_t row = seed;
  if (row < 0) row += r->N; /* torch negative indexing (:256) */
  const int32_t *ids = r->ids + row * B;
  const int64_t *tt = r->times + row * B;
  int64_t wp = (int64_t)r->wpos[row];
  int last = -1;
  for (int j = 0; j < B; ++j) {
    int slot = (int)((((wp - (B - j)) % B) + B) % B);
    if (ids[slot] != PAD && tt[slot] < tq) last = j;
  }
  for (int c = 0; c < k; ++c) {
    int col = last - (k - 1 - c);
    if (col >= 0) {
      int slot = (int)((((wp - (B - col)) % B) + B) % B);
      o_nid[c] = ids[slot];
      o_t[c] = tt[slot];
      if (D) memcpy(o_x + (size_t)c * D, r->feats + ((size_t)row * B + slot) * D, D * sizeof(float));
    } else {
      o_nid[c] = PAD;
      o_t[c] = 0;
      if (D) memset(o_x + (size_t)c * D, 0, D * sizeof(float));
    }
  }
}

void ring_query(const ring_t *r, const int32_t *seeds, const int64_t *tq, int64_t S, int32_t k,
                int32_t *out_nid, int64_t *out_t, float *out_x) {
  const int D = r->D;
#pragma omp parallel for schedule(static) if (S >= 4096)
  for (int64_t s = 0; s < S; ++s)
    query_one(r, seeds[s], tq[s], k, out_nid + s * k, out_t + s * k,
              out_x ? out_x + (size_t)s * k * D : NULL);
}

/* recency.py:323-399.  Entries = [src->dst for every edge] ++ [dst->src for every edge]
 * (:339-342; directed: the first half only, :331-336); stable sort by (node, time) (:347-349);
 * keep the last B of each node (:373); write at (write_pos + j) % B (:389-394); write_pos +=
 * number written (:397-399). */
static const int32_t *g_node;
static const int64_t *g_time;
static void merge_sort(int32_t *a, int32_t *tmp, int64_t n) { /* stable */
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort(a, tmp, h);
  merge_sort(a + h, tmp, n - h);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) {
    int32_t x = a[i], y = a[j];
    int less_y = (g_node[y] < g_node[x]) || (g_node[y] == g_node[x] && g_time[y] < g_time[x]);
    tmp[o++] = less_y ? a[j++] : a[i++];
  }
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(int32_t));
}

int ring_update(ring_t *r, const int32_t *src, const int32_t *dst, const int64_t *t,
                const float *x, int64_t Eb, int directed) {
  const int B = r->B, D = r->D;
  const int64_t n = directed ? Eb : 2 * Eb;
  if (n == 0) return 0;
  if (n > r->cap) {
    free(r->perm), free(r->tmp);
    r->cap = n + n / 2;
    r->perm = (int32_t *)malloc((size_t)r->cap * sizeof(int32_t));
    r->tmp = (int32_t *)malloc((size_t)r->cap * sizeof(int32_t));
    if (!r->perm || !r->tmp) return -1;
  }
  int32_t *node = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  int64_t *tt = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  if (!node || !tt) return -1;
  for (int64_t i = 0; i < n; ++i) {
    int64_t e = i < Eb ? i : i - Eb;
    node[i] = i < Eb ? src[e] : dst[e];
    tt[i] = t[e];
    r->perm[i] = (int32_t)i;
  }
  g_node = node, g_time = tt;
  merge_sort(r->perm, r->tmp, n);
  for (int64_t a = 0; a < n;) {
    int64_t b = a;
    const int32_t v = node[r->perm[a]];
    while (b < n && node[r->perm[b]] == v) ++b;
    const int64_t cnt = b - a, first = cnt > B ? cnt - B : 0;
    const int64_t wp = r->wpos[v];
    for (int64_t j = first; j < cnt; ++j) {
      const int64_t i = r->perm[a + j], e = i < Eb ? i : i - Eb;
      const int slot = (int)((((wp + (j - first)) % B) + B) % B);
      const size_t at = (size_t)v * B + slot;
      r->ids[at] = i < Eb ? dst[e] : src[e];
      r->times[at] = t[e];
      if (D) {
        if (x) memcpy(r->feats + at * D, x + (size_t)e * D, D * sizeof(float));
        else memset(r->feats + at * D, 0, D * sizeof(float)); /* :325-329 */
      }
    }
    r->wpos[v] = (int32_t)(wp + (cnt - first));
    a = b;
  }
  free(node), free(tt);
  return 0;
}

/* Position-sensitive 64-bit checksums (wrapping arithmetic) of an output block whose first
 * element has global index `base`: sum_i v_i * ((base+i) * 0x9E3779B97F4A7C15 + 1).  The GPU
 * tests compute the same sums with torch int64 ops. */
#define GOLD 0x9E3779B97F4A7C15ull
static uint64_t csum_i32(const int32_t *v, int64_t n, uint64_t base) {
  uint64_t s = 0;
  for (int64_t i = 0; i < n; ++i) s += (uint64_t)(int64_t)v[i] * ((base + (uint64_t)i) * GOLD + 1ull);
  return s;
}
static uint64_t csum_i64(const int64_t *v, int64_t n, uint64_t base) {
  uint64_t s = 0;
  for (int64_t i = 0; i < n; ++i) s += (uint64_t)v[i] * ((base + (uint64_t)i) * GOLD + 1ull);
  return s;
}

/* One pass of the loader + hook over edges [e_lo, e_hi) of a stream (recency.py:119-171 driven
 * by tgm/data/loader.py:136-160): per batch of `bs` edges, seeds = [src | dst] with the edge
 * times, every hop queried (hop h>0 seeds = flattened hop h-1 neighbours, :141-143), then the
 * batch is pushed.  Outputs are not kept; per hop h the running checksums of (nid, time, feature
 * bit patterns) are accumulated into csum[3*h .. 3*h+2], rows numbered globally in batch order.
 * If out_nid/out_t/out_x are non-NULL the hop-0 outputs are also stored ([2*(e_hi-e_lo), k0]).
 * Returns the number of sampled slots (sum over hops of S_h * k_h), <0 on allocation failure. */
int64_t ring_run_stream(ring_t *r, const int32_t *src, const int32_t *dst, const int64_t *t,
                        const float *x, int64_t e_lo, int64_t e_hi, int64_t bs,
                        const int32_t *num_nbrs, int32_t nhops, int directed, uint64_t *csum,
                        int32_t *out_nid, int64_t *out_t, float *out_x) {
  const int D = r->D;
  int64_t smax = 2 * bs, slots = 0;
  int64_t *rows_done = (int64_t *)calloc((size_t)nhops, sizeof(int64_t));
  int32_t **nid = (int32_t **)calloc((size_t)nhops, sizeof(void *));
  int64_t **nt = (int64_t **)calloc((size_t)nhops, sizeof(void *));
  float **nx = (float **)calloc((size_t)nhops, sizeof(void *));
  int32_t *seed0 = (int32_t *)malloc((size_t)smax * sizeof(int32_t));
  int64_t *tq0 = (int64_t *)malloc((size_t)smax * sizeof(int64_t));
  if (!rows_done || !nid || !nt || !nx || !seed0 || !tq0) return -1;
  {
    int64_t s = smax;
    for (int h = 0; h < nhops; ++h) {
      size_t cells = (size_t)s * num_nbrs[h];
      nid[h] = (int32_t *)malloc(cells * sizeof(int32_t));
      nt[h] = (int64_t *)malloc(cells * sizeof(int64_t));
      nx[h] = (float *)malloc((cells * (D ? D : 1)) * sizeof(float));
      if (!nid[h] || !nt[h] || !nx[h]) return -1;
      s = (int64_t)cells;
    }
  }
  if (csum) memset(csum, 0, (size_t)nhops * 3 * sizeof(uint64_t));
  for (int64_t lo = e_lo; lo < e_hi; lo += bs) {
    const int64_t hi = lo + bs < e_hi ? lo + bs : e_hi, nb = hi - lo;
    for (int64_t i = 0; i < nb; ++i) {
      seed0[i] = src[lo + i], seed0[nb + i] = dst[lo + i];
      tq0[i] = tq0[nb + i] = t[lo + i];
    }
    const int32_t *seeds = seed0;
    const int64_t *tq = tq0;
    int64_t S = 2 * nb;
    for (int h = 0; h < nhops; ++h) {
      const int k = num_nbrs[h];
      ring_query(r, seeds, tq, S, k, nid[h], nt[h], nx[h]);
      const int64_t cells = S * k;
      if (csum) {
        const uint64_t base = (uint64_t)rows_done[h] * (uint64_t)k;
        csum[3 * h + 0] += csum_i32(nid[h], cells, base);
        csum[3 * h + 1] += csum_i64(nt[h], cells, base);
        if (D) csum[3 * h + 2] += csum_i32((const int32_t *)nx[h], cells * D, base * (uint64_t)D);
      }
      if (h == 0 && out_nid) {
        const size_t at = (size_t)rows_done[0] * k;
        memcpy(out_nid + at, nid[0], (size_t)cells * sizeof(int32_t));
        memcpy(out_t + at, nt[0], (size_t)cells * sizeof(int64_t));
        if (D && out_x) memcpy(out_x + at * D, nx[0], (size_t)cells * D * sizeof(float));
      }
      rows_done[h] += S;
      slots += cells;
      seeds = nid[h], tq = nt[h], S = cells;
    }
    if (ring_update(r, src + lo, dst + lo, t + lo, x ? x + (size_t)lo * D : NULL, nb, directed))
      return -1;
  }
  for (int h = 0; h < nhops; ++h) free(nid[h]), free(nt[h]), free(nx[h]);
  free(nid), free(nt), free(nx), free(seed0), free(tq0), free(rows_done);
  return slots;
}

/* examples/linkproppred/graphmixer.py:131-135: sum_c z[s,c,:]*mask / max(1, #valid), fp32, the
 * k terms added left to right. */
void masked_mean_ref(const float *z, const int32_t *nid, int64_t S, int32_t k, int32_t D,
                     float *out) {
  for (int64_t s = 0; s < S; ++s) {
    int cnt = 0;
    for (int c = 0; c < k; ++c) cnt += nid[s * k + c] != PAD;
    const float den = (float)(cnt > 1 ? cnt : 1);
    for (int d = 0; d < D; ++d) {
      volatile float acc = 0.f; /* volatile: no reassociation / vector reduction reorder */
      for (int c = 0; c < k; ++c) {
        const float m = nid[s * k + c] != PAD ? 1.f : 0.f;
        acc = acc + z[((size_t)s * k + c) * D + d] * m;
      }
      out[(size_t)s * D + d] = acc / den;
    }
  }
}
