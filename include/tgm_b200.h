/*
 * tgm_b200.h -- C ABI of the B200-native temporal neighbor-sampling engine.
 *
 * The reference (tgm-team/tgm @ 5183dc9) is pure Python and has no FFI of its own; its plug-in
 * points for this path are Python protocols.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  The Python face that TGM
 * calls (DGStorageBase subclass, DGHook object) lives in tgm_b200/ and binds this library with
 * ctypes; INTEGRATION.md shows the stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns TGM_OK (0) or a negative TGM_ERR_* code; tgm_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - plain pointers and sizes only.  Unless a parameter says "host", array arguments are DEVICE
 *     pointers owned by the caller and must stay valid until the stream operation completes.
 *   - all work is stream-ordered on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); no call synchronises the device except *_create / *_destroy / *_host.
 *   - handles are not thread-safe; one handle = one device.
 *   - node ids are int32, timestamps int64, edge features float32 (tgm/data/dg_data.py:129-161).
 *   - TGM_PADDED_NODE_ID (-1) marks padded neighbour slots (tgm/constants.py:3); padded
 *     timestamps are 0 and padded features 0.0f (tgm/hooks/neighbors/recency.py:317-319).
 */
#ifndef TGM_B200_H
#define TGM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGM_OK 0
#define TGM_ERR_INVALID (-1)   /* bad argument */
#define TGM_ERR_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define TGM_ERR_NO_DEVICE (-3) /* handle was created without a device (metadata-only store) */
#define TGM_ERR_OOM (-4)

#define TGM_PADDED_NODE_ID (-1)

#define TGM_MEM_DEVICE 0 /* array arguments are device pointers, adopted as zero-copy views */
#define TGM_MEM_HOST 1   /* array arguments are host pointers, the library uploads and owns copies */

typedef struct tgm_store tgm_store;     /* device-resident time-sorted COO edge store */
typedef struct tgm_recency tgm_recency; /* per-node ring buffers (stateful sampler) */
typedef struct tgm_csr tgm_csr;         /* per-node chronological adjacency (stateless sampler) */
typedef void *tgm_stream;               /* cudaStream_t */

const char *tgm_last_error(void);
int tgm_version(void);
/* number of visible CUDA devices (0 on a CPU-only host; never an error). */
int tgm_device_count(void);

/* Tuning switches (process-wide).  "csr_feature_copy": 1 (default) = the hop-0 window sampler moves
 * feature rows with the TMA unit (cp.async.bulk through shared-memory stages), 0 = with the warp's
 * own loads/stores.  Results are identical.  "gemm_fastf32": 1 (default) = token-sized fp32 GEMMs of
 * the transformer layers run on the tensor cores (tcgen05, fp32-accurate 9xBF16 emulation), 0 = on
 * cuBLAS's SIMT SGEMM; both hold the 1e-5 parity bar.  "tc_linear": which of those GEMMs run on the
 * hand-written tcgen05 kernel of tgm_tc_linear instead: 0 = none, 1 = all, 2 (default) = the ones
 * where it measured faster than the template instantiation (the GELU-fused FFN linear); a non-zero
 * value also puts TGAT's merge-layer and folded output products on it.  "attn_folded": 1
 * (default) = tgm_attn_forward* without caller time features run the folded inference chain
 * (projection weights pre-multiplied, one warp per seed), 0 = the unfolded chain the backward
 * pass uses.  "dyg_fused_attn": 1 (default) = DyGFormer's
 * per-head QK^T / softmax / PV is one kernel with the scores in shared memory, 0 = batched cuBLAS
 * products with the scores in HBM.  "csr_tma_ctas_per_sm": cap on resident CTAs of the TMA sampler
 * (0 = automatic).  "trace": 1 = tgm_csr_build prints its phase timings on stderr. */
int tgm_set_option(const char *name, int value);

/* ------------------------------------------------------------------------------------------
 * Edge store.  Replaces DGStorageArrayBackend's edge arrays and its slice lookup:
 *   tgm/core/_storage/backends/array_backend.py:15-21   (__init__ holding DGData)
 *   tgm/core/_storage/backends/array_backend.py:301-321 (_binary_search)
 *   tgm/core/_storage/backends/array_backend.py:57-68, 259-268 (get_edges / get_edge_x)
 * The reference rebuilds O(E) boolean masks per batch; here a slice is two binary searches over the
 * timestamps and a pointer offset into the device slabs.
 *
 * src,dst: int32[E]; t: int64[E] non-decreasing; edge_x: float32[E*D] row-major or NULL (D=0).
 * mem = TGM_MEM_HOST: pointers are host memory, uploaded on `device`.
 * mem = TGM_MEM_DEVICE: pointers are device memory on `device`, adopted without a copy; t_host
 *       may pass a host copy of t, BORROWED for the life of the store (bounds are then searched on
 *       the host).  With t_host = NULL the store keeps no host mirror: the order of t is verified
 *       by one device pass and time bounds are searched on the device (16 bytes read back).
 * device = -1 (TGM_MEM_HOST only): metadata-only store; bounds work, slabs/kernels do not.
 */
int tgm_store_create(tgm_store **out, const int32_t *src, const int32_t *dst, const int64_t *t,
                     const float *edge_x, int64_t E, int32_t D, int32_t num_nodes, int device,
                     int mem, const int64_t *t_host);
void tgm_store_destroy(tgm_store *);
int tgm_store_info(const tgm_store *, int64_t *E, int32_t *D, int32_t *num_nodes, int *device);
/* [lb, ub) edge-index bounds of a slice; == _binary_search (array_backend.py:301-321):
 * lb = first index with t >= t_lo (has_lo=0: 0), ub = first index with t > t_hi (has_hi=0: E),
 * both clamped into [idx_lo, idx_hi] (idx_lo < 0: 0, idx_hi < 0: E). */
int tgm_store_bounds(const tgm_store *, int64_t t_lo, int has_lo, int64_t t_hi, int has_hi,
                     int64_t idx_lo, int64_t idx_hi, int64_t *lb, int64_t *ub);
/* zero-copy device views of edges [lb, ub); *x is NULL when the store has no features. */
int tgm_store_slab(const tgm_store *, int64_t lb, int64_t ub, const int32_t **src,
                   const int32_t **dst, const int64_t **t, const float **x);

/* ------------------------------------------------------------------------------------------
 * Stateful recency sampler: per-node ring buffers.  Replaces RecencyNeighborHook's state and
 * its two per-batch operations (tgm/hooks/neighbors/recency.py):
 *   :93-97,:410-416  state  ids int32[N,B], times int64[N,B], feats f32[N,B,D], write_pos int32[N]
 *   :111-117         reset_state            -> tgm_recency_reset
 *   :239-321         _get_recency_neighbors -> tgm_recency_query
 *   :323-399         _update                -> tgm_recency_update
 * B = max(num_nbrs) (:73).
 */
int tgm_recency_create(tgm_recency **out, int32_t num_nodes, int32_t B, int32_t D, int device);
void tgm_recency_destroy(tgm_recency *);
int tgm_recency_reset(tgm_recency *, tgm_stream stream);
/* For each seed (v, tq): among the ring of v in chronological order find the right-most entry
 * with id != -1 and time < tq, and return the k entries ending there, right-aligned and
 * left-padded with (-1, 0, 0.0f).  seeds int32[S] (negative ids wrap like torch indexing, so the
 * padded id -1 reads row N-1), tq int64[S]; out_nid int32[S*k], out_t int64[S*k],
 * out_x float32[S*k*D] (may be NULL when D == 0).  1 <= k <= B. */
int tgm_recency_query(const tgm_recency *, const int32_t *seeds, const int64_t *tq, int64_t S,
                      int32_t k, int32_t *out_nid, int64_t *out_t, float *out_x,
                      tgm_stream stream);
/* Push one batch of edges: entries (src->dst) then, unless directed, (dst->src); per node ordered
 * by (time, position in that concatenation); the last B per node are written at
 * (write_pos + j) % B and write_pos advances by the number written.  x float32[Eb*D] or NULL
 * (zeros, recency.py:325-329). */
int tgm_recency_update(tgm_recency *, const int32_t *src, const int32_t *dst, const int64_t *t,
                       const float *x, int64_t Eb, int directed, tgm_stream stream);
/* One whole hook call for the standard seed configuration -- seed_nodes_keys [edge_src, edge_dst],
 * seed_times_keys [edge_time, edge_time] (recency.py:119-171): writes the hop-0 seeds/times
 * ([src|dst], [t|t]; int32[2Eb], int64[2Eb]), queries every hop (hop h+1 seeds = flattened hop-h
 * neighbours, :141-143), THEN pushes the batch (:161-163).  num_nbrs: HOST int32[num_hops];
 * out_nid/out_t/out_x: HOST arrays [num_hops] of device pointers sized (S_h,k_h), (S_h,k_h),
 * (S_h,k_h,D) with S_0 = 2Eb, S_{h+1} = S_h*k_h (out_x entries may be NULL when D == 0). */
int tgm_recency_step(tgm_recency *, const int32_t *src, const int32_t *dst, const int64_t *t,
                     const float *x /* nullable */, int64_t Eb, int directed, int32_t num_hops,
                     const int32_t *num_nbrs, int32_t *seed_nids0, int64_t *seed_times0,
                     int32_t *const *out_nid, int64_t *const *out_t, float *const *out_x,
                     tgm_stream stream);
/* device pointers to the live state (for checkpointing / inspection); any may be NULL. */
int tgm_recency_state(const tgm_recency *, int32_t **ids, int64_t **times, float **feats,
                      int32_t **write_pos);

int tgm_recency_dims(const tgm_recency *, int32_t *num_nodes, int32_t *B, int32_t *D);

/* ------------------------------------------------------------------------------------------
 * Stateless recency sampler over a per-node chronological adjacency ("CSR").  Same answers as
 * the ring sampler driven batch by batch (recency.py:119-171), but one launch serves the seeds
 * of thousands of loader batches: the history a batch may see is encoded by an edge-index cut.
 *
 * Built for one loader geometry: batches are [e_start + i*batch_size, e_start + (i+1)*batch_size)
 * (tgm/data/loader.py:136-148,158-160).  Entries of a node are ordered (batch, time, side, edge)
 * with side 0 = node is the edge source -- the order in which the reference's stable sort
 * (recency.py:347-349) appends them to the ring.
 * colocate_x != 0 additionally stores the feature rows in adjacency order so a seed's window is
 * one contiguous read (costs 2*E*D*4 bytes of HBM when undirected).
 */
int tgm_csr_build(tgm_csr **out, const tgm_store *, int64_t e_start, int64_t batch_size,
                  int directed, int colocate_x, tgm_stream stream);
void tgm_csr_destroy(tgm_csr *);
int tgm_csr_info(const tgm_csr *, int64_t *num_entries, int64_t *e_start, int64_t *batch_size,
                 int *directed, int *colocate_x);
/* General seeds.  Seed i may only see entries whose edge index is < cut[i / cut_group]
 * (cut = first edge of the loader batch the seed belongs to; hop h>0 seeds inherit the cut of
 * their root, hence cut_group = k of the previous hops multiplied).  Among those the window is
 * the last B; the rest is tgm_recency_query's rule.  seeds may contain -1 (all-padding row). */
int tgm_csr_sample(const tgm_csr *, const int32_t *seeds, const int64_t *tq, const int64_t *cut,
                   int64_t cut_group, int64_t S, int32_t B, int32_t k, int32_t *out_nid,
                   int64_t *out_t, float *out_x, tgm_stream stream);
/* Fast path for hop-0 seeds that are the endpoints of the stream edges [e_lo, e_hi) themselves
 * (seed_nodes_keys = ['edge_src','edge_dst'], seed_times_keys = ['edge_time','edge_time']):
 * no seed arrays are read and the history cut comes from a prebuilt per-edge anchor table.
 * e_lo must sit on a batch boundary.  Rows are laid out as the concatenation of the per-batch
 * hook outputs: batch j (nb edges) owns rows [2*(lb_j - e_lo), 2*(lb_j - e_lo) + 2*nb), src seeds
 * first, then dst seeds (recency.py:181-233 with the keys above). */
int tgm_csr_sample_edges(const tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                         int32_t *out_nid, int64_t *out_t, float *out_x, tgm_stream stream);

/* Materialise, into a stateful sampler, the ring state RecencyNeighborHook would hold after the
 * stream edges [e_start, e_cut) were pushed batch by batch (equivalent for every later query; the
 * physical slot rotation follows one-by-one pushes).  Lets a windowed (pre-sampled) run hand over
 * to tgm_recency_query/_update, e.g. train stream -> validation stream with shared hook state
 * (examples/linkproppred/tgat.py:168).  e_cut must sit on a batch boundary of the adjacency. */
int tgm_csr_export_ring(const tgm_csr *, int64_t e_cut, tgm_recency *ring, tgm_stream stream);

/* Uniform full-history sampling == DGStorageArrayBackend.get_nbrs
 * (tgm/core/_storage/backends/array_backend.py:108-171, called by NeighborSamplerHook,
 * tgm/hooks/neighbors/uniform.py:122-127).  The adjacency must have been built with batch_size 1
 * and e_start 0 (entries of a node ordered by (edge, side), the reference's candidate order).
 * For every seed: its candidates are the entries with e_lo <= edge < e_hi; at most k are returned
 * LEFT-aligned, right-padded with (-1, 0, 0.0f).  With <= k candidates the output equals the
 * reference bit for bit; with more, the reference draws with CPython's random.sample and this
 * call draws a uniform k-subset from a counter-based generator keyed by (rng_seed, node id).
 * seeds int32[S] (ids outside [0, N) give padding rows). */
int tgm_csr_sample_uniform(const tgm_csr *, const int32_t *seeds, int64_t S, int64_t e_lo,
                           int64_t e_hi, int32_t k, uint64_t rng_seed, int32_t *out_nid,
                           int64_t *out_t, float *out_x, tgm_stream stream);
/* The same with the candidate range given as the slice's own closed time interval
 * [t_lo, t_hi] (has_lo / has_hi = 0: unbounded on that side) instead of edge indices: for a
 * time-sorted stream these are the same candidates (e < upper_bound(t, T) <=> t[e] <= T), and the
 * hook's per-batch cut `end_time = min(batch time) - 1` (uniform.py:125) needs no slice -> index
 * resolution (tgm_store_bounds) and no host round trip first. */
int tgm_csr_sample_uniform_time(const tgm_csr *, const int32_t *seeds, int64_t S, int64_t t_lo,
                                int has_lo, int64_t t_hi, int has_hi, int32_t k,
                                uint64_t rng_seed, int32_t *out_nid, int64_t *out_t, float *out_x,
                                tgm_stream stream);
/* Reference-exact sub-sampling.  The reference keeps random.sample(candidates, k) -- CPython's
 * global generator -- per unique seed node with more than k candidates, visiting the unique nodes
 * in ascending order (array_backend.py:118, :147-153).  That draw depends only on the candidate
 * COUNT, so the host can make it with the very same call:
 *   tgm_csr_candidate_counts: out_counts int64[S] = number of candidates of every seed;
 *   (host) picks = random.sample(range(count), k) per unique node, in ascending node order;
 *   tgm_csr_gather_picks: column c of seed s takes candidate ordinal picks[s*k + c]
 *     (int32, -1 = padding), left to right as the reference writes its sampled list (:155-169).
 * Same adjacency requirements and output layout as tgm_csr_sample_uniform. */
int tgm_csr_candidate_counts(const tgm_csr *, const int32_t *seeds, int64_t S, int64_t e_lo,
                             int64_t e_hi, int64_t *out_counts, tgm_stream stream);
int tgm_csr_gather_picks(const tgm_csr *, const int32_t *seeds, int64_t S, int64_t e_lo,
                         int64_t e_hi, int32_t k, const int32_t *picks, int32_t *out_nid,
                         int64_t *out_t, float *out_x, tgm_stream stream);

/* Host-buffer form of tgm_csr_sample_edges: what a CPU-resident caller binds (the reference
 * keeps its arrays on the CPU and moves every batch property with .to(device),
 * tgm/core/graph.py:232-263, and its CPU hook returns CPU tensors).  Stream-ordered on `stream`:
 *   1. H2D of the caller's slab of stream edges [e_lo, e_hi) (h_src/h_dst int32[n], h_t int64[n],
 *      h_x float32[n*D]; any may be NULL = already resident) into the store,
 *   2. the sampling kernel into a device staging block (`slot` in [0, TGM_HOST_SLOTS) selects
 *      it, so calls on different streams with different slots overlap),
 *   3. D2H of out_nid int32[2n*k], out_t int64[2n*k], out_x float32[2n*k*D] to host memory.
 * The host outputs are valid once `stream` is synchronised; pinned host memory keeps the copies
 * asynchronous.  The slab written in step 1 must equal what the adjacency was built from. */
#define TGM_HOST_SLOTS 4
int tgm_csr_sample_edges_host(tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                              const int32_t *h_src, const int32_t *h_dst, const int64_t *h_t,
                              const float *h_x, int32_t *h_out_nid, int64_t *h_out_t,
                              float *h_out_x, int slot, tgm_stream stream);

/* Hop-0 forms that never materialise the (S, k, D) feature block (with D = 172 a
 * sampled slot is 700 B of feature row against 12 B of id + time).  Same seeds, row layout, window
 * and time rule as tgm_csr_sample_edges; B <= 32.
 *   search = 0: history cuts from the prebuilt anchor table;
 *   search = 1: every seed is read from the store's src/dst slab and its cut found by a binary
 *               search over the node's adjacency (what the host forms below use, so the slab
 *               they upload is what the kernel consumes).
 * tgm_csr_sample_edges_ids: out_nid int32[2n*k], out_t int64[2n*k] as tgm_csr_sample_edges, plus
 *   out_eid int32[2n*k] = store edge index of every slot (-1 for padding), so that
 *   nbr_edge_x[s, c, :] == edge_x[out_eid[s, c]] (zeros where -1): a caller that owns the feature
 *   table gathers the rows itself.  Any output may be NULL.
 * tgm_csr_sample_edges_mean: fused sample + masked mean over the sampled rows
 *   (examples/linkproppred/graphmixer.py:131-135 applied to recency.py:239-321's output):
 *   out_mean float32[2n*D] = sum of the seed's valid feature rows (oldest to newest, fp32) /
 *   max(1, #valid) -- bit-identical to tgm_masked_mean over tgm_csr_sample_edges' output.  Needs
 *   colocated feature rows and D % 4 == 0.  out_nid/out_t/out_eid are optional (NULL = skip). */
int tgm_csr_sample_edges_ids(const tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                             int search, int32_t *out_nid, int64_t *out_t, int32_t *out_eid,
                             tgm_stream stream);
/* The same id form for general seeds (negatives, hop h > 0): tgm_csr_sample's arguments and rows,
 * out_eid instead of out_x.  B <= 32.  With it a multi-hop neighbourhood is sampled without ever
 * writing feature rows: tgm_attn_forward_rows reads them from the store by edge id. */
int tgm_csr_sample_ids(const tgm_csr *, const int32_t *seeds, const int64_t *tq, const int64_t *cut,
                       int64_t cut_group, int64_t S, int32_t B, int32_t k, int32_t *out_nid,
                       int64_t *out_t, int32_t *out_eid, tgm_stream stream);
int tgm_csr_sample_edges_mean(const tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                              int search, int32_t *out_nid, int64_t *out_t, int32_t *out_eid,
                              float *out_mean, tgm_stream stream);
/* Host-buffer forms of the two calls above (search = 1): H2D of the slab (h_src/h_dst int32[n],
 * h_t int64[n]; NULL = already resident), the kernel, D2H of the outputs -- 16 bytes per sampled
 * slot (ids form) or 4*D bytes per seed (+ optional ids/times) instead of 12 + 4*D bytes per slot.
 * The host gathers nbr_edge_x = edge_x[eid] from the table it already owns when it needs rows. */
int tgm_csr_sample_edges_host_ids(tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                                  const int32_t *h_src, const int32_t *h_dst, const int64_t *h_t,
                                  int32_t *h_out_nid, int64_t *h_out_t, int32_t *h_out_eid,
                                  int slot, tgm_stream stream);
int tgm_csr_sample_edges_host_mean(tgm_csr *, int64_t e_lo, int64_t e_hi, int32_t B, int32_t k,
                                   const int32_t *h_src, const int32_t *h_dst, const int64_t *h_t,
                                   int32_t *h_out_nid, int64_t *h_out_t, float *h_out_mean,
                                   int slot, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * TGN node-memory join at time-shard boundaries (multi-GPU, BASELINE config 4).  The reference
 * keeps memory f32[N, M] / last_update int64[N] in one process (tgm/nn/encoder/tgn.py:95-110,
 * written by :192-216); time-sharded over GPUs, the rows a shard touched travel through ONE
 * all-gather as packed records { int32 id | int32 0 | int64 last_update | float memory[M] }
 * (tgm_join_row_bytes(M) = 16 + 4 M bytes; M % 4 == 0).
 *   tgm_join_pack: rows[i] = record of node ids[i] for i < n; rows n .. cap-1 are padding records
 *     (id -1) so every rank sends a block of the common size `cap`.
 *   tgm_join_scatter: applies n records to memory / last_update (ids < 0 or >= num_nodes are
 *     skipped).  Applying the ranks' blocks in ascending rank order makes the later time shard win
 *     a row two shards touched. */
int64_t tgm_join_row_bytes(int32_t M);
int tgm_join_pack(const float *memory, const int64_t *last_update, int32_t M, const int32_t *ids,
                  int64_t n, int64_t cap, void *rows, tgm_stream stream);
int tgm_join_scatter(const void *rows, int64_t n, int32_t M, int32_t num_nodes, float *memory,
                     int64_t *last_update, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Negative destinations for a window of loader batches in one launch.  Replaces the per-batch
 * torch.randint(low, high, (n,), dtype=int32, device=dg.device) of RandomNegativeEdgeSamplerHook
 * (tgm/hooks/negatives/sampler.py:45-65).  out int32[total]: element g belongs to batch
 * j = g / per_batch of the window (the last batch may be short) and is element i = g % per_batch
 * of the j-th randint call:  low + philox4x32_10(key=seed, counter={(offset + 4j)/4, i}).x %
 * (high - low) -- the numbers `ceil(total / per_batch)` consecutive torch.randint calls draw from
 * a CUDA generator at (seed, offset), i.e. the reference's own device='cuda' stream.  The caller
 * advances its generator by 4 per batch.  offset % 4 == 0, per_batch <= 65536, range < 2^28 (from
 * there on ATen switches to 64-bit draws). */
int tgm_negatives_window(uint64_t seed, uint64_t offset, int64_t low, int64_t high,
                         int64_t per_batch, int64_t total, int32_t *out, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Frontier compaction (hop h+1 seeds = flatten(hop h), recency.py:141-143; the non-padded
 * subset is also what DeduplicationHook keeps, tgm/hooks/dedup.py:44-48).
 * Writes the indices i with nid[i] != -1 in increasing order to out_idx (capacity n) and their
 * number to *out_count (device int64).  One pass over the ids (4 bytes read per slot, 8 written
 * per kept slot).  Stream-ordered; the launches of a device share a small status area, so a call
 * on another stream first waits (cudaStreamWaitEvent) for the previous call's launch.  Inputs of
 * more than ~2.4e8 slots are handled by consecutive launches inside the call. */
int tgm_frontier_compact(const int32_t *nid, int64_t n, int64_t *out_idx, int64_t *out_count,
                         tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Aggregation over sampled neighbours.
 * masked mean (examples/linkproppred/graphmixer.py:131-135):
 *   out[s,:] = sum_c z[s,c,:]*[nid[s,c] != -1] / max(1, #valid), accumulated left to right in
 *   fp32.  z float32[S*k*D], nid int32[S*k], out float32[S*D]. */
int tgm_masked_mean(const float *z, const int32_t *nid, int64_t S, int32_t k, int32_t D,
                    float *out, tgm_stream stream);
/* Time2Vec (tgm/nn/modules/time_encoding.py:22-24): out[i,j] = cosf(fma(float(dt[i]), w[j], b[j]))
 * -- a single rounding of the argument, as the reference's batched nn.Linear(1,d) GEMM does --
 * and full-range cosf.
 * dt int64[n], w,b float32[d], out float32[n*d]. */
int tgm_time2vec(const int64_t *dt, int64_t n, const float *w, const float *b, int32_t d,
                 float *out, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Time-encoded attention aggregation (TGAT).  Replaces TemporalAttention.forward together with
 * the Time2Vec calls that feed it, and MergeLayer:
 *   tgm/nn/modules/attention.py:28-56   parameters: W_Q [out,out] (no bias), W_KV [2*out,key]
 *                                       (no bias), W_O [out,out] + bias, LayerNorm(out);
 *                                       out = node_dim + time_dim padded up to a multiple of
 *                                       n_heads, key = node_dim + edge_dim + time_dim
 *   tgm/nn/modules/attention.py:58-128  forward (eval mode: dropout is the identity)
 *   tgm/nn/modules/time_encoding.py:12-24  Time2Vec weight [time_dim] / bias [time_dim]
 *   tgm/nn/encoder/tgat.py:136-147      call site; tgat.py:11-38 MergeLayer
 * All parameter pointers are float32 in torch's row-major [out_features, in_features] layout, host
 * or device (copied at creation).  fp32 throughout; results are within 1e-5 of the reference.
 */
typedef struct tgm_attn tgm_attn;
typedef struct tgm_mlp2 tgm_mlp2;
int tgm_attn_create(tgm_attn **out, int32_t n_heads, int32_t node_dim, int32_t edge_dim,
                    int32_t time_dim, const float *W_Q, const float *W_KV, const float *W_O,
                    const float *b_O, const float *ln_w, const float *ln_b, float ln_eps,
                    const float *t2v_w, const float *t2v_b, int device);
/* Refresh the parameter copies in place (after an optimizer step), stream-ordered. */
int tgm_attn_set_params(tgm_attn *, const float *W_Q, const float *W_KV, const float *W_O,
                        const float *b_O, const float *ln_w, const float *ln_b,
                        const float *t2v_w, const float *t2v_b, tgm_stream stream);
void tgm_attn_destroy(tgm_attn *);
int tgm_attn_out_dim(const tgm_attn *);
/* node_x float32[S,node_dim]; nbr_node_feat float32[S,k,node_dim]; edge_feat float32[S,k,edge_dim];
 * seed_t int64[S]; nbr_t int64[S,k]; nbr_id int32[S,k] (-1 = masked slot); out float32[S,out].
 * The time features are computed inside: Time2Vec(0) for the seed, Time2Vec(seed_t - nbr_t) per
 * slot.  Masked slots take part exactly as in the reference (logit -1e10, attention.py:110-113),
 * so the caller passes the rows the reference would gather for them. */
int tgm_attn_forward(tgm_attn *, const float *node_x, const float *nbr_node_feat,
                     const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                     const int32_t *nbr_id, int64_t S, int32_t k, float *out, tgm_stream stream);
/* The plain signature of attention.py:58-66 -- time features supplied by the caller: time_feat
 * float32[S,time_dim], nbr_time_feat float32[S,k,time_dim] (argument order as in the reference). */
/* tgm_attn_forward with the sampled edge features read IN PLACE: edge_table float32[E, edge_dim]
 * is the store's feature table and edge_rows int32[S*k] the edge id of every slot (-1 = padding:
 * zeros), i.e. tgm_csr_sample_ids / tgm_csr_sample_edges_ids' out_eid.  The (S, k, edge_dim)
 * nbr_edge_x block (165 MB per 200-edge batch for hop 1 of tgat.py:136-147 at k = [20, 20],
 * D = 172) is neither written by the sampler nor re-read here.  Same result bit for bit. */
int tgm_attn_forward_rows(tgm_attn *, const float *node_x, const float *nbr_node_feat,
                          const float *edge_table, const int32_t *edge_rows, const int64_t *seed_t,
                          const int64_t *nbr_t, const int32_t *nbr_id, int64_t S, int32_t k,
                          float *out, tgm_stream stream);
/* One call for several hops of a TGAT layer (tgat.py:136-147 runs the same attention module once
 * per hop): the seeds of all hops are one row range [0, S); node_x, nbr_node_feat, seed_t, nbr_t,
 * nbr_id cover it contiguously (hop i+1's rows ARE hop i's neighbour slots, so the hop recursion
 * already stores them back to back), and the dense edge-feature blocks stay where the sampler
 * wrote them: segment i is float32[seg_rows[i], k, edge_dim], sum(seg_rows) == S, n_segs <= 4.
 * Needs tgm_attn_folded_covers(handle, k) == 1 (k <= 32, heads <= 2, node/edge_dim <= 192,
 * time_dim <= 128, out_dim <= 384, option "attn_folded" on).  Same result as the per-hop calls. */
int tgm_attn_folded_covers(const tgm_attn *, int32_t k);
int tgm_attn_forward_segments(tgm_attn *, const float *node_x, const float *nbr_node_feat,
                              const float *const *edge_feat_segs, const int64_t *seg_rows,
                              int32_t n_segs, const int64_t *seed_t, const int64_t *nbr_t,
                              const int32_t *nbr_id, int64_t S, int32_t k, float *out,
                              tgm_stream stream);
/* C[M,N] = act(A[M,K] W[N,K]^T + bias[N]) (bias nullable; act 0 = none, 2 = ReLU), fp32 FMA in
 * ascending k, for short matrices (M <= 2^20; built for the few-hundred-row products of TGAT's last
 * layer, tgat.py:136-149): 32x32 output tiles, one launch, epilogue fused. */
int tgm_small_gemm(int64_t M, int32_t N, int32_t K, const float *A, const float *W, const float *bias,
                   int32_t act, float *C, tgm_stream stream);
/* TGAT.forward (tgat.py:122-149) for inference as one call: the handle ties the L attention and
 * merge-layer handles together (borrowed: they must outlive it) and owns the scratch rows.
 * seed_ids int32[S0]; per hop i < L: nbr_ids[i] int32[S_i, k], seed_t[i] int64[S_i],
 * nbr_t[i] int64[S_i, k] with S_0 = S0, S_{i+1} = S_i k (the hop recursion of the sampler hooks:
 * hop i+1's seeds are hop i's slots), and the edge features either as dense blocks
 * edge_feat[i] float32[S_i, k, edge_dim] (edge_table NULL) or as row ids edge_rows[i] int32[S_i, k]
 * (-1 = zeros) into edge_table (edge_feat NULL).  node_x float32[num_nodes, node_dim] is indexed
 * with torch's negative-index rule (id -1 reads the last row).  out float32[S0, embed].  Every layer
 * needs tgm_attn_folded_covers(attn[j], k) == 1.  Same result as the per-hop calls (<= 1e-5). */
typedef struct tgm_tgat tgm_tgat;
int tgm_tgat_create(tgm_tgat **out, int32_t num_layers, tgm_attn *const *attn,
                    tgm_mlp2 *const *merge, int device);
void tgm_tgat_destroy(tgm_tgat *);
int tgm_tgat_forward(tgm_tgat *, const float *node_x, int64_t num_nodes, const int32_t *seed_ids,
                     int64_t S0, const int32_t *const *nbr_ids, const int64_t *const *seed_t,
                     const int64_t *const *nbr_t, const float *const *edge_feat,
                     const float *edge_table, const int32_t *const *edge_rows, int32_t k,
                     float *out, tgm_stream stream);
int tgm_attn_forward_feats(tgm_attn *, const float *node_x, const float *time_feat,
                           const float *edge_feat, const float *nbr_node_feat,
                           const float *nbr_time_feat, const int32_t *nbr_id, int64_t S, int32_t k,
                           float *out, tgm_stream stream);
/* Backward of tgm_attn_forward (what autograd computes through attention.py:58-128 and the two
 * Time2Vec calls of tgat.py:141-146).  d_out float32[S,out] is the gradient of the output.  Input
 * gradients are WRITTEN: d_node_x float32[S,node_dim], d_nbr_node_feat float32[S,k,node_dim]
 * (nullable), d_edge_feat float32[S,k,edge_dim] (nullable).  Parameter gradients are ACCUMULATED
 * (+=) into dW_Q [out,out], dW_KV [2*out,key], dW_O [out,out], db_O, dln_w, dln_b [out], dt2v_w,
 * dt2v_b [time_dim] (device pointers, torch layouts), so they can alias .grad buffers.  The
 * forward intermediates are recomputed internally; the handle must hold the same parameters the
 * forward used.  Masked slots carry no logit gradient (their logit is the constant -1e10). */
int tgm_attn_backward(tgm_attn *, const float *node_x, const float *nbr_node_feat,
                      const float *edge_feat, const int64_t *seed_t, const int64_t *nbr_t,
                      const int32_t *nbr_id, int64_t S, int32_t k, const float *d_out,
                      float *d_node_x, float *d_nbr_node_feat, float *d_edge_feat, float *dW_Q,
                      float *dW_KV, float *dW_O, float *db_O, float *dln_w, float *dln_b,
                      float *dt2v_w, float *dt2v_b, tgm_stream stream);
/* MergeLayer (tgat.py:11-38): out = fc2(relu(fc1(cat[x1, x2]))); W1 [hidden, in1+in2], W2
 * [out, hidden].  x1 float32[S,in1], x2 float32[S,in2], out float32[S,out]. */
int tgm_mlp2_create(tgm_mlp2 **out, int32_t in1, int32_t in2, int32_t hidden, int32_t out_dim,
                    const float *W1, const float *b1, const float *W2, const float *b2, int device);
void tgm_mlp2_destroy(tgm_mlp2 *);
int tgm_mlp2_forward(tgm_mlp2 *, const float *x1, const float *x2, int64_t S, float *out,
                     tgm_stream stream);
/* out[i,:] = table[ids[i]] with torch's negative indexing (id -1 reads the last row, as
 * node_x[nbr_nids] does in tgat.py:131-134).  table float32[num_rows,dim], ids int32[n]. */
int tgm_gather_rows(const float *table, int64_t num_rows, int32_t dim, const int32_t *ids,
                    int64_t n, float *out, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * TGN node memory with IdentityMessage + LastAggregator.  Replaces TGNMemory's state machine
 * (tgm/nn/encoder/tgn.py):
 *   :128-133 state (memory f32[N,M], last_update int64[N], per-node message stores)
 *   :149-152 reset_state -> tgm_tgn_reset        :157-163 forward      -> tgm_tgn_forward
 *   :165-178 update_state -> tgm_tgn_update_state :245-251 train(False) -> tgm_tgn_flush
 * memory_updater = GRUCell(raw_msg_dim + 2*memory_dim + time_dim, memory_dim): gru_w_ih
 * [3M, in], gru_w_hh [3M, M], gru_b_ih [3M], gru_b_hh [3M] in torch's (r, z, n) gate order;
 * time_enc = Time2Vec(time_dim).  Parameter pointers may be host or device (copied).
 */
typedef struct tgm_tgn tgm_tgn;
int tgm_tgn_create(tgm_tgn **out, int32_t num_nodes, int32_t raw_msg_dim, int32_t memory_dim,
                   int32_t time_dim, const float *gru_w_ih, const float *gru_w_hh,
                   const float *gru_b_ih, const float *gru_b_hh, const float *t2v_w,
                   const float *t2v_b, int device);
void tgm_tgn_destroy(tgm_tgn *);
int tgm_tgn_reset(tgm_tgn *, tgm_stream stream);
/* device pointers to the live memory [N,M] / last_update [N] (checkpointing, all-gather). */
int tgm_tgn_state(const tgm_tgn *, float **memory, int64_t **last_update);
/* n_id int64[n] -> out_memory f32[n,M], out_last_update int64[n].  training != 0: the memory the
 * nodes would have after consuming their stored messages, WITHOUT writing it (tgn.py:158-159);
 * training == 0: the stored rows (tgn.py:160-161). */
int tgm_tgn_forward(tgm_tgn *, const int64_t *n_id, int64_t n, int training, float *out_memory,
                    int64_t *out_last_update, tgm_stream stream);
/* One batch of events: src,dst int32[Eb], t int64[Eb], raw_msg f32[Eb,raw_msg_dim].
 * training != 0: update the memory of the batch's nodes from their stored messages, then store
 * this batch's messages; training == 0: store first, then update (tgn.py:170-177). */
int tgm_tgn_update_state(tgm_tgn *, const int32_t *src, const int32_t *dst, const int64_t *t,
                         const float *raw_msg, int64_t Eb, int training, tgm_stream stream);
/* train() -> eval() transition: consume every stored message into memory, clear the stores. */
int tgm_tgn_flush(tgm_tgn *, tgm_stream stream);
/* Training (autograd of memory(n_id) in examples/linkproppred/tgn.py:100-118, where
 * loss.backward() runs AFTER memory.update_state()).  The state has moved on by then, so the
 * training forward hands the caller what the backward needs of its rows:
 *   saved_x f32[n,in] (in = raw_msg_dim + 2*memory_dim + time_dim: the LastAggregator's message per
 *   node, zeros without one), saved_h f32[n,M] (memory[n_id]), saved_aux f32[n,
 *   tgm_tgn_saved_aux_width()] = {float32(t - last_update) of that message, 1 if the node has a
 *   message else 0} for the LastAggregator.
 * tgm_tgn_forward_saved == tgm_tgn_forward(training = 1) + those rows.  tgm_tgn_backward is a
 * function of the saved rows, the handle's CURRENT parameters and d_memory f32[n,M] only; it ADDS
 * into the gradient buffers (torch layouts: g_w_ih [3M,in], g_w_hh [3M,M], g_b_ih/g_b_hh [3M],
 * g_t2v_w/g_t2v_b [time_dim]; zero them first for plain gradients).  memory[...] and the raw
 * messages are buffers/inputs without gradient in the reference (tgn.py:128-133, :154-155).
 * tgm_tgn_set_params refreshes the parameter copies in place after an optimizer step (the node
 * state and the message stores are kept); pointers may be host or device. */
/* Message aggregator (tgn.py:43-63).  TGM_TGN_AGGR_LAST (the default, LastAggregator): the message
 * with the largest t per node.  TGM_TGN_AGGR_MEAN (MeanAggregator = scatter(mean)): the mean over
 * all messages of the node's last batch as source and as destination; the batches pushed since the
 * last reset/flush are then kept in an append-only device log (log_capacity events reserved up
 * front, grown by doubling -- the one synchronising path).  Switching resets the state; call it
 * right after tgm_tgn_create.  tgm_tgn_saved_aux_width: floats per row of saved_aux below
 * (2 for LAST, 2 * time_dim for MEAN: the means of sin(arg) and sin(arg) * dt over the messages). */
#define TGM_TGN_AGGR_LAST 0
#define TGM_TGN_AGGR_MEAN 1
int tgm_tgn_set_aggregator(tgm_tgn *, int kind, int64_t log_capacity, tgm_stream stream);
int tgm_tgn_saved_aux_width(const tgm_tgn *);
int tgm_tgn_set_params(tgm_tgn *, const float *gru_w_ih, const float *gru_w_hh,
                       const float *gru_b_ih, const float *gru_b_hh, const float *t2v_w,
                       const float *t2v_b, tgm_stream stream);
int tgm_tgn_forward_saved(tgm_tgn *, const int64_t *n_id, int64_t n, float *out_memory,
                          int64_t *out_last_update, float *saved_x, float *saved_h,
                          float *saved_aux, tgm_stream stream);
int tgm_tgn_backward(tgm_tgn *, const float *saved_x, const float *saved_h, const float *saved_aux,
                     int64_t n, const float *d_memory, float *g_w_ih, float *g_w_hh, float *g_b_ih,
                     float *g_b_hh, float *g_t2v_w, float *g_t2v_b, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * DyGFormer forward (eval mode).  Replaces DyGFormer.forward with its co-occurrence encoder,
 * patching and transformer layers (tgm/nn/encoder/dygformer.py:13-77, 80-143, 243-444).
 * Parameters use torch's layouts ([out_features, in_features] row-major; MultiheadAttention
 * in_proj_weight [3E, E]); E = 4 * channel_dim; pointers may be host or device (copied).
 */
typedef struct {
  const float *in_proj_w, *in_proj_b;   /* transformers.i.multi_head_attention.in_proj_{weight,bias} */
  const float *out_proj_w, *out_proj_b; /* ....multi_head_attention.out_proj.{weight,bias} */
  const float *ffn1_w, *ffn1_b;         /* ....linear_layers.0  [4E, E] */
  const float *ffn2_w, *ffn2_b;         /* ....linear_layers.1  [E, 4E] */
  const float *ln0_w, *ln0_b, *ln1_w, *ln1_b; /* ....norm_layers.{0,1} */
} tgm_dyg_layer;
typedef struct {
  int32_t node_dim, edge_dim, time_dim, channel_dim, out_dim, patch_size, num_layers, num_heads;
  int32_t seq_len; /* max_input_sequence_length = 1 + sampled neighbours per node */
  float ln_eps;
  const float *t2v_w, *t2v_b;                     /* time_encoder.w.{weight,bias} */
  const float *cooc_w1, *cooc_b1, *cooc_w2, *cooc_b2; /* co-occurrence MLP: Linear(1,C), Linear(C,C) */
  const float *proj_w[4], *proj_b[4];             /* projection_layer.{node,edge,time,neighbor_co_occurrence} */
  const tgm_dyg_layer *layers;                    /* [num_layers] */
  const float *out_w, *out_b;                     /* output_layer */
} tgm_dyg_params;
typedef struct tgm_dyg tgm_dyg;
int tgm_dyg_create(tgm_dyg **out, const tgm_dyg_params *params, int device);
void tgm_dyg_destroy(tgm_dyg *);
/* node_x f32[num_nodes,node_dim]; src,dst int32[B] (edge_index rows); edge_time int64[B];
 * nbrs int32[2B,k], nbr_t int64[2B,k], nbr_x f32[2B,k,edge_dim] with k = seq_len - 1: rows [0,B)
 * are the neighbours of the sources, [B,2B) of the destinations (dygformer.py:262-270);
 * out_src,out_dst f32[B,out_dim]. */
int tgm_dyg_forward(tgm_dyg *, const float *node_x, int64_t num_nodes, const int32_t *src,
                    const int32_t *dst, const int64_t *edge_time, const int32_t *nbrs,
                    const int64_t *nbr_t, const float *nbr_x, int64_t B, float *out_src,
                    float *out_dst, tgm_stream stream);

/* Training support for DyGFormer (dropout 0).  tgm_dyg_set_params refreshes the handle's parameter
 * copies in place after an optimizer step (same shapes as at creation).  tgm_dyg_backward
 * recomputes the forward pass, keeping the activations the chain rule needs, and OVERWRITES every
 * gradient buffer of `grads` (torch layouts, as in tgm_dyg_params) with the gradient of
 * sum(out_src * d_src) + sum(out_dst * d_dst) -- what loss.backward() leaves in .grad of the
 * reference module (tgm/nn/encoder/dygformer.py:243-431 under autograd).  Inputs as tgm_dyg_forward;
 * d_src, d_dst f32[B,out_dim].  Input features receive no gradient. */
typedef struct {
  float *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b, *ffn1_w, *ffn1_b, *ffn2_w, *ffn2_b;
  float *ln0_w, *ln0_b, *ln1_w, *ln1_b;
} tgm_dyg_layer_grads;
typedef struct {
  float *t2v_w, *t2v_b, *cooc_w1, *cooc_b1, *cooc_w2, *cooc_b2;
  float *proj_w[4], *proj_b[4];
  tgm_dyg_layer_grads *layers; /* [num_layers] */
  float *out_w, *out_b;
} tgm_dyg_grads;
int tgm_dyg_set_params(tgm_dyg *, const tgm_dyg_params *params, tgm_stream stream);
int tgm_dyg_backward(tgm_dyg *, const float *node_x, int64_t num_nodes, const int32_t *src,
                     const int32_t *dst, const int64_t *edge_time, const int32_t *nbrs,
                     const int64_t *nbr_t, const float *nbr_x, int64_t B, const float *d_src,
                     const float *d_dst, const tgm_dyg_grads *grads, tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Dense token-by-weight linear on the tensor cores, hand-written for sm_100a (tcgen05.mma
 * kind::tf32, accumulator in tensor memory; every fp32 operand is split in flight into two TF32
 * terms and three products are accumulated, so the result holds the 1e-5 parity bar):
 *   out[S,N] = act(A[S,K] W[N,K]^T + bias[N] (+ residual[S,N])),  act = exact GELU when gelu != 0.
 * This is what DyGFormer's in/out projections and FFN linears run on (reference:
 * tgm/nn/encoder/dygformer.py:80-143, torch fp32 GEMMs).  residual may alias out; gelu and
 * residual are exclusive; N % 4 == 0, K % 4 == 0, arrays 16-byte aligned; all device pointers. */
int tgm_tc_linear(int64_t S, int32_t N, int32_t K, const float *A, const float *W,
                  const float *bias, const float *residual, int gelu, float *out,
                  tgm_stream stream);
/* The same contract on the CUTLASS sm_100 FastF32 collective (9xBF16 emulation, TMA, 2-SM tcgen05
 * tiles of 256x128x16 -- K steps of 32 / 64 and 1-SM 128x128 tiles measured 6-20 % slower on the
 * DyGFormer shapes) that DyGFormer's linears use by default.  TGM_ERR_INVALID when the shape is
 * unsupported or the library was built without the CUTLASS header tree. */
int tgm_fastf32_linear(int64_t S, int32_t N, int32_t K, const float *A, const float *W,
                       const float *bias, const float *residual, int gelu, float *out,
                       tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * TGN embedding.  Replaces GraphAttentionEmbedding (tgm/nn/encoder/tgn.py:14-40) as called from
 * examples/linkproppred/tgn.py:74-98: rel_t = last_update[edge_src] - t, edge_attr =
 * [Time2Vec(rel_t) | msg], then torch_geometric's TransformerConv(in_channels, out_channels/heads,
 * heads, edge_dim = msg_dim + time_dim) -- third-party arithmetic (torch-geometric 2.6.1, not
 * vendored, reference tests shape-only): restated from the published algorithm, PARITY UNPINNED.
 * Eval-mode forward (attention dropout = identity).  Weights in torch layout [out, in]; out_channels
 * is the full output width (heads * per-head channels); W_edge [out_channels, time_dim + msg_dim]
 * with the time-encoding columns first (tgn.py:39); pointers may be host or device (copied).
 */
typedef struct tgm_gae tgm_gae;
int tgm_gae_create(tgm_gae **out, int32_t in_channels, int32_t out_channels, int32_t heads,
                   int32_t msg_dim, int32_t time_dim, const float *W_query, const float *b_query,
                   const float *W_key, const float *b_key, const float *W_value,
                   const float *b_value, const float *W_edge, const float *W_skip,
                   const float *b_skip, const float *t2v_w, const float *t2v_b, int device);
void tgm_gae_destroy(tgm_gae *);
/* x f32[n,in_channels], last_update int64[n] (TGNMemory.forward's outputs for the batch's unique
 * nodes); edge_src,edge_dst int64[m] = edge_index rows (LOCAL indices in [0,n): messages flow
 * edge_src -> edge_dst and are normalised per edge_dst), t int64[m], msg f32[m,msg_dim];
 * out f32[n,out_channels].  Nodes without incoming edges get the skip projection only. */
int tgm_gae_forward(tgm_gae *, const float *x, const int64_t *last_update, int64_t n,
                    const int64_t *edge_src, const int64_t *edge_dst, const int64_t *t,
                    const float *msg, int64_t m, float *out, tgm_stream stream);
/* Training (dropout 0: the convolution's attention dropout cannot follow the reference's RNG).
 * tgm_gae_backward recomputes the forward from the same inputs and ADDS the gradients of
 * sum(out * d_out), d_out f32[n,out_channels], into: d_x f32[n,in_channels] (nullable),
 * g_W_qkvs f32[4*out_channels,in_channels] and g_b_qkvs f32[4*out_channels] = the query, key, value
 * and skip linears stacked in that order, g_W_edge f32[out_channels,time_dim+msg_dim],
 * g_t2v_w/g_t2v_b f32[time_dim] (Time2Vec, through the first time_dim columns of edge_attr).
 * Source-node rows of d key / d value are accumulated with atomics (summation order varies).
 * tgm_gae_set_params refreshes the parameter copies in place after an optimizer step. */
int tgm_gae_set_params(tgm_gae *, const float *W_query, const float *b_query, const float *W_key,
                       const float *b_key, const float *W_value, const float *b_value,
                       const float *W_edge, const float *W_skip, const float *b_skip,
                       const float *t2v_w, const float *t2v_b, tgm_stream stream);
int tgm_gae_backward(tgm_gae *, const float *x, const int64_t *last_update, int64_t n,
                     const int64_t *edge_src, const int64_t *edge_dst, const int64_t *t,
                     const float *msg, int64_t m, const float *d_out, float *d_x, float *g_W_qkvs,
                     float *g_b_qkvs, float *g_W_edge, float *g_t2v_w, float *g_t2v_b,
                     tgm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Batch de-duplication.  Replaces DeduplicationHook.__call__ (tgm/hooks/dedup.py:35-67): masked
 * gather per hop + torch.cat + torch.unique(sorted=True) + a torch.searchsorted per lookup.
 * Node ids are dense in [0, num_nodes): the set is a bitmap, unique ids come out ascending
 * without a sort and global_to_local(v) = #{u in set: u < v} (== searchsorted left, for members
 * and non-members alike) is an O(1) rank lookup.
 * The caller owns the state of one batch: bitmap uint32[bitmap_words], prefix int32[prefix_len]
 * and tmp (tmp_bytes) from tgm_dedup_sizes; they must stay alive while tgm_dedup_map is used.
 */
int tgm_dedup_sizes(int32_t num_nodes, int64_t *bitmap_words, int64_t *prefix_len,
                    int64_t *tmp_bytes);
/* parts/sizes/skip_padded: HOST arrays of n_parts (<= 8) entries: device pointers to int32 ids,
 * their lengths, and whether -1 entries are padding to drop (nbr_nids, dedup.py:46-48) or
 * ordinary elements (seed arrays, as torch.unique would keep them).  out_unique int32[capacity >=
 * min(sum sizes, num_nodes) + 1]; out_count int64[1] on the device: the number of unique ids, or
 * -1 when an id fell outside [-1, num_nodes). */
int tgm_dedup_unique(const int32_t *const *parts, const int64_t *sizes, const int32_t *skip_padded,
                     int32_t n_parts, int32_t num_nodes, uint32_t *bitmap, int32_t *prefix,
                     void *tmp, int64_t tmp_bytes, int32_t *out_unique, int64_t *out_count,
                     tgm_stream stream);
/* out_local[i] = searchsorted(unique, ids[i]) as int32 (dedup.py:57-59), ids int32[n]. */
int tgm_dedup_map(const uint32_t *bitmap, const int32_t *prefix, const void *tmp,
                  int32_t num_nodes, const int32_t *ids, int64_t n, int32_t *out_local,
                  tgm_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TGM_B200_H */
