// Hand-written sm_100a tensor-core linear for the dense contractions of the aggregation path:
//     out[S,N] = act(A[S,K] W[N,K]^T + bias[N] (+ residual[S,N])),   act = exact GELU or identity
// (DyGFormer's token-by-weight linears, reference tgm-team/tgm @ 5183dc9:
//  tgm/nn/encoder/dygformer.py:80-143 -- nn.MultiheadAttention in/out projections and the two
//  FFN linears, 97 % of the model's flops -- evaluated there by torch's fp32 GEMMs.)
//
// The 1e-5 parity bar rules out plain TF32/BF16 inputs, so every fp32 operand is used as two TF32
// terms, x = hi + lo: hi = x truncated to 10 mantissa bits -- exactly what the tensor core reads of
// an fp32 word (checked bit for bit on the B200: feeding raw words == feeding masked words), so the
// RAW tile is the hi operand -- and lo = x - hi (exact in fp32).  Three tcgen05 products are
// accumulated in fp32 in tensor memory:  A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T
// (the dropped lo x lo term is 2^-20 relative; measured max abs error vs float64 in
// tests/test_gpu_tc_linear.py).
//
// Structure (no library code: plain PTX for tcgen05 / mbarrier / cp.async / fences):
//   * one CTA (512 threads, one per SM) = one 128 x BN output tile at a time, persistent over tiles
//     (n fastest, so the CTAs of one wave share A rows in L2)
//   * K runs in chunks of 40.  The raw fp32 rows of A and W go global -> shared by 16-byte
//     `cp.async`, each granule straight to its place in the canonical no-swizzle K-major UMMA
//     layout (8-row x 16-byte core matrices; LBO = 128 B between the core matrices of one k-step,
//     SBO = 1280 B between 8-row groups) -- no registers, two chunks ahead of the tensor core
//   * a shared -> shared pass computes lo = x - trunc(x) for the NEXT chunk while the tensor core
//     works on the current one (two raw + two lo stages, 210 KB)
//   * ONE thread issues the chunk's 15 `tcgen05.mma.cta_group::1.kind::tf32` instructions
//     (5 k-steps of 8 x 3 products) and commits them to an mbarrier.  The hi x hi products go to
//     one TMEM accumulator, the two correction products to a second one; the first is drained
//     into fp32 registers after every chunk (tcgen05.ld, then round-to-nearest adds): the tensor
//     core's own accumulate step truncates, and a 75-step chain in a single TMEM accumulator
//     drifts by ~1e-5 (measured); 5-step chains promoted in registers stay at fp32-GEMM accuracy.
//     The correction accumulator (2^-11 of the other) runs over the whole tile
//   * epilogue: bias / residual / exact GELU on the register accumulators, transposed through
//     shared memory so that every warp writes whole 800-byte output rows
#include <algorithm>

#include "common.cuh"

using namespace tgm;

namespace {

constexpr int kThreads = 512;
constexpr int BM = 128;           // UMMA M
constexpr int KC = 40;            // K chunk: 5 k-steps of 8 (tf32 UMMA K)
constexpr int kMaxBN = 208;       // UMMA N (multiple of 16, <= 256): 200 columns + padding
constexpr int kTmemCols = 512;    // two accumulators: columns [0, 208) and [256, 464)
constexpr int kCorrCol = 256;
constexpr int kColsPerThread = kMaxBN / 4;     // 4 warps share a TMEM lane quarter: 52 columns each
constexpr uint32_t kLBO = 128;                 // bytes between the two core matrices of a k-step
constexpr uint32_t kSBO = (KC / 4) * 128;      // bytes between 8-row groups
constexpr uint32_t kABytes = BM * KC * 4, kWBytes = kMaxBN * KC * 4;
constexpr uint32_t kStageBytes = kABytes + kWBytes;  // one stage: A rows | W rows
constexpr int kEpiStride = kMaxBN + 1;         // odd row stride of the epilogue transpose: no conflicts

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return uint32_t(__cvta_generic_to_shared(p));
}

// no-swizzle K-major shared-memory matrix descriptor (tcgen05 "version 1")
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3fffu);        // start address, 16-byte units
  d |= uint64_t((kLBO >> 4) & 0x3fffu) << 16;   // leading (K) byte offset
  d |= uint64_t((kSBO >> 4) & 0x3fffu) << 32;   // stride (M/N) byte offset
  d |= uint64_t(1) << 46;                       // descriptor version: Blackwell
  return d;                                     // base offset 0, layout type 0 = no swizzle
}

// kind::tf32, fp32 accumulate, A and B K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                   // c_format = F32
  d |= 2u << 7;                   // a_format = TF32
  d |= 2u << 10;                  // b_format = TF32
  d |= uint32_t(N >> 3) << 17;    // n_dim
  d |= uint32_t(M >> 4) << 24;    // m_dim
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// element (row r, k column kk) of an operand chunk, in bytes from the chunk's base
__device__ __forceinline__ uint32_t umma_off(int r, int c4) {
  return uint32_t(r >> 3) * kSBO + uint32_t(c4) * kLBO + uint32_t(r & 7) * 16u;
}

// Staging.  Item idx of a chunk = one 16-byte K column group of one row; 8 consecutive items are 8
// consecutive rows of one group (a 128-byte core matrix in shared memory), the next 8 the
// neighbouring group (so a warp reads whole 32-byte sectors of 8 rows).  A thread owns the same
// (at most 7) items in every chunk, so their global / shared offsets are computed ONCE per kernel
// (`StageMap`); per chunk only the K offset moves.  Rows / columns outside the matrices are
// zero-filled by cp.async itself (src-size 0).
constexpr int kC4 = KC / 4;
constexpr int kAItems = BM * kC4;                                   // 1280
constexpr int kMaxItems = (kAItems + kMaxBN * kC4 + kThreads - 1) / kThreads;  // 7

struct StageMap {
  uint32_t g_off[kMaxItems];  // element offset from the tile's first row: r * K + 4 * c4
  uint32_t s_off[kMaxItems];  // byte offset inside a stage (A rows | W rows)
  uint32_t r_pack[2];         // row inside the tile, 8 bits per item (255 = no item)
  uint32_t c_pack[2];         // 4 * c4, 8 bits per item
  uint32_t w_mask;            // bit i: item i belongs to W
  uint32_t have;              // bit i: item i exists
};

__device__ __forceinline__ void stage_map_init(StageMap &m, int UN, int K, int tid) {
  const int total = kAItems + UN * kC4;
  m.r_pack[0] = m.r_pack[1] = m.c_pack[0] = m.c_pack[1] = m.w_mask = m.have = 0u;
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    const int idx = tid + i * kThreads;
    const bool is_w = idx >= kAItems;
    const int it = is_w ? idx - kAItems : idx;
    const int r = idx < total ? (it & 7) + 8 * (it / (8 * kC4)) : 255;
    const int c4 = (it >> 3) % kC4;
    m.g_off[i] = uint32_t(r) * uint32_t(K) + 4u * c4;
    m.s_off[i] = (is_w ? kABytes : 0u) + umma_off(r == 255 ? 0 : r, c4);
    m.r_pack[i >> 2] |= uint32_t(r) << (8 * (i & 3));
    m.c_pack[i >> 2] |= uint32_t(4 * c4) << (8 * (i & 3));
    m.w_mask |= uint32_t(is_w) << i;
    m.have |= uint32_t(idx < total) << i;
  }
}

// which of this thread's items lie inside the matrices for a tile with a_rows x w_rows valid rows
__device__ __forceinline__ uint32_t stage_row_mask(const StageMap &m, int a_rows, int w_rows) {
  uint32_t ok = 0u;
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    const int r = int((m.r_pack[i >> 2] >> (8 * (i & 3))) & 255u);
    ok |= uint32_t(r < (((m.w_mask >> i) & 1u) ? w_rows : a_rows)) << i;
  }
  return ok & m.have;
}

// raw fp32 rows of one chunk -> shared memory, asynchronously (one commit group per chunk).
// `rows_ok` = stage_row_mask of the tile; `full_k`: the chunk lies entirely inside K.
__device__ __forceinline__ void stage_raw_async(const StageMap &m, const float *__restrict__ At,
                                                const float *__restrict__ Wt, uint32_t rows_ok,
                                                int K, int k0, uint32_t raw_s) {
  const bool full_k = k0 + KC <= K;
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    if ((m.have >> i) & 1u) {
      bool ok = (rows_ok >> i) & 1u;
      if (!full_k) ok = ok && k0 + int((m.c_pack[i >> 2] >> (8 * (i & 3))) & 255u) < K;
      const float *src = ok ? (((m.w_mask >> i) & 1u) ? Wt : At) + m.g_off[i] + k0 : At;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(raw_s + m.s_off[i]),
                   "l"(src), "r"(ok ? 16u : 0u)
                   : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// lo = x - trunc_tf32(x) of this thread's own granules (it copied them itself: its own
// cp.async.wait_group makes them visible to it, no barrier needed in between)
__device__ __forceinline__ void stage_lo(const StageMap &m, const unsigned char *raw,
                                         unsigned char *lo) {
#pragma unroll
  for (int i0 = 0; i0 < kMaxItems; i0 += 2) {  // two granules in flight: short dependent chains
    float4 v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (i0 + u < kMaxItems && ((m.have >> (i0 + u)) & 1u))
        v[u] = *reinterpret_cast<const float4 *>(raw + m.s_off[i0 + u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i0 + u < kMaxItems && ((m.have >> (i0 + u)) & 1u)) {
        float4 l;
        l.x = v[u].x - __uint_as_float(__float_as_uint(v[u].x) & 0xffffe000u);
        l.y = v[u].y - __uint_as_float(__float_as_uint(v[u].y) & 0xffffe000u);
        l.z = v[u].z - __uint_as_float(__float_as_uint(v[u].z) & 0xffffe000u);
        l.w = v[u].w - __uint_as_float(__float_as_uint(v[u].w) & 0xffffe000u);
        *reinterpret_cast<float4 *>(lo + m.s_off[i0 + u]) = l;
      }
    }
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__global__ void __launch_bounds__(kThreads, 1)
tc3_linear_kernel(const float *__restrict__ A, const float *__restrict__ W,
                  const float *__restrict__ bias, const float *residual, float *out, int64_t S,
                  int N, int K, int BN, int gelu) {
  extern __shared__ __align__(1024) unsigned char smem[];  // raw0 | raw1 | lo0 | lo1
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int UN = (BN + 15) & ~15;  // UMMA N
  const uint32_t bar = smem_addr(&s_bar);
  unsigned char *lo_base = smem + 2 * kStageBytes;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_addr(&s_tmem)),
                 "r"(uint32_t(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = umma_idesc(BM, UN);

  StageMap map;
  stage_map_init(map, UN, K, tid);
  const int64_t m_tiles = (S + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t tiles = m_tiles * n_tiles;
  const int k_chunks = (K + KC - 1) / KC;
  // raw chunk kc of tile `t` -> raw stage (kc & 1), asynchronously
  auto prefetch = [&](int64_t t, int kc) {
    const int64_t m0 = (t / n_tiles) * BM;
    const int n0 = int(t % n_tiles) * BN;
    const int a_rows = int(S - m0 < BM ? S - m0 : BM);
    const int w_rows = (n0 + BN < N ? n0 + BN : N) - n0;
    stage_raw_async(map, A + m0 * K, W + int64_t(n0) * K, stage_row_mask(map, a_rows, w_rows), K,
                    kc * KC, smem_addr(smem) + uint32_t(kc & 1) * kStageBytes);
  };
  // this thread's share of the accumulator: TMEM lane (= tile row) 32 * (warp & 3) + lane,
  // columns [cbase, cbase + UN / 4)
  const int q = warp & 3, cols = UN >> 2, cbase = (warp >> 2) * cols;
  const uint32_t lane_addr = tmem + (uint32_t(32 * q) << 16);
  uint32_t phase = 0;  // parity of the next completion of the MMA barrier

  int64_t tile = blockIdx.x;
  if (tile < tiles) {  // the first tile's first two chunks
    prefetch(tile, 0);
    if (k_chunks > 1) prefetch(tile, 1);
  }
  for (; tile < tiles; tile += gridDim.x) {
    const int64_t m0 = (tile / n_tiles) * BM;
    const int n0 = int(tile % n_tiles) * BN;
    float acc[kColsPerThread];
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) acc[j] = 0.f;
    // acc += this thread's columns of the accumulator at TMEM column `col0`, 16 columns per load
    // (columns past this thread's share may be read -- they stay inside the
    // 256-column region of the accumulator -- but are never used)
    auto drain = [&](uint32_t col0) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; j += 16) {
        if (j < cols) {
          uint32_t t[16];
          tmem_ld16(lane_addr + col0 + uint32_t(cbase + j), t);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int u = 0; u < 16; ++u)
            if (j + u < kColsPerThread && j + u < cols) acc[j + u] += __uint_as_float(t[u]);
        }
      }
    };
    // chunk 0: its raw tile has landed (requested during the previous tile) -> its lo terms
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    stage_lo(map, smem, lo_base);
    for (int kc = 0; kc < k_chunks; ++kc) {
      // raw(kc) and lo(kc) written by all threads -> visible to the tensor core
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ah = smem_addr(smem) + uint32_t(kc & 1) * kStageBytes, wh = ah + kABytes;
        const uint32_t al = smem_addr(lo_base) + uint32_t(kc & 1) * kStageBytes, wl = al + kABytes;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint32_t o = uint32_t(ks) * 2u * kLBO;  // two core matrices per k-step
          umma_tf32(tmem, umma_desc(ah + o), umma_desc(wh + o), idesc, ks != 0);
          umma_tf32(tmem + kCorrCol, umma_desc(ah + o), umma_desc(wl + o), idesc, (kc | ks) != 0);
          umma_tf32(tmem + kCorrCol, umma_desc(al + o), umma_desc(wh + o), idesc, 1u);
        }
        // arrives on the barrier when every MMA issued so far has completed (implies
        // tcgen05.fence::before_thread_sync)
        asm volatile(
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
            : "memory");
      }
      // while the tensor core runs: the lo terms of the next chunk (its raw tile was requested
      // two iterations ago; lo stage (kc+1)&1 was last read by the MMAs of chunk kc-1, complete)
      if (kc + 1 < k_chunks) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        stage_lo(map, smem + size_t((kc + 1) & 1) * kStageBytes,
                 lo_base + size_t((kc + 1) & 1) * kStageBytes);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // promote the chunk's hi x hi partial sum into the fp32 register accumulators (the
      // correction accumulator is 2^-11 of it: its own truncation is far below fp32 resolution, so
      // it runs over the whole tile and is added once, below)
      drain(0u);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      // raw stage kc&1 is free again (its MMAs completed): request chunk kc+2 of this tile, or
      // the first chunks of the next tile
      if (kc + 2 < k_chunks) {
        prefetch(tile, kc + 2);
      } else if (tile + gridDim.x < tiles && (kc & 1) < k_chunks) {
        // tail of the tile: the released stage takes the next tile's chunk of the same parity
        // (chunk 0 lives in stage 0, chunk 1 in stage 1)
        prefetch(tile + gridDim.x, kc & 1);
      }
    }
    drain(uint32_t(kCorrCol));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // next tile's MMAs overwrite it
    // ---- epilogue: bias / residual / GELU in registers, transposed through the lo stages (free:
    // every MMA of the tile is complete) so that warps write whole output rows --------------------
    float *ep = reinterpret_cast<float *>(lo_base);
    {
      const int row = 32 * q + lane;
#pragma unroll
      for (int j = 0; j < kColsPerThread; ++j)
        if (j < cols) ep[row * kEpiStride + cbase + j] = acc[j];
    }
    __syncthreads();
    {
      // lane owns columns lane, lane + 32, ... of every row its warp writes: their bias values
      // stay in registers, and the loads of a row are all issued before its first store
      const int n_valid = (n0 + BN < N ? n0 + BN : N) - n0;
      constexpr int kCols = (kMaxBN + 31) / 32;  // 7
      float bv[kCols];
#pragma unroll
      for (int i = 0; i < kCols; ++i) {
        const int c = lane + 32 * i;
        bv[i] = c < n_valid ? __ldg(bias + n0 + c) : 0.f;
      }
      const int r_end = int(S - m0 < BM ? S - m0 : BM);
      for (int r = warp; r < r_end; r += kThreads / 32) {
        const int64_t o = (m0 + r) * N + n0;
        float y[kCols];
#pragma unroll
        for (int i = 0; i < kCols; ++i) {
          const int c = lane + 32 * i;
          y[i] = c < n_valid ? ep[r * kEpiStride + c] + bv[i] : 0.f;
        }
        if (residual) {
#pragma unroll
          for (int i = 0; i < kCols; ++i) {
            const int c = lane + 32 * i;
            if (c < n_valid) y[i] += residual[o + c];
          }
        }
#pragma unroll
        for (int i = 0; i < kCols; ++i) {
          const int c = lane + 32 * i;
          if (c < n_valid) {
            float v = y[i];
            if (gelu == 1) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // exact GELU
            else if (gelu == 2) v = fmaxf(v, 0.f);                                       // ReLU
            out[o + c] = v;
          }
        }
      }
    }
    __syncthreads();  // the lo stages are reused by the next tile
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(uint32_t(kTmemCols))
                 : "memory");
}

constexpr size_t kSmemBytes = 4 * size_t(kStageBytes);

}  // namespace

namespace tgm {

// tgm_set_option("tc_linear", 0|1|2): 0 = never; 1 (default) = every token linear of DyGFormer
// runs here -- one DyGFormer forward at the default 12800 tokens: 0.772 ms against 0.770 ms with
// the CUTLASS FastF32 collective for all of them (scratch/dyg_ab.py, same box, interleaved);
// 2 = only the activation-fused linear here and the other three on the collective, the fastest
// mix measured (0.743 ms: in isolation 82 / 80 / 29 / 60 us here vs 93 / 80 / 30 / 51 us for
// FFN1 / FFN2 / out_proj / in_proj, profiles/r2_tc_linear_timings.txt).  TGAT's products choose
// per shape in attn_fold.cu (any non-zero value enables this kernel there).
int g_tc_linear = 1;
int g_tc_bn = 0;  // tgm_set_option("tc_bn", v): probe switch, v > 0 forces the column tile width

// 1 = computed, 0 = shape / alignment not supported (caller falls back), < 0 = error
int tc3_linear(int64_t S, int N, int K, const float *A, const float *W, const float *bias,
               const float *residual, int gelu, float *out, cudaStream_t stream) {
  if (S < 1 || N < 4 || K < 4 || N % 4 || K % 4 || !bias || !aligned16(A) || !aligned16(W) ||
      !aligned16(out) || !aligned16(bias) || (residual && !aligned16(residual)) ||
      (gelu && residual))
    return 0;
  // column tile: as wide as fits the 208-column accumulator while dividing N evenly; a short
  // matrix (fewer row tiles than half the SMs) is cut into narrower column tiles -- down to 48
  // columns -- so that more SMs share it and each tile's chunk loop gets shorter
  int n_tiles = (N + 199) / 200;
  {
    const int64_t m_tiles = (S + BM - 1) / BM;
    if (m_tiles * n_tiles * 2 <= kSmCount) {
      const int fit = int(kSmCount / m_tiles), narrow = (N + 47) / 48;
      n_tiles = std::max(n_tiles, std::min(fit, narrow));
    }
  }
  int BN = ((N + n_tiles - 1) / n_tiles + 3) & ~3;
  if (BN > 200) BN = 200;
  {
    // one more column tile when that saves time over whole waves of CTAs: a tile costs about
    // (128 + BN) column units (fitted on the B200: 23.3 us at BN = 200, 19.9 us at BN = 152,
    // K = 200), a launch ceil(tiles / 148) of them -- e.g. 12800 x 600: 300 tiles of 200 columns
    // are 3 waves (70 us), 400 tiles of 152 columns also 3 (60 us)
    const int64_t m_tiles = (S + BM - 1) / BM;
    auto cost = [&](int nt, int bn) {
      const int64_t waves = (m_tiles * nt + kSmCount - 1) / kSmCount;
      return waves * (128 + bn);
    };
    const int nt2 = n_tiles + 1;
    const int bn2 = ((N + nt2 - 1) / nt2 + 7) & ~7;
    if (m_tiles * n_tiles > kSmCount && bn2 >= 48 && cost(nt2, bn2) < cost(n_tiles, BN)) BN = bn2;
  }
  if (g_tc_bn > 0) BN = std::min(200, (g_tc_bn + 3) & ~3);
  static bool configured = false;
  if (!configured) {
    TGM_CUDA(cudaFuncSetAttribute(tc3_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(kSmemBytes)));
    configured = true;
  }
  const int64_t tiles = ((S + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = int(tiles < kSmCount ? tiles : kSmCount);
  tc3_linear_kernel<<<grid, kThreads, kSmemBytes, stream>>>(A, W, bias, residual, out, S, N, K, BN,
                                                            gelu);
  TGM_LAUNCH_CHECK();
  return 1;
}

}  // namespace tgm

extern "C" int tgm_tc_linear(int64_t S, int32_t N, int32_t K, const float *A, const float *W,
                             const float *bias, const float *residual, int gelu, float *out,
                             tgm_stream stream) {
  TGM_REQUIRE(A && W && bias && out, "tgm_tc_linear: NULL array argument");
  const int rc = tgm::tc3_linear(S, N, K, A, W, bias, residual, gelu, out, as_stream(stream));
  if (rc == 0)
    return fail(TGM_ERR_INVALID, "tgm_tc_linear: needs S >= 1, N % 4 == 0, K % 4 == 0, 16-byte "
                                 "aligned arrays and not both gelu and residual");
  return rc < 0 ? rc : TGM_OK;
}

// The CUTLASS FastF32 (9xBF16) instantiation behind DyGFormer's token linears, same contract
// (gemm_fastf32.cu).
extern "C" int tgm_fastf32_linear(int64_t S, int32_t N, int32_t K, const float *A, const float *W,
                                  const float *bias, const float *residual, int gelu, float *out,
                                  tgm_stream stream) {
  TGM_REQUIRE(A && W && bias && out, "tgm_fastf32_linear: NULL array argument");
  const int rc = tgm::fastf32_linear(S, N, K, A, W, bias, residual, gelu, out, as_stream(stream));
  if (rc == 0)
    return fail(TGM_ERR_INVALID, "tgm_fastf32_linear: shape / alignment not supported, or the "
                                 "library was built without the CUTLASS headers");
  return rc < 0 ? rc : TGM_OK;
}
