// Hand-written sm_100a tensor-core linear for the dense contractions of the aggregation path:
//     out[S,N] = act(A[S,K] W[N,K]^T + bias[N] (+ residual[S,N])),   act = exact GELU or identity
// (DyGFormer's token-by-weight linears, reference tgm-team/tgm @ 5183dc9:
//  tgm/nn/encoder/dygformer.py:80-143 -- nn.MultiheadAttention in/out projections and the two
//  FFN linears, 97 % of the model's flops -- evaluated there by torch's fp32 GEMMs.)
//
// The 1e-5 parity bar rules out plain TF32/BF16 inputs, so every fp32 operand is split in flight
// into two TF32 terms, x = hi + lo with hi = x truncated to 10 mantissa bits (what the tensor core
// reads of an fp32 word) and lo = x - hi (exact in fp32), and three tcgen05 products are
// accumulated in fp32 in tensor memory:  A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T
// (the dropped lo x lo term is 2^-20 relative; measured max abs error vs float64 in
// tests/test_gpu_tc_linear.py).
//
// Structure (no library code: plain PTX for tcgen05 / mbarrier / fences):
//   * one CTA (512 threads, one per SM) = one 128 x BN output tile at a time, persistent over tiles
//     (n fastest, so the CTAs of one wave share A rows in L2)
//   * per K chunk of 40: every thread loads A and W rows from global with 128-bit loads, splits
//     them and stores hi / lo into shared memory in the canonical no-swizzle K-major UMMA layout
//     (8-row x 16-byte core matrices; LBO = 128 B between the core matrices of one k-step, SBO =
//     1280 B between 8-row groups); two operand buffers, so chunk c + 1 is staged while the
//     tensor core works on chunk c
//   * ONE thread issues the chunk's 15 `tcgen05.mma.cta_group::1.kind::tf32` instructions
//     (5 k-steps of 8 x 3 products) and commits them to an mbarrier.  The hi x hi products go to
//     one TMEM accumulator, the two correction products to a second one, and BOTH are drained
//     into fp32 registers after every chunk (tcgen05.ld, then round-to-nearest adds): the tensor
//     core's own accumulate step truncates, and a 75-step chain in a single TMEM accumulator
//     drifts by ~1e-5 (measured); 5-step chains promoted in registers stay at fp32-GEMM accuracy
//   * epilogue straight from the register accumulators: bias / residual / exact GELU, 208-byte
//     row segments to global
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
constexpr size_t kBufBytes = size_t(2) * BM * KC * 4 + size_t(2) * kMaxBN * KC * 4;  // one stage

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

// Staging of one K chunk of both operands as hi / lo: BM rows of A and UN rows of W, KC floats
// each (rows / columns outside the matrices read as zeros).  Item idx of a chunk = one 16-byte K
// column group of one row; 8 consecutive items are 8 consecutive rows of one group (a
// conflict-free 128-byte core matrix in shared memory), the next 8 the neighbouring group (so a
// warp reads whole 32-byte sectors of 8 rows).  A thread owns the same (at most 7) items in every
// chunk, so their global / shared offsets are computed ONCE per kernel (`StageMap`); per chunk
// only the K offset moves.  All of a thread's loads are issued before the first split / store:
// one L2 latency per chunk, not one per item.
constexpr int kC4 = KC / 4;
constexpr int kAItems = BM * kC4;                                   // 1280
constexpr int kMaxItems = (kAItems + kMaxBN * kC4 + kThreads - 1) / kThreads;  // 7

struct StageMap {
  uint32_t g_off[kMaxItems];  // element offset from the tile's first row: r * K + 4 * c4
  uint32_t s_off[kMaxItems];  // byte offset of the hi copy inside a stage (A_hi | A_lo | W_hi | W_lo)
  uint32_t r_pack[2];         // row inside the tile, 8 bits per item (255 = no item)
  uint32_t c_pack[2];         // 4 * c4, 8 bits per item
  uint32_t w_mask;            // bit i: item i belongs to W
};

__device__ __forceinline__ void stage_map_init(StageMap &m, int UN, int K, int tid) {
  const int total = kAItems + UN * kC4;
  m.r_pack[0] = m.r_pack[1] = m.c_pack[0] = m.c_pack[1] = m.w_mask = 0u;
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    const int idx = tid + i * kThreads;
    const bool is_w = idx >= kAItems;
    const int it = is_w ? idx - kAItems : idx;
    const int r = idx < total ? (it & 7) + 8 * (it / (8 * kC4)) : 255;
    const int c4 = (it >> 3) % kC4;
    m.g_off[i] = uint32_t(r) * uint32_t(K) + 4u * c4;
    m.s_off[i] = (is_w ? 2u * BM * KC * 4u : 0u) + umma_off(r == 255 ? 0 : r, c4);
    m.r_pack[i >> 2] |= uint32_t(r) << (8 * (i & 3));
    m.c_pack[i >> 2] |= uint32_t(4 * c4) << (8 * (i & 3));
    m.w_mask |= uint32_t(is_w) << i;
  }
}

__device__ __forceinline__ void stage_chunk(const StageMap &m, const float *__restrict__ At,
                                            const float *__restrict__ Wt, int a_rows, int w_rows,
                                            int K, int k0, unsigned char *base) {
  float4 v[kMaxItems];
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    const int r = int((m.r_pack[i >> 2] >> (8 * (i & 3))) & 255u);
    const int c = int((m.c_pack[i >> 2] >> (8 * (i & 3))) & 255u);
    const bool is_w = (m.w_mask >> i) & 1u;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < (is_w ? w_rows : a_rows) && k0 + c < K)
      v[i] = __ldg(reinterpret_cast<const float4 *>((is_w ? Wt : At) + m.g_off[i] + k0));
  }
#pragma unroll
  for (int i = 0; i < kMaxItems; ++i) {
    if (((m.r_pack[i >> 2] >> (8 * (i & 3))) & 255u) != 255u) {
      const bool is_w = (m.w_mask >> i) & 1u;
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u);
      h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u);
      h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u);
      h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u);
      l.x = v[i].x - h.x, l.y = v[i].y - h.y, l.z = v[i].z - h.z, l.w = v[i].w - h.w;
      unsigned char *hi = base + m.s_off[i];
      *reinterpret_cast<float4 *>(hi) = h;
      *reinterpret_cast<float4 *>(hi + (is_w ? kMaxBN : BM) * KC * 4) = l;
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
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int UN = (BN + 15) & ~15;  // UMMA N
  const uint32_t bar = smem_addr(&s_bar);

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
  auto stage = [&](int64_t m0, int n0, int kc) {
    const int a_rows = int(S - m0 < BM ? S - m0 : BM);
    const int w_rows = (n0 + BN < N ? n0 + BN : N) - n0;
    stage_chunk(map, A + m0 * K, W + int64_t(n0) * K, a_rows, w_rows, K, kc * KC,
                smem + size_t(kc & 1) * kBufBytes);
  };

  const int64_t m_tiles = (S + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_chunks = (K + KC - 1) / KC;
  // this thread's share of the accumulator: TMEM lane (= tile row) 32 * (warp & 3) + lane,
  // columns [cbase, cbase + UN / 4)
  const int q = warp & 3, cols = UN >> 2, cbase = (warp >> 2) * cols;
  const uint32_t lane_addr = tmem + (uint32_t(32 * q) << 16);
  uint32_t phase = 0;  // parity of the next completion of the MMA barrier
  for (int64_t tile = blockIdx.x; tile < m_tiles * n_tiles; tile += gridDim.x) {
    const int64_t m0 = (tile / n_tiles) * BM;
    const int n0 = int(tile % n_tiles) * BN;
    float acc[kColsPerThread];
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) acc[j] = 0.f;
    // acc += this thread's columns of the accumulator at TMEM column `col0`: 16-column loads, one
    // wait per 32 columns
    // (columns past this thread's share may be read -- they stay inside the 256-column region of
    // the accumulator -- but are never used)
    auto drain = [&](uint32_t col0) {
      uint32_t t[32];
#pragma unroll
      for (int j = 0; j < kColsPerThread; j += 32) {
        if (j < cols) {
          const uint32_t ta = lane_addr + col0 + uint32_t(cbase + j);
          tmem_ld16(ta, t);
          if (j + 16 < cols) tmem_ld16(ta + 16, t + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (j + u < kColsPerThread && j + u < cols) acc[j + u] += __uint_as_float(t[u]);
        }
      }
    };
    stage(m0, n0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy
    __syncthreads();
    for (int kc = 0; kc < k_chunks; ++kc) {
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ah = smem_addr(smem + size_t(kc & 1) * kBufBytes);
        const uint32_t al = ah + BM * KC * 4, wh = al + BM * KC * 4, wl = wh + kMaxBN * KC * 4;
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
      // the other buffer was last read by the MMAs of chunk kc - 1, which completed before the
      // drain of the previous iteration: stage the next chunk while the tensor core runs
      if (kc + 1 < k_chunks) stage(m0, n0, kc + 1);
      mbar_wait(bar, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // promote the chunk's hi x hi partial sum into the fp32 register accumulators (the
      // correction accumulator is 2^-11 of it: its own truncation is far below fp32 resolution, so
      // it runs over the whole tile and is added once, below)
      drain(0u);
      // TMEM drained and the next chunk's operands written: both visible before the next MMAs
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    drain(uint32_t(kCorrCol));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // next tile's MMAs overwrite it
    // ---- epilogue: bias / residual / GELU on the register accumulators -> global ---------------
    const int64_t row = m0 + 32 * q + lane;
    if (row < S) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; j += 4) {
        const int c = cbase + j, n = n0 + c;
        if (j < cols && c < BN && n < N) {  // N and BN are multiples of 4
          const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + n));
          float4 y = make_float4(acc[j] + b.x, acc[j + 1] + b.y, acc[j + 2] + b.z, acc[j + 3] + b.w);
          if (residual) {
            const float4 r = *reinterpret_cast<const float4 *>(residual + row * N + n);
            y.x += r.x, y.y += r.y, y.z += r.z, y.w += r.w;
          }
          if (gelu) {  // exact GELU (F.gelu)
            y.x = 0.5f * y.x * (1.f + erff(y.x * 0.70710678118654752440f));
            y.y = 0.5f * y.y * (1.f + erff(y.y * 0.70710678118654752440f));
            y.z = 0.5f * y.z * (1.f + erff(y.z * 0.70710678118654752440f));
            y.w = 0.5f * y.w * (1.f + erff(y.w * 0.70710678118654752440f));
          }
          *reinterpret_cast<float4 *>(out + row * N + n) = y;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(uint32_t(kTmemCols))
                 : "memory");
}

constexpr size_t kSmemBytes = 2 * kBufBytes;

}  // namespace

namespace tgm {

// tgm_set_option("tc_linear", 0|1).  Default 0: measured on the B200 (profiles/README.md) this
// kernel reaches 45 TFLOP/s fp32-equivalent on the DyGFormer shapes (1.5x cuBLAS fp32) against 64
// for the CUTLASS FastF32 collective, which therefore stays the default for the token linears.
int g_tc_linear = 0;

// 1 = computed, 0 = shape / alignment not supported (caller falls back), < 0 = error
int tc3_linear(int64_t S, int N, int K, const float *A, const float *W, const float *bias,
               const float *residual, int gelu, float *out, cudaStream_t stream) {
  if (S < 1 || N < 4 || K < 4 || N % 4 || K % 4 || !bias || !aligned16(A) || !aligned16(W) ||
      !aligned16(out) || !aligned16(bias) || (residual && !aligned16(residual)) ||
      (gelu && residual))
    return 0;
  // column tile: as wide as fits the 208-column accumulator while dividing N evenly
  const int n_tiles = (N + 199) / 200;
  int BN = ((N + n_tiles - 1) / n_tiles + 3) & ~3;
  if (BN > 200) BN = 200;
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
