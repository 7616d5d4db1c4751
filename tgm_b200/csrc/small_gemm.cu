// Short-matrix fp32 products with the bias / activation in the epilogue:
//     C[M, N] = act(A[M, K] W[N, K]^T + bias[N]),   all row-major, fp32 FMA in ascending k
// for the few-hundred-row products of the aggregation path (TGAT's last layer runs on the 600 seeds
// of a batch: tgm/nn/encoder/tgat.py:136-149 -> 600x888x172, 600x272x888, 600x172x444, 600x172x172).
// A library SGEMM serves them with split-K + a reduction kernel + a separate bias pass (three
// launches, 18-31 us), the 128-row tensor-core tile leaves most SMs idle; here one launch of 32x32
// output tiles (171 CTAs for 600x272) walks K in 32-wide chunks through double-buffered shared
// memory, 8 outputs per thread, and writes act(acc + bias) directly.
#include "common.cuh"

using namespace tgm;

namespace {

constexpr int kBM = 32, kBN = 32, kBK = 32, kThreadsSG = 128, kPad = 4;

template <bool VEC>
__device__ __forceinline__ void load_tile(const float *__restrict__ src, int ld, int rows_valid,
                                          int k0, int K, int tid, float (&reg)[8]) {
  // tile rows r = tid / 8 + 16 i, columns (tid % 8) * 4 .. + 3
  const int kc = (tid & 7) * 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (tid >> 3) + 16 * i;
    const float *p = src + int64_t(r) * ld + k0 + kc;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_valid && k0 + kc < K) v = __ldg(reinterpret_cast<const float4 *>(p));
      reg[4 * i] = v.x, reg[4 * i + 1] = v.y, reg[4 * i + 2] = v.z, reg[4 * i + 3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reg[4 * i + j] = (r < rows_valid && k0 + kc + j < K) ? __ldg(p + j) : 0.f;
    }
  }
}

__device__ __forceinline__ void store_tile(float (*dst)[kBM + kPad], int tid, const float (&reg)[8]) {
  const int kc = (tid & 7) * 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (tid >> 3) + 16 * i;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[kc + j][r] = reg[4 * i + j];
  }
}

// act: 0 none, 2 ReLU.  grid (N tiles, M tiles, sum of the groups' batch counts); a group is a
// batch of equally shaped products (same W, A / C advancing by a stride); all groups share M, N, ldc.
__global__ void __launch_bounds__(kThreadsSG)
sgemm_nt_small_kernel(const SmallGemmGroups g, int ldc, int M, int N, int act) {
  __shared__ __align__(16) float As[2][kBK][kBM + kPad];
  __shared__ __align__(16) float Ws[2][kBK][kBN + kPad];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;  // 4 columns x 2 rows per thread
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  int gi = 0, z = int(blockIdx.z);
#pragma unroll
  for (int i = 0; i < kSmallGemmMaxGroups - 1; ++i)
    if (i + 1 < g.n && z >= g.batch[gi]) z -= g.batch[gi], gi = i + 1;
  const int lda = g.lda[gi], ldw = g.ldw[gi], K = g.K[gi];
  const float *__restrict__ A = g.A[gi] + z * g.strideA[gi] + int64_t(m0) * lda;
  const float *__restrict__ W = g.W[gi] + int64_t(n0) * ldw;
  const float *__restrict__ bias = g.bias[gi];
  float *__restrict__ C = g.C[gi] + z * g.strideC[gi];
  const int a_rows = min(kBM, M - m0), w_rows = min(kBN, N - n0);
  const bool vec = g.vec[gi];  // 128-bit loads: K, the leading dimensions and the bases allow it
  float ra[8], rw[8];
  auto load = [&](int k0) {
    if (vec) {
      load_tile<true>(A, lda, a_rows, k0, K, tid, ra);
      load_tile<true>(W, ldw, w_rows, k0, K, tid, rw);
    } else {
      load_tile<false>(A, lda, a_rows, k0, K, tid, ra);
      load_tile<false>(W, ldw, w_rows, k0, K, tid, rw);
    }
  };
  load(0);
  store_tile(As[0], tid, ra);
  store_tile(Ws[0], tid, rw);
  __syncthreads();
  float acc[2][4] = {};
  const int chunks = (K + kBK - 1) / kBK;
  for (int c = 0; c < chunks; ++c) {
    const int cur = c & 1;
    if (c + 1 < chunks) load((c + 1) * kBK);  // the next chunk's loads fly during this chunk's FMAs
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float2 a = *reinterpret_cast<const float2 *>(&As[cur][k][2 * ty]);
      const float4 w = *reinterpret_cast<const float4 *>(&Ws[cur][k][4 * tx]);
      acc[0][0] = fmaf(a.x, w.x, acc[0][0]), acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, w.z, acc[0][2]), acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, w.x, acc[1][0]), acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, w.z, acc[1][2]), acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
    }
    if (c + 1 < chunks) {
      store_tile(As[cur ^ 1], tid, ra);
      store_tile(Ws[cur ^ 1], tid, rw);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + 2 * ty + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tx + j;
      if (n < N) {
        float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
        if (act == 2) v = fmaxf(v, 0.f);
        C[int64_t(m) * ldc + n] = v;
      }
    }
  }
}

}  // namespace

namespace tgm {

// 1 = computed, 0 = not applicable (batch or tile count out of range), < 0 = error
int small_gemm_groups(const SmallGemmGroups &g_in, int64_t M, int N, int ldc, int act,
                      cudaStream_t st) {
  SmallGemmGroups g = g_in;
  if (M < 1 || N < 1 || g.n < 1 || g.n > kSmallGemmMaxGroups || M > (int64_t(1) << 20)) return 0;
  int64_t batches = 0;
  for (int i = 0; i < g.n; ++i) {
    if (g.K[i] < 1 || g.batch[i] < 1) return 0;
    batches += g.batch[i];
    g.vec[i] = g.K[i] % 4 == 0 && g.lda[i] % 4 == 0 && g.ldw[i] % 4 == 0 && g.strideA[i] % 4 == 0 &&
               aligned16(g.A[i]) && aligned16(g.W[i]);
  }
  const dim3 grid(unsigned((N + kBN - 1) / kBN), unsigned((M + kBM - 1) / kBM), unsigned(batches));
  if (grid.y > 65535 || batches > 65535) return 0;
  sgemm_nt_small_kernel<<<grid, kThreadsSG, 0, st>>>(g, ldc, int(M), N, act);
  TGM_LAUNCH_CHECK();
  return 1;
}

int small_gemm_nt(int64_t M, int N, int K, const float *A, int lda, int64_t strideA, const float *W,
                  int ldw, int64_t strideW, const float *bias, int act, float *C, int ldc,
                  int64_t strideC, int batch, cudaStream_t st) {
  if (strideW != 0 && batch != 1) return 0;  // one W per group
  SmallGemmGroups g{};
  g.n = 1;
  g.A[0] = A, g.W[0] = W, g.bias[0] = bias, g.C[0] = C;
  g.K[0] = K, g.lda[0] = lda, g.ldw[0] = ldw, g.strideA[0] = strideA, g.strideC[0] = strideC;
  g.batch[0] = batch;
  return small_gemm_groups(g, M, N, ldc, act, st);
}

}  // namespace tgm

extern "C" int tgm_small_gemm(int64_t M, int32_t N, int32_t K, const float *A, const float *W,
                              const float *bias, int32_t act, float *C, tgm_stream stream) {
  TGM_REQUIRE(M >= 0 && N >= 1 && K >= 1, "tgm_small_gemm: bad sizes");
  TGM_REQUIRE(act == 0 || act == 2, "tgm_small_gemm: act must be 0 (none) or 2 (ReLU)");
  if (M == 0) return TGM_OK;
  TGM_REQUIRE(A && W && C, "tgm_small_gemm: NULL array argument");
  const int rc = tgm::small_gemm_nt(M, N, K, A, K, 0, W, K, 0, bias, act, C, N, 0, 1, as_stream(stream));
  if (rc == 0) return fail(TGM_ERR_INVALID, "tgm_small_gemm: M must be <= 2^20");
  return rc < 0 ? rc : TGM_OK;
}
