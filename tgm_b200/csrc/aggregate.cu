// Aggregation over sampled neighbours: masked mean and Time2Vec.
// Reference (tgm-team/tgm @ 5183dc9): examples/linkproppred/graphmixer.py:131-135 (masked mean
// over the k sampled neighbours), tgm/nn/modules/time_encoding.py:22-24 (Time2Vec).
// Both are HBM-bound elementwise/segmented work: one pass, 128-bit accesses, no tensor cores.
#include "common.cuh"

using namespace tgm;

namespace {

// out[s,:] = sum_c z[s,c,:] * [nid[s,c] != -1] / max(1, #valid); the k terms are accumulated left
// to right in fp32, multiplying by the 0/1 mask exactly as the reference does.
template <bool VEC4>
__global__ void __launch_bounds__(256)
masked_mean_kernel(const float *__restrict__ z, const int32_t *__restrict__ nid, int64_t S, int k,
                   int D, float *__restrict__ out) {
  if (VEC4) {
    const int D4 = D >> 2;
    const int64_t total = S * D4;
    const float4 *z4 = reinterpret_cast<const float4 *>(z);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t s = i / D4;
      const int d = int(i - s * D4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
#pragma unroll 4  // four independent 128-bit loads in flight per thread; the adds stay in order
      for (int c = 0; c < k; ++c) {
        const float m = nid[s * k + c] != TGM_PADDED_NODE_ID ? 1.f : 0.f;
        cnt += m != 0.f;
        const float4 v = ldg_stream_f4(z4 + (s * k + c) * D4 + d);
        acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, m));
        acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, m));
        acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, m));
        acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, m));
      }
      const float den = float(cnt > 1 ? cnt : 1);
      o4[i] = make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
    }
  } else {
    const int64_t total = S * D;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t s = i / D;
      const int d = int(i - s * D);
      float acc = 0.f;
      int cnt = 0;
      for (int c = 0; c < k; ++c) {
        const float m = nid[s * k + c] != TGM_PADDED_NODE_ID ? 1.f : 0.f;
        cnt += m != 0.f;
        acc = __fadd_rn(acc, __fmul_rn(__ldg(z + (s * k + c) * D + d), m));
      }
      out[i] = acc / float(cnt > 1 ? cnt : 1);
    }
  }
}

// out[r,j] = cos(fma(float(dt[r]), w[j], b[j])): int64 -> fp32 cast (time_encoding.py:23), then
// nn.Linear(1,d), whose batched CPU GEMM (and cuBLAS on the reference's CUDA path) fuses the
// multiply-add into ONE rounding -- verified against torch CPU: 100% of arguments equal the fused
// form, 80% the two-rounding form; with arguments up to 2.7e6 (ulp 0.25) the choice decides the
// result.  One warp per row, lanes along the d columns (coalesced 4*d-byte row stores, no integer
// divisions); the weights of the first four column tiles stay in registers; cosine = t2v_cos
// (common.cuh: 1.6e-7 abs error, 15 instructions).
__global__ void __launch_bounds__(256)
time2vec_kernel(const int64_t *__restrict__ dt, int64_t n, const float *__restrict__ w,
                const float *__restrict__ b, int d, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  float wr[4], br[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = lane + 32 * t;
    wr[t] = c < d ? __ldg(w + c) : 0.f;
    br[t] = c < d ? __ldg(b + c) : 0.f;
  }
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const float x = float(dt[r]);
    float *o = out + r * d;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = lane + 32 * t;
      if (c < d) o[c] = t2v_cos(__fmaf_rn(x, wr[t], br[t]));
    }
    for (int c = 128 + lane; c < d; c += 32) o[c] = t2v_cos(__fmaf_rn(x, __ldg(w + c), __ldg(b + c)));
  }
}

}  // namespace

extern "C" int tgm_masked_mean(const float *z, const int32_t *nid, int64_t S, int32_t k,
                               int32_t D, float *out, tgm_stream stream) {
  TGM_REQUIRE(S >= 0 && k >= 1 && D >= 0, "tgm_masked_mean: bad sizes");
  if (S == 0 || D == 0) return TGM_OK;
  TGM_REQUIRE(z && nid && out, "tgm_masked_mean: NULL array argument");
  const bool vec4 = (D % 4 == 0) && aligned16(z) && aligned16(out);
  cudaStream_t st = as_stream(stream);
  if (vec4)
    masked_mean_kernel<true><<<grid_for(S * (D / 4), 256, 8), 256, 0, st>>>(z, nid, S, k, D, out);
  else
    masked_mean_kernel<false><<<grid_for(S * D, 256, 8), 256, 0, st>>>(z, nid, S, k, D, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_time2vec(const int64_t *dt, int64_t n, const float *w, const float *b,
                            int32_t d, float *out, tgm_stream stream) {
  TGM_REQUIRE(n >= 0 && d >= 1, "tgm_time2vec: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(dt && w && b && out, "tgm_time2vec: NULL array argument");
  time2vec_kernel<<<grid_for(n, 8, 8), 256, 0, as_stream(stream)>>>(dt, n, w, b, d, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
