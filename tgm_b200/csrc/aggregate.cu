// Aggregation over sampled neighbours: masked mean and Time2Vec.
// Reference (tgm-team/tgm @ 5183dc9): examples/linkproppred/graphmixer.py:131-135 (masked mean
// over the k sampled neighbours), tgm/nn/modules/time_encoding.py:22-24 (Time2Vec).
// Both are HBM-bound elementwise/segmented work: one pass, 128-bit accesses, no tensor cores.
#include "common.cuh"

using namespace tgm;

namespace {

// out[s,:] = sum_c z[s,c,:] * [nid[s,c] != -1] / max(1, #valid); the k terms are accumulated left
// to right in fp32, multiplying by the 0/1 mask exactly as the reference does.
// DQ: D / 4 as a compile-time constant (the row stride of the batch's loads becomes an immediate:
// one address register instead of one per load), 0 = run-time width
template <bool VEC4, int DQ>
__global__ void __launch_bounds__(256, DQ ? 3 : 2)
masked_mean_kernel(const float *__restrict__ z, const int32_t *__restrict__ nid, int64_t S, int k,
                   int D, float *__restrict__ out) {
  if (VEC4) {
    const int D4 = DQ ? DQ : D >> 2;
    const int64_t total = S * D4;
    const float4 *z4 = reinterpret_cast<const float4 *>(z);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t s = i / D4;
      const int d = int(i - s * D4);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      // rows go in batches of 10: their 128-bit loads (and mask loads) are all issued before the
      // first add -- 160 bytes in flight per thread (ncu on the 4-deep version: 92 % of the stall
      // cycles were waits on these loads at 65 % occupancy); the adds stay in row order
      constexpr int kBatch = 10;
      const int32_t *nrow = nid + s * k;
      const float4 *zrow = z4 + s * int64_t(k) * D4 + d;
      for (int c0 = 0; c0 < k; c0 += kBatch) {
        float4 v[kBatch];
        float m[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const bool in = c0 + u < k;
          v[u] = in ? ldg_stream_f4(zrow + int64_t(c0 + u) * D4) : make_float4(0.f, 0.f, 0.f, 0.f);
          m[u] = (in && __ldg(nrow + c0 + u) != TGM_PADDED_NODE_ID) ? 1.f : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (c0 + u < k) {
            cnt += m[u] != 0.f;
            acc.x = __fadd_rn(acc.x, __fmul_rn(v[u].x, m[u]));
            acc.y = __fadd_rn(acc.y, __fmul_rn(v[u].y, m[u]));
            acc.z = __fadd_rn(acc.z, __fmul_rn(v[u].z, m[u]));
            acc.w = __fadd_rn(acc.w, __fmul_rn(v[u].w, m[u]));
          }
        }
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


// The same values with every lane busy and one 128-bit store per lane (d % 4 == 0): the output
// is walked as a flat array of float4 -- element e = (row e / (d/4), columns 4 (e % (d/4)) ..+3)
// -- so a warp writes 512 contiguous bytes whatever d is (d = 100: the row kernel above leaves
// 28 of 32 lanes idle in its fourth column tile and issues four 128-byte stores per 400-byte
// row).  (row, column group) advance by a constant stride without a division; w and b sit in
// shared memory as float4.  The reduced-range cosine is used when |x| * max|w| + max|b| < 2^22,
// checked per element with one compare; the rare rest goes through t2v_cos.
constexpr int kT2vFlatMaxD = 2048;
__device__ __noinline__ float4 t2v_cos4_general(float a0, float a1, float a2, float a3) {
  return make_float4(t2v_cos(a0), t2v_cos(a1), t2v_cos(a2), t2v_cos(a3));
}
__global__ void __launch_bounds__(256)
time2vec_flat4_kernel(const int64_t *__restrict__ dt, int64_t n, const float *__restrict__ w,
                      const float *__restrict__ b, int d4, float4 *__restrict__ out) {
  __shared__ float4 s_w[kT2vFlatMaxD / 4], s_b[kT2vFlatMaxD / 4];
  __shared__ float s_red[2][8];
  float wm = 0.f, bm = 0.f;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    const float4 wv = __ldg(reinterpret_cast<const float4 *>(w) + c);
    const float4 bv = __ldg(reinterpret_cast<const float4 *>(b) + c);
    s_w[c] = wv, s_b[c] = bv;
    wm = fmaxf(wm, fmaxf(fmaxf(fabsf(wv.x), fabsf(wv.y)), fmaxf(fabsf(wv.z), fabsf(wv.w))));
    bm = fmaxf(bm, fmaxf(fmaxf(fabsf(bv.x), fabsf(bv.y)), fmaxf(fabsf(bv.z), fabsf(bv.w))));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
  }
  if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = wm, s_red[1][threadIdx.x >> 5] = bm;
  __syncthreads();
  wm = bm = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) wm = fmaxf(wm, s_red[0][i]), bm = fmaxf(bm, s_red[1][i]);
  // |x| <= xlim  =>  |fma(x, w, b)| < 2^22 for every column (a NaN / zero wm leaves xlim non-finite
  // or negative: the comparison then sends everything to the general path or the fast one, both exact)
  const float xlim = wm > 0.f ? (4194304.f - bm) / wm * 0.999f : 3.0e38f;

  const int64_t total = n * d4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t srow = stride / d4;
  const int scol = int(stride - srow * d4);
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  int64_t row = e / d4;
  int c = int(e - row * d4);
  for (; e < total; e += stride) {
    const float x = float(__ldg(dt + row));
    const float4 wv = s_w[c], bv = s_b[c];
    const float a0 = __fmaf_rn(x, wv.x, bv.x), a1 = __fmaf_rn(x, wv.y, bv.y),
                a2 = __fmaf_rn(x, wv.z, bv.z), a3 = __fmaf_rn(x, wv.w, bv.w);
    float4 y;
    if (fabsf(x) <= xlim)
      y = make_float4(t2v_cos_fast(a0), t2v_cos_fast(a1), t2v_cos_fast(a2), t2v_cos_fast(a3));
    else
      y = t2v_cos4_general(a0, a1, a2, a3);
    out[e] = y;
    row += srow;
    c += scol;
    if (c >= d4) c -= d4, ++row;
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
    switch (D) {
      case 16: masked_mean_kernel<true, 4><<<grid_for(S * 4, 256, 3), 256, 0, st>>>(z, nid, S, k, D, out); break;
      case 100: masked_mean_kernel<true, 25><<<grid_for(S * 25, 256, 3), 256, 0, st>>>(z, nid, S, k, D, out); break;
      case 172: masked_mean_kernel<true, 43><<<grid_for(S * 43, 256, 3), 256, 0, st>>>(z, nid, S, k, D, out); break;
      default: masked_mean_kernel<true, 0><<<grid_for(S * (D / 4), 256, 2), 256, 0, st>>>(z, nid, S, k, D, out);
    }
  else
    masked_mean_kernel<false, 0><<<grid_for(S * D, 256, 2), 256, 0, st>>>(z, nid, S, k, D, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}

extern "C" int tgm_time2vec(const int64_t *dt, int64_t n, const float *w, const float *b,
                            int32_t d, float *out, tgm_stream stream) {
  TGM_REQUIRE(n >= 0 && d >= 1, "tgm_time2vec: bad sizes");
  if (n == 0) return TGM_OK;
  TGM_REQUIRE(dt && w && b && out, "tgm_time2vec: NULL array argument");
  if (d % 4 == 0 && d <= kT2vFlatMaxD && aligned16(w) && aligned16(b) && aligned16(out))
    time2vec_flat4_kernel<<<grid_for(n * (d / 4), 256, 8), 256, 0, as_stream(stream)>>>(
        dt, n, w, b, d / 4, reinterpret_cast<float4 *>(out));
  else
    time2vec_kernel<<<grid_for(n, 8, 8), 256, 0, as_stream(stream)>>>(dt, n, w, b, d, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
