// Shared helpers for the tgm_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/tgm_b200.h"

namespace tgm {

constexpr int kWarp = 32;
constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

// RAII device guard: every entry point runs on the handle's device and restores the caller's.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

// gemm_fastf32.cu: fp32-accurate GEMM on the tcgen05 tensor cores (CUTLASS FastF32 collective).
// out[S,N] = act(A[S,K] W[N,K]^T + bias[N] (+ residual[S,N])), act = exact GELU when `gelu`;
// residual may alias out.  Returns 1 = computed, 0 = not applicable (caller uses cuBLAS), < 0 = error.
bool fastf32_available();
int fastf32_linear(int64_t S, int N, int K, const float *A, const float *W, const float *bias,
                   const float *residual, int gelu, float *out, cudaStream_t stream);
extern int g_gemm_fastf32;  // tgm_set_option("gemm_fastf32", 0|1); default 1
// tc_linear.cu: the same contract on a hand-written tcgen05 kernel (fp32 operands split in flight
// into two TF32 terms, three products accumulated in TMEM).  N % 4 == 0, K % 4 == 0.  `gelu`: 0 no
// activation, 1 exact GELU, 2 ReLU.
int tc3_linear(int64_t S, int N, int K, const float *A, const float *W, const float *bias,
               const float *residual, int gelu, float *out, cudaStream_t stream);
// small_gemm.cu: C[M,N] = act(A[M,K] W[N,K]^T + bias) for short matrices (32x32 SIMT tiles, fused
// epilogue, optional batch via strides); act 0 none / 2 ReLU; bias may be NULL.  1 = computed.
constexpr int kSmallGemmMaxGroups = 8;
struct SmallGemmGroups {  // group i: batch[i] products A_i[z] (M x K_i) W_i^T -> C_i[z] (M x N)
  const float *A[kSmallGemmMaxGroups], *W[kSmallGemmMaxGroups], *bias[kSmallGemmMaxGroups];
  float *C[kSmallGemmMaxGroups];
  int K[kSmallGemmMaxGroups], lda[kSmallGemmMaxGroups], ldw[kSmallGemmMaxGroups],
      batch[kSmallGemmMaxGroups];
  int64_t strideA[kSmallGemmMaxGroups], strideC[kSmallGemmMaxGroups];
  bool vec[kSmallGemmMaxGroups];  // filled by small_gemm_groups
  int n;
};
int small_gemm_groups(const SmallGemmGroups &g, int64_t M, int N, int ldc, int act, cudaStream_t st);
int small_gemm_nt(int64_t M, int N, int K, const float *A, int lda, int64_t strideA, const float *W,
                  int ldw, int64_t strideW, const float *bias, int act, float *C, int ldc,
                  int64_t strideC, int batch, cudaStream_t st);
extern int g_tc_bn;      // probe switch: forced column tile width of tc3_linear (0 = automatic)
extern int g_tc_linear;  // tgm_set_option("tc_linear", 0|1|2); default 1 (see tc_linear.cu)
extern int g_attn_folded;  // tgm_set_option("attn_folded", 0|1); default 1 (attn_fold.cu)
extern int g_dyg_fused_attn;  // tgm_set_option("dyg_fused_attn", 0|1); default 1

inline cudaStream_t as_stream(tgm_stream s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid size for a grid-stride kernel: enough CTAs for `work_items`, capped at a whole number of
// waves over the 148 SMs.
inline int grid_for(int64_t work_items, int items_per_cta, int ctas_per_sm) {
  int64_t need = (work_items + items_per_cta - 1) / items_per_cta;
  int64_t cap = int64_t(kSmCount) * ctas_per_sm;
  if (need < 1) need = 1;
  return int(need < cap ? need : cap);
}

}  // namespace tgm

#define TGM_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return tgm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define TGM_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) return tgm::fail(TGM_ERR_INVALID, (msg));      \
  } while (0)

#define TGM_LAUNCH_CHECK() TGM_CUDA(cudaGetLastError())

// ---- device helpers ------------------------------------------------------------------------
namespace tgm {

// One adjacency / ring entry as it travels through the sampler: 16 bytes, one 128-bit load.
struct __align__(16) Entry {
  int32_t nbr;  // neighbour node id
  int32_t eid;  // edge index in the store (feature row)
  int64_t t;    // edge timestamp
};
static_assert(sizeof(Entry) == 16, "Entry must be one 16-byte vector");

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming 128-bit load that does not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// one adjacency entry as a single streaming 128-bit load
__device__ __forceinline__ Entry ldg_stream_entry(const Entry *p) {
  int32_t a, b;
  int64_t c;
  asm volatile("{\n\t.reg .b64 lo;\n\t"
               "ld.global.nc.L1::no_allocate.v2.b64 {lo, %2}, [%3];\n\t"
               "mov.b64 {%0, %1}, lo;\n\t}"
               : "=r"(a), "=r"(b), "=l"(c)
               : "l"(p));
  Entry e;
  e.nbr = a;
  e.eid = b;
  e.t = c;
  return e;
}
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src_lane) {
  int lo = __shfl_sync(0xffffffffu, int(uint64_t(v) & 0xffffffffull), src_lane);
  int hi = __shfl_sync(0xffffffffu, int(uint64_t(v) >> 32), src_lane);
  return int64_t((uint64_t(uint32_t(hi)) << 32) | uint64_t(uint32_t(lo)));
}
// cos(a) for the Time2Vec argument a = fl32(fma(dt, w, b)), |error| <= 1.6e-7 for |a| < 2^22
// (checked against float64 over 1e7 arguments up to 4e6): q = rint(a / pi) by the magic-number
// add, a three-term Cody-Waite reduction with FMAs (pi = C1 - D1 - D2), a degree-14 Taylor
// polynomial on |r| <~ 1.75 and the sign from the parity of q.  15 instructions against ~47 for
// libdevice cosf's fast path; larger arguments take cosf.
// the reduced-range path alone: the caller guarantees |a| < 4194304
__device__ __forceinline__ float t2v_cos_fast(float a) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float t = __fmaf_rn(a, 0.318309886183790672f, kMagic);
  const float q = t - kMagic;
  float r = __fmaf_rn(q, -3.1415927410125732f, a);
  r = __fmaf_rn(q, 8.742277657347586e-08f, r);
  r = __fmaf_rn(q, 3.4302490200117637e-15f, r);
  const float x2 = r * r;
  float p = -1.1470745597729725e-11f;          // -1/14!
  p = __fmaf_rn(p, x2, 2.08767569878681e-09f);  //  1/12!
  p = __fmaf_rn(p, x2, -2.755731922398589e-07f);
  p = __fmaf_rn(p, x2, 2.48015873015873e-05f);
  p = __fmaf_rn(p, x2, -1.388888888888889e-03f);
  p = __fmaf_rn(p, x2, 4.1666666666666664e-02f);
  p = __fmaf_rn(p, x2, -0.5f);
  p = __fmaf_rn(p, x2, 1.0f);
  return __uint_as_float(__float_as_uint(p) ^ ((__float_as_uint(t) & 1u) << 31));
}
__device__ __forceinline__ float t2v_cos(float a) {
  if (!(fabsf(a) < 4194304.f)) return cosf(a);
  return t2v_cos_fast(a);
}

__device__ __forceinline__ void stg_stream_f4(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace tgm
