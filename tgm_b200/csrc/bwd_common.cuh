// Small reductions shared by the TGN backward passes (tgn_memory.cu, graph_attn.cu).
// `static`: each translation unit gets its own copy (the library is built without -rdc).
#pragma once

#include "common.cuh"

namespace tgm {

// out[c] += sum_r A[r, c]   (A row-major [n, cols] with leading dimension ld).
// blockIdx.x tiles the columns (threads = consecutive columns: coalesced), blockIdx.y strides the
// rows; one atomicAdd per thread at the end.
static __global__ void __launch_bounds__(128)
colsum_add_kernel(const float *__restrict__ A, int64_t n, int cols, int ld,
                  float *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) acc += A[r * ld + c];
  atomicAdd(out + c, acc);
}

inline dim3 colsum_grid(int64_t n, int cols) {
  const int gx = (cols + 127) / 128;
  int64_t gy = (n + 15) / 16;  // >= 16 rows per thread
  const int64_t cap = (int64_t(kSmCount) * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(unsigned(gx), unsigned(gy));
}

// Time2Vec parameter gradients.  enc[r, c] = cos(arg), arg = fl32(fma(dt[r], w[c], b[c])):
//     g = -sin(arg) * d_enc[r, c];   gw[c] += g * dt[r];   gb[c] += g
// dt is read as dt[r * dt_stride]; rows with valid != nullptr and valid[r * dt_stride] == 0 are
// skipped (nodes without a message).  Same launch shape as colsum_add_kernel.
static __global__ void __launch_bounds__(128)
t2v_grad_kernel(const float *__restrict__ dt, const float *__restrict__ valid, int dt_stride,
                const float *__restrict__ d_enc, int ld, int64_t n, int TD,
                const float *__restrict__ w, const float *__restrict__ b,
                float *__restrict__ gw, float *__restrict__ gb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= TD) return;
  const float wc = __ldg(w + c), bc = __ldg(b + c);
  float aw = 0.f, ab = 0.f;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
    if (valid != nullptr && valid[r * dt_stride] == 0.f) continue;
    const float x = dt[r * dt_stride];
    const float g = -sinf(__fmaf_rn(x, wc, bc)) * d_enc[r * ld + c];
    aw = __fmaf_rn(g, x, aw);
    ab += g;
  }
  atomicAdd(gw + c, aw);
  atomicAdd(gb + c, ab);
}

}  // namespace tgm
