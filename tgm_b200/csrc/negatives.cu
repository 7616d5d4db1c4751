// Negative destinations for a whole window of loader batches in ONE launch.
//
// Replaces, for the seed-producing step of link prediction, the per-batch call
//   torch.randint(low, high, (n,), dtype=int32, device=dg.device)
// of RandomNegativeEdgeSamplerHook (reference tgm-team/tgm @ 5183dc9,
// tgm/hooks/negatives/sampler.py:45-65).  The stream of numbers is documented and deliberately the
// one the reference's own `device='cuda'` mode draws: ATen's randint kernel gives element i of a
// call made at Philox offset `off` (a multiple of 4, numel <= 256 * grid, range < 2^28) the value
//   low + philox4x32_10(key = seed, counter = {off / 4, subsequence = i}).x % (high - low)
// and every call advances the generator's offset by 4.  Batch j of the window therefore uses
// offset off0 + 4 * j: the output equals `num_batches` consecutive torch.randint calls from a
// generator at (seed, off0), and the caller advances the generator by 4 * num_batches
// (tests/test_gpu_negatives.py pins the equality on hardware).
#include "common.cuh"

using namespace tgm;

namespace {

__device__ __forceinline__ uint32_t philox4x32_10_x(uint32_t k0, uint32_t k1, uint32_t c0,
                                                    uint32_t c1, uint32_t c2, uint32_t c3) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
    k0 += W0, k1 += W1;
  }
  return c0;
}

__global__ void __launch_bounds__(256)
negatives_window_kernel(uint64_t seed, uint64_t offset, int64_t low, uint64_t range,
                        int64_t per_batch, int64_t total, int32_t *__restrict__ out) {
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += int64_t(gridDim.x) * blockDim.x) {
    const int64_t j = g / per_batch;           // batch of the window = torch call number
    const uint64_t i = uint64_t(g - j * per_batch);  // element of that call = Philox subsequence
    const uint64_t ctr = (offset >> 2) + uint64_t(j);
    const uint32_t r = philox4x32_10_x(k0, k1, uint32_t(ctr), uint32_t(ctr >> 32), uint32_t(i),
                                       uint32_t(i >> 32));
    out[g] = int32_t(int64_t(uint64_t(r) % range) + low);
  }
}

}  // namespace

extern "C" int tgm_negatives_window(uint64_t seed, uint64_t offset, int64_t low, int64_t high,
                                    int64_t per_batch, int64_t total, int32_t *out,
                                    tgm_stream stream) {
  TGM_REQUIRE(low < high, "tgm_negatives_window: low must be < high");
  // from 2^28 on ATen draws 64-bit numbers (two Philox words per element) to bound the modulo bias
  TGM_REQUIRE(high - low < (int64_t(1) << 28), "tgm_negatives_window: range must be < 2^28");
  TGM_REQUIRE(low >= INT32_MIN && high - 1 <= INT32_MAX,
              "tgm_negatives_window: [low, high) must fit int32");
  TGM_REQUIRE((offset & 3u) == 0, "tgm_negatives_window: the Philox offset must be a multiple of 4");
  // one ATen call covers numel <= 256 * grid elements with one curand4 per thread; above that its
  // threads loop and the element -> (subsequence, lane) map changes
  TGM_REQUIRE(per_batch >= 1 && per_batch <= 65536,
              "tgm_negatives_window: per_batch must be in [1, 65536]");
  TGM_REQUIRE(total >= 0, "tgm_negatives_window: total must be >= 0");
  if (total == 0) return TGM_OK;
  TGM_REQUIRE(out != nullptr, "tgm_negatives_window: out is NULL");
  negatives_window_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(
      seed, offset, low, uint64_t(high - low), per_batch, total, out);
  TGM_LAUNCH_CHECK();
  return TGM_OK;
}
