// Device-resident time-sorted COO edge store + error plumbing.
// Replaces the edge arrays / slice lookup of tgm/core/_storage/backends/array_backend.py
// (reference tgm-team/tgm @ 5183dc9): _binary_search :301-321, get_edges :57-68,
// get_edge_x :259-268.  A slice is two binary searches over the timestamps (a host mirror when the
// caller has one, else a two-thread device search whose 16-byte answer is read back) and a pointer
// offset into the device slabs (no O(E) masks, no per-batch H2D).
#include "store.cuh"

#include <algorithm>
#include <cstring>
#include <new>

namespace tgm {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in `%s`", int(e), cudaGetErrorString(e),
           file, line, what);
  g_last_error = buf;
  cudaGetLastError();  // clear sticky-less errors so later calls report their own
  return e == cudaErrorMemoryAllocation ? TGM_ERR_OOM : TGM_ERR_CUDA;
}

}  // namespace tgm

using namespace tgm;

extern "C" const char *tgm_last_error(void) { return g_last_error.c_str(); }
extern "C" int tgm_version(void) { return 100; }  // 0.1.0
extern "C" int tgm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

tgm_store::~tgm_store() {
  if (scratch && device >= 0) {
    DeviceGuard g(device);
    cudaFree(scratch);
  }
  if (owns_device && device >= 0) {
    DeviceGuard g(device);
    cudaFree(const_cast<int32_t *>(src));
    cudaFree(const_cast<int32_t *>(dst));
    cudaFree(const_cast<int64_t *>(t));
    cudaFree(const_cast<float *>(x));
  }
}

namespace {
// number of adjacent pairs out of order (0 = the stream is time-sorted)
__global__ void store_unsorted_kernel(const int64_t *__restrict__ t, int64_t E,
                                      unsigned long long *__restrict__ bad) {
  unsigned long long mine = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i + 1 < E;
       i += int64_t(gridDim.x) * blockDim.x)
    mine += t[i] > t[i + 1];
  mine = __reduce_add_sync(0xffffffffu, unsigned(mine));
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(bad, mine);
}
// thread 0: first index with t >= t_lo; thread 1: first index with t > t_hi
__global__ void store_bounds_kernel(const int64_t *__restrict__ t, int64_t E, int64_t t_lo,
                                    int64_t t_hi, int64_t *__restrict__ out) {
  const bool upper = threadIdx.x == 1;
  const int64_t key = upper ? t_hi : t_lo;
  int64_t lo = 0, hi = E;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = t[mid];
    if (upper ? v <= key : v < key) lo = mid + 1; else hi = mid;
  }
  out[threadIdx.x] = lo;
}
}  // namespace

extern "C" int tgm_store_create(tgm_store **out, const int32_t *src, const int32_t *dst,
                                const int64_t *t, const float *edge_x, int64_t E, int32_t D,
                                int32_t num_nodes, int device, int mem, const int64_t *t_host) {
  TGM_REQUIRE(out != nullptr, "tgm_store_create: out is NULL");
  *out = nullptr;
  TGM_REQUIRE(E >= 0 && E < (int64_t(1) << 31), "tgm_store_create: E must be in [0, 2^31)");
  TGM_REQUIRE(D >= 0, "tgm_store_create: D must be >= 0");
  TGM_REQUIRE(num_nodes >= 0, "tgm_store_create: num_nodes must be >= 0");
  TGM_REQUIRE(mem == TGM_MEM_HOST || mem == TGM_MEM_DEVICE, "tgm_store_create: bad mem kind");
  TGM_REQUIRE(E == 0 || (src && dst && t), "tgm_store_create: src/dst/t must be non-NULL");
  TGM_REQUIRE((D == 0) == (edge_x == nullptr) || E == 0,
              "tgm_store_create: edge_x must be NULL exactly when D == 0");
  TGM_REQUIRE(device >= 0 || mem == TGM_MEM_HOST,
              "tgm_store_create: a metadata-only store (device < 0) needs host arrays");

  tgm_store *s = new (std::nothrow) tgm_store();
  if (!s) return fail(TGM_ERR_OOM, "tgm_store_create: host allocation failed");
  s->E = E;
  s->D = D;
  s->num_nodes = num_nodes;
  s->device = device;

  auto bail = [&](int code) {
    delete s;
    return code;
  };

  if (mem == TGM_MEM_HOST) {
    // the caller's arrays may be temporaries: the timestamps are copied into an owned mirror
    try {
      s->t_owned.assign(t, t + size_t(E));
    } catch (...) {
      return bail(fail(TGM_ERR_OOM, "tgm_store_create: host mirror allocation failed"));
    }
    s->t_host = s->t_owned.data();
    if (device >= 0) {
      DeviceGuard g(device);
      if (!g.ok) return bail(fail(TGM_ERR_CUDA, "tgm_store_create: cannot select device"));
      s->owns_device = true;
      size_t n = size_t(E);
      int32_t *dsrc = nullptr, *ddst = nullptr;
      int64_t *dt = nullptr;
      float *dx = nullptr;
      cudaError_t e = cudaSuccess;
      if (n) {
        if ((e = cudaMalloc(&dsrc, n * 4)) == cudaSuccess) s->src = dsrc;
        if (e == cudaSuccess && (e = cudaMalloc(&ddst, n * 4)) == cudaSuccess) s->dst = ddst;
        if (e == cudaSuccess && (e = cudaMalloc(&dt, n * 8)) == cudaSuccess) s->t = dt;
        if (e == cudaSuccess && D > 0 && (e = cudaMalloc(&dx, n * size_t(D) * 4)) == cudaSuccess)
          s->x = dx;
        if (e == cudaSuccess) e = cudaMemcpy(dsrc, src, n * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ddst, dst, n * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dt, t, n * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && D > 0)
          e = cudaMemcpy(dx, edge_x, n * size_t(D) * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
          return bail(cuda_fail(e, "store upload", __FILE__, __LINE__));
      }
    }
    if (!std::is_sorted(s->t_owned.begin(), s->t_owned.end()))
      return bail(fail(TGM_ERR_INVALID, "tgm_store_create: timestamps must be non-decreasing"));
  } else {
    DeviceGuard g(device);
    if (!g.ok) return bail(fail(TGM_ERR_CUDA, "tgm_store_create: cannot select device"));
    s->src = src;
    s->dst = dst;
    s->t = t;
    s->x = edge_x;
    // t_host, when given, is BORROWED (it must outlive the store); without it the store keeps no
    // host mirror at all: order is verified and bounds are searched on the device
    s->t_host = t_host;
    cudaError_t e = cudaMalloc(&s->scratch, 2 * sizeof(int64_t));
    if (e != cudaSuccess) return bail(cuda_fail(e, "store scratch", __FILE__, __LINE__));
    if (E > 1) {
      e = cudaMemset(s->scratch, 0, 2 * sizeof(int64_t));
      if (e != cudaSuccess) return bail(cuda_fail(e, "store scratch", __FILE__, __LINE__));
      store_unsorted_kernel<<<grid_for(E - 1, 256 * 8, 8), 256>>>(
          t, E, reinterpret_cast<unsigned long long *>(s->scratch));
      int64_t bad = 0;
      e = cudaMemcpy(&bad, s->scratch, sizeof bad, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return bail(cuda_fail(e, "timestamp order check", __FILE__, __LINE__));
      if (bad) return bail(fail(TGM_ERR_INVALID, "tgm_store_create: timestamps must be non-decreasing"));
    }
  }
  *out = s;
  return TGM_OK;
}

extern "C" void tgm_store_destroy(tgm_store *s) { delete s; }

extern "C" int tgm_store_info(const tgm_store *s, int64_t *E, int32_t *D, int32_t *num_nodes,
                              int *device) {
  TGM_REQUIRE(s != nullptr, "tgm_store_info: store is NULL");
  if (E) *E = s->E;
  if (D) *D = s->D;
  if (num_nodes) *num_nodes = s->num_nodes;
  if (device) *device = s->device;
  return TGM_OK;
}

extern "C" int tgm_store_bounds(const tgm_store *s, int64_t t_lo, int has_lo, int64_t t_hi,
                                int has_hi, int64_t idx_lo, int64_t idx_hi, int64_t *lb,
                                int64_t *ub) {
  TGM_REQUIRE(s != nullptr && lb != nullptr && ub != nullptr, "tgm_store_bounds: NULL argument");
  int64_t lo = 0, hi = s->E;
  if (s->t_host) {
    const int64_t *ts = s->t_host, *te = s->t_host + s->E;
    if (has_lo) lo = std::lower_bound(ts, te, t_lo) - ts;
    if (has_hi) hi = std::upper_bound(ts, te, t_hi) - ts;
  } else if ((has_lo || has_hi) && s->E > 0) {  // no host mirror: search on the device
    DeviceGuard g(s->device);
    int64_t got[2];
    store_bounds_kernel<<<1, 2>>>(s->t, s->E, t_lo, t_hi, s->scratch);
    TGM_LAUNCH_CHECK();
    TGM_CUDA(cudaMemcpy(got, s->scratch, sizeof got, cudaMemcpyDeviceToHost));
    if (has_lo) lo = got[0];
    if (has_hi) hi = got[1];
  }
  int64_t cl = idx_lo < 0 ? 0 : idx_lo;
  int64_t ch = idx_hi < 0 ? s->E : idx_hi;
  // clamp(x, cl, ch) = max(cl, min(ch, x))  (array_backend.py:318-320)
  *lb = std::max(cl, std::min(ch, lo));
  *ub = std::max(cl, std::min(ch, hi));
  return TGM_OK;
}

extern "C" int tgm_store_slab(const tgm_store *s, int64_t lb, int64_t ub, const int32_t **src,
                              const int32_t **dst, const int64_t **t, const float **x) {
  TGM_REQUIRE(s != nullptr, "tgm_store_slab: store is NULL");
  if (s->device < 0) return fail(TGM_ERR_NO_DEVICE, "tgm_store_slab: metadata-only store");
  TGM_REQUIRE(0 <= lb && lb <= ub && ub <= s->E, "tgm_store_slab: bounds outside [0, E]");
  if (src) *src = s->src + lb;
  if (dst) *dst = s->dst + lb;
  if (t) *t = s->t + lb;
  if (x) *x = s->x ? s->x + size_t(lb) * size_t(s->D) : nullptr;
  return TGM_OK;
}
