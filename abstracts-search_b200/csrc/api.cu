// api.cu — the extern "C" surface of libabsb200.so (include/absb200.h) for the index half of the
// path: library/device queries, the synthetic generator, IndexFlatIP and IndexIVFFlat.
// The encoder entry points live in encoder.cu.
//
// Every function is a thin shell: validate, pick the stream, stage host buffers, call the C++
// object, translate exceptions into error codes.  No exception crosses the boundary.
#include <mutex>

#include "common.cuh"
#include "ivf.cuh"

namespace absb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }

DeviceProps device_props(int device) {
  static std::mutex mu;
  static std::vector<DeviceProps> cache;
  std::lock_guard<std::mutex> lock(mu);
  int count = 0;
  ABSB_CUDA(cudaGetDeviceCount(&count));
  ABSB_CHECK(device >= 0 && device < count, ABSB_ERR_INVALID, "device %d out of range (count=%d)", device, count);
  if ((int)cache.size() < count) cache.resize(count);
  DeviceProps& p = cache[device];
  if (p.sm_count == 0) {
    int v = 0;
    ABSB_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, device));
    ABSB_CUDA(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, device));
    ABSB_CUDA(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    ABSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    p.smem_optin = (size_t)v;
  }
  return p;
}

// The library carries sm_100a code only: anything else cannot run it, and there is no fallback.
static void require_sm100(int device) {
  const DeviceProps p = device_props(device);
  ABSB_CHECK(p.cc_major == 10, ABSB_ERR_UNSUPPORTED,
             "device %d is sm_%d%d; libabsb200 is built for sm_100a (B200) only and has no fallback",
             device, p.cc_major, p.cc_minor);
}

namespace {

// Stages a host array into a device workspace on st (async; caller syncs before reusing `src`).
template <typename T>
T* stage_in(DBuf<T>& ws, const T* src, size_t n, cudaStream_t st) {
  ws.reserve(std::max<size_t>(n, 1));
  if (n) ABSB_CUDA(cudaMemcpyAsync(ws.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return ws.p;
}

template <typename T>
void fetch_out(T* dst, const T* src_dev, size_t n, cudaStream_t st) {
  if (n) ABSB_CUDA(cudaMemcpyAsync(dst, src_dev, n * sizeof(T), cudaMemcpyDeviceToHost, st));
}

}  // namespace
}  // namespace absb

using namespace absb;

struct absb_ivf_s {
  IvfIndex ix;
  absb_ivf_s(int d, int nlist, int device) : ix(d, nlist, device) {}
};
struct absb_flat_s {
  FlatIndex ix;
  absb_flat_s(int d, int device) : ix(d, device) {}
};

struct absb_peer_s {
  PeerExchange px;
  absb_peer_s(int device, int rank, int world, size_t slot_bytes) : px(device, rank, world, slot_bytes) {}
};

#define NEED(p) ABSB_CHECK((p) != nullptr, ABSB_ERR_INVALID, "null argument: " #p)

namespace absb {
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ABSB_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace absb

extern "C" {

// ------------------------------------------------------------------ library -----------------
int absb_version(void) { return 100; }

const char* absb_last_error(void) { return g_last_error.c_str(); }

int absb_device_count(int* count) {
  ABSB_API_BEGIN
  NEED(count);
  ABSB_CUDA(cudaGetDeviceCount(count));
  ABSB_API_END
}

int absb_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
  ABSB_API_BEGIN
  const DeviceProps p = device_props(device);
  if (name && name_len > 0) {
    cudaDeviceProp prop;
    ABSB_CUDA(cudaGetDeviceProperties(&prop, device));
    snprintf(name, (size_t)name_len, "%s", prop.name);
  }
  if (sm_count) *sm_count = p.sm_count;
  if (cc_major) *cc_major = p.cc_major;
  if (cc_minor) *cc_minor = p.cc_minor;
  ABSB_API_END
}

// ------------------------------------------------------------------ synthetic ---------------
int absb_synth_fill_dev(int kind, uint64_t seed, int64_t row0, int64_t n, int d, int nlist,
                        int64_t corpus_rows, float* out_dev, void* stream) {
  ABSB_API_BEGIN
  ABSB_CHECK(n == 0 || out_dev, ABSB_ERR_INVALID, "null output");
  synth_fill(kind, seed, row0, nullptr, n, d, nlist, corpus_rows, out_dev, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_synth_fill_rows_dev(int kind, uint64_t seed, const int64_t* rows_dev, int64_t n, int d, int nlist,
                             int64_t corpus_rows, float* out_dev, void* stream) {
  ABSB_API_BEGIN
  ABSB_CHECK(n == 0 || (out_dev && rows_dev), ABSB_ERR_INVALID, "null argument");
  synth_fill(kind, seed, 0, reinterpret_cast<const long long*>(rows_dev), n, d, nlist, corpus_rows, out_dev,
             (cudaStream_t)stream);
  ABSB_API_END
}

int absb_synth_cluster_dev(uint64_t seed, int64_t row0, int64_t n, int nlist, int64_t* out_dev,
                           void* stream) {
  ABSB_API_BEGIN
  ABSB_CHECK(n == 0 || out_dev, ABSB_ERR_INVALID, "null output");
  ABSB_CHECK(nlist > 0, ABSB_ERR_INVALID, "nlist=%d", nlist);
  synth_cluster(seed, row0, n, nlist, reinterpret_cast<long long*>(out_dev), (cudaStream_t)stream);
  ABSB_API_END
}

// ------------------------------------------------------------------ IndexFlatIP -------------
int absb_flat_create(int d, int metric, int device, absb_flat_t* out) {
  ABSB_API_BEGIN
  NEED(out);
  ABSB_CHECK(metric == ABSB_METRIC_INNER_PRODUCT, ABSB_ERR_UNSUPPORTED,
             "only METRIC_INNER_PRODUCT is on the abstracts-search path (metric=%d)", metric);
  require_sm100(device);
  *out = new absb_flat_s(d, device);
  ABSB_API_END
}

int absb_flat_destroy(absb_flat_t h) {
  ABSB_API_BEGIN
  delete h;
  ABSB_API_END
}

int absb_flat_reset(absb_flat_t h) {
  ABSB_API_BEGIN
  NEED(h);
  DeviceGuard g(h->ix.device);
  ABSB_CUDA(cudaStreamSynchronize(h->ix.own_stream));
  h->ix.xb.release();
  h->ix.ntotal = 0;
  ABSB_API_END
}

int absb_flat_ntotal(absb_flat_t h, int64_t* ntotal) {
  ABSB_API_BEGIN
  NEED(h); NEED(ntotal);
  *ntotal = h->ix.ntotal;
  ABSB_API_END
}

int absb_flat_add_dev(absb_flat_t h, int64_t n, const float* x_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || x_dev), ABSB_ERR_INVALID, "bad add arguments");
  DeviceGuard g(h->ix.device);
  h->ix.add_dev(n, x_dev, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_flat_add(absb_flat_t h, int64_t n, const float* x) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || x), ABSB_ERR_INVALID, "bad add arguments");
  FlatIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  const int64_t step = std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)ix.d * 4));
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* xd = stage_in(ix.ws_x, x + r0 * ix.d, (size_t)nr * ix.d, st);
    ix.add_dev(nr, xd, st);
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  ABSB_API_END
}

int absb_flat_search_dev(absb_flat_t h, int64_t n, const float* q_dev, int k, float* D_dev,
                         int64_t* I_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q_dev && D_dev && I_dev)), ABSB_ERR_INVALID, "bad search arguments");
  DeviceGuard g(h->ix.device);
  h->ix.search_dev(n, q_dev, k, D_dev, reinterpret_cast<long long*>(I_dev), (cudaStream_t)stream);
  ABSB_API_END
}

int absb_flat_search(absb_flat_t h, int64_t n, const float* q, int k, float* D, int64_t* I) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q && D && I)), ABSB_ERR_INVALID, "bad search arguments");
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  FlatIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  const int64_t step = 8192;
  ix.ws_D.reserve((size_t)std::min(step, std::max<int64_t>(n, 1)) * k);
  ix.ws_I.reserve((size_t)std::min(step, std::max<int64_t>(n, 1)) * k);
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* qd = stage_in(ix.ws_x, q + r0 * ix.d, (size_t)nr * ix.d, st);
    ix.search_dev(nr, qd, k, ix.ws_D.p, ix.ws_I.p, st);
    fetch_out(D + r0 * k, ix.ws_D.p, (size_t)nr * k, st);
    fetch_out(reinterpret_cast<long long*>(I) + r0 * k, ix.ws_I.p, (size_t)nr * k, st);
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  ABSB_API_END
}

int absb_flat_reconstruct(absb_flat_t h, int64_t i0, int64_t n, float* x) {
  ABSB_API_BEGIN
  NEED(h);
  FlatIndex& ix = h->ix;
  ABSB_CHECK(i0 >= 0 && n >= 0 && i0 + n <= ix.ntotal && (n == 0 || x), ABSB_ERR_INVALID,
             "reconstruct range [%lld,%lld) outside [0,%lld)", (long long)i0, (long long)(i0 + n), (long long)ix.ntotal);
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaStreamSynchronize(ix.own_stream));
  if (n) ABSB_CUDA(cudaMemcpy(x, ix.xb.p + (size_t)i0 * ix.d, sizeof(float) * (size_t)n * ix.d, cudaMemcpyDeviceToHost));
  ABSB_API_END
}

// ------------------------------------------------------------------ IndexIVFFlat ------------
int absb_ivf_create(int d, int nlist, int metric, int device, absb_ivf_t* out) {
  ABSB_API_BEGIN
  NEED(out);
  ABSB_CHECK(metric == ABSB_METRIC_INNER_PRODUCT, ABSB_ERR_UNSUPPORTED,
             "only METRIC_INNER_PRODUCT is on the abstracts-search path (metric=%d)", metric);
  require_sm100(device);
  *out = new absb_ivf_s(d, nlist, device);
  ABSB_API_END
}

int absb_ivf_destroy(absb_ivf_t h) {
  ABSB_API_BEGIN
  delete h;
  ABSB_API_END
}

int absb_ivf_reset(absb_ivf_t h) {
  ABSB_API_BEGIN
  NEED(h);
  h->ix.reset();
  ABSB_API_END
}

int absb_ivf_ntotal(absb_ivf_t h, int64_t* ntotal) {
  ABSB_API_BEGIN
  NEED(h); NEED(ntotal);
  *ntotal = h->ix.ntotal;
  ABSB_API_END
}

int absb_ivf_is_trained(absb_ivf_t h, int* trained) {
  ABSB_API_BEGIN
  NEED(h); NEED(trained);
  *trained = h->ix.trained ? 1 : 0;
  ABSB_API_END
}

int absb_ivf_set_clustering(absb_ivf_t h, int niter, int max_points_per_centroid,
                            int min_points_per_centroid, int64_t seed) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(niter >= 0 && max_points_per_centroid >= 1, ABSB_ERR_INVALID, "bad clustering parameters");
  h->ix.cp.niter = niter;
  h->ix.cp.max_points_per_centroid = max_points_per_centroid;
  h->ix.cp.min_points_per_centroid = min_points_per_centroid;
  h->ix.cp.seed = seed;
  ABSB_API_END
}

int absb_ivf_set_clustering_spherical(absb_ivf_t h, int spherical) {
  ABSB_API_BEGIN
  NEED(h);
  h->ix.cp.spherical = spherical != 0;
  ABSB_API_END
}

int absb_renorm_rows_dev(int device, int64_t n, int d, float* x_dev, void* stream) {
  ABSB_API_BEGIN
  ABSB_CHECK(n >= 0 && d > 0 && (n == 0 || x_dev), ABSB_ERR_INVALID, "bad renorm arguments");
  require_sm100(device);
  DeviceGuard g(device);
  renorm_rows(n, d, x_dev, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_train(absb_ivf_t h, int64_t n, const float* x) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || x), ABSB_ERR_INVALID, "bad train arguments");
  DeviceGuard g(h->ix.device);
  h->ix.train_host(n, x);
  ABSB_API_END
}

int absb_ivf_train_dev(absb_ivf_t h, int64_t n, const float* x_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || x_dev), ABSB_ERR_INVALID, "bad train arguments");
  DeviceGuard g(h->ix.device);
  cudaStream_t st = (cudaStream_t)stream;
  h->ix.train_dev(n, x_dev, st);
  ABSB_CUDA(cudaStreamSynchronize(st));
  ABSB_API_END
}

int absb_ivf_set_centroids(absb_ivf_t h, const float* centroids) {
  ABSB_API_BEGIN
  NEED(h); NEED(centroids);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaStreamSynchronize(ix.own_stream));
  ABSB_CUDA(cudaMemcpy(ix.centroids.p, centroids, sizeof(float) * (size_t)ix.nlist * ix.d, cudaMemcpyHostToDevice));
  ix.trained = true;
  ix.c3_dirty = true;
  ABSB_API_END
}

int absb_ivf_get_centroids(absb_ivf_t h, float* centroids) {
  ABSB_API_BEGIN
  NEED(h); NEED(centroids);
  IvfIndex& ix = h->ix;
  ABSB_CHECK(ix.trained, ABSB_ERR_STATE, "index is not trained");
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaStreamSynchronize(ix.own_stream));
  ABSB_CUDA(cudaMemcpy(centroids, ix.centroids.p, sizeof(float) * (size_t)ix.nlist * ix.d, cudaMemcpyDeviceToHost));
  ABSB_API_END
}

int absb_ivf_set_centroids_dev(absb_ivf_t h, const float* centroids_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h); NEED(centroids_dev);
  DeviceGuard g(h->ix.device);
  h->ix.set_centroids_dev(centroids_dev, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_add_dev(absb_ivf_t h, int64_t n, const float* x_dev, const int64_t* ids_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || x_dev), ABSB_ERR_INVALID, "bad add arguments");
  DeviceGuard g(h->ix.device);
  h->ix.add_dev(n, x_dev, reinterpret_cast<const long long*>(ids_dev), (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_add_preassigned_dev(absb_ivf_t h, int64_t n, const float* x_dev, const int64_t* ids_dev,
                                 const int64_t* list_ids_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (x_dev && list_ids_dev)), ABSB_ERR_INVALID, "bad add arguments");
  DeviceGuard g(h->ix.device);
  const int64_t step = (int64_t)1 << 22;
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    h->ix.add_core_dev(nr, x_dev + r0 * h->ix.d,
                       ids_dev ? reinterpret_cast<const long long*>(ids_dev) + r0 : nullptr,
                       reinterpret_cast<const long long*>(list_ids_dev) + r0, (cudaStream_t)stream);
  }
  ABSB_API_END
}

static void ivf_add_host(IvfIndex& ix, int64_t n, const float* x, const int64_t* ids, const int64_t* list_ids) {
  ABSB_CHECK(n >= 0 && (n == 0 || x), ABSB_ERR_INVALID, "bad add arguments");
  ABSB_CHECK(ix.trained, ABSB_ERR_STATE, "index is not trained");
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  const int64_t step = std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)ix.d * 4));
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* xd = stage_in(ix.ws_x, x + r0 * ix.d, (size_t)nr * ix.d, st);
    const long long* idd = nullptr;
    if (ids) idd = stage_in(ix.ws_ids, reinterpret_cast<const long long*>(ids) + r0, (size_t)nr, st);
    if (list_ids) {
      ix.ws_list_ids.reserve((size_t)nr);
      ABSB_CUDA(cudaMemcpyAsync(ix.ws_list_ids.p, list_ids + r0, sizeof(long long) * nr, cudaMemcpyHostToDevice, st));
      ix.add_core_dev(nr, xd, idd, ix.ws_list_ids.p, st);
    } else {
      ix.add_dev(nr, xd, idd, st);
    }
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
}

int absb_ivf_add(absb_ivf_t h, int64_t n, const float* x, const int64_t* ids) {
  ABSB_API_BEGIN
  NEED(h);
  ivf_add_host(h->ix, n, x, ids, nullptr);
  ABSB_API_END
}

int absb_ivf_add_preassigned(absb_ivf_t h, int64_t n, const float* x, const int64_t* ids,
                             const int64_t* list_ids) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n == 0 || list_ids, ABSB_ERR_INVALID, "null list_ids");
  ivf_add_host(h->ix, n, x, ids, list_ids);
  ABSB_API_END
}

int absb_ivf_compact(absb_ivf_t h) {
  ABSB_API_BEGIN
  NEED(h);
  DeviceGuard g(h->ix.device);
  h->ix.compact(0, h->ix.own_stream);
  ABSB_API_END
}

int absb_ivf_compact_scratch(absb_ivf_t h, int64_t scratch_pages) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(scratch_pages >= 1, ABSB_ERR_INVALID, "scratch_pages=%lld", (long long)scratch_pages);
  DeviceGuard g(h->ix.device);
  h->ix.compact(scratch_pages, h->ix.own_stream);
  ABSB_API_END
}

int absb_plan_page_compaction(int64_t n, const int32_t* src, int64_t scratch_pages, int32_t* moves,
                              int64_t moves_cap, int64_t* phase_end, int64_t phases_cap, int64_t* n_moves,
                              int64_t* n_phases) {
  ABSB_API_BEGIN
  ABSB_CHECK(n >= 0 && (n == 0 || src) && n_moves && n_phases, ABSB_ERR_INVALID, "bad arguments");
  std::vector<PageMove> mv;
  std::vector<int64_t> pe;
  plan_page_compaction(std::vector<int>(src, src + n), scratch_pages, mv, pe);
  *n_moves = (int64_t)mv.size();
  *n_phases = (int64_t)pe.size();
  ABSB_CHECK((int64_t)mv.size() <= moves_cap && (int64_t)pe.size() <= phases_cap, ABSB_ERR_INVALID,
             "output too small: %lld moves, %lld phases", (long long)mv.size(), (long long)pe.size());
  for (size_t i = 0; i < mv.size(); ++i) {
    moves[2 * i] = mv[i].from;
    moves[2 * i + 1] = mv[i].to;
  }
  std::copy(pe.begin(), pe.end(), phase_end);
  ABSB_API_END
}

int absb_ivf_coarse_dev(absb_ivf_t h, int64_t n, const float* q_dev, int nprobe, float* Dc_dev,
                        int64_t* Ic_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q_dev && Dc_dev && Ic_dev)), ABSB_ERR_INVALID, "bad coarse arguments");
  ABSB_CHECK(nprobe >= 1 && nprobe <= h->ix.nlist, ABSB_ERR_INVALID, "nprobe=%d outside [1,nlist]", nprobe);
  DeviceGuard g(h->ix.device);
  h->ix.coarse_dev(n, q_dev, nprobe, Dc_dev, reinterpret_cast<long long*>(Ic_dev), true, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_coarse(absb_ivf_t h, int64_t n, const float* q, int nprobe, float* Dc, int64_t* Ic) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q && Dc && Ic)), ABSB_ERR_INVALID, "bad coarse arguments");
  IvfIndex& ix = h->ix;
  ABSB_CHECK(nprobe >= 1 && nprobe <= ix.nlist && nprobe <= ABSB_MAX_K, ABSB_ERR_INVALID,
             "nprobe=%d outside [1,min(nlist,%d)]", nprobe, ABSB_MAX_K);
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  const int64_t step = 4096;
  ix.ws_D.reserve((size_t)step * nprobe);
  ix.ws_I.reserve((size_t)step * nprobe);
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* qd = stage_in(ix.ws_x, q + r0 * ix.d, (size_t)nr * ix.d, st);
    ix.coarse_dev(nr, qd, nprobe, ix.ws_D.p, ix.ws_I.p, true, st);
    fetch_out(Dc + r0 * nprobe, ix.ws_D.p, (size_t)nr * nprobe, st);
    fetch_out(reinterpret_cast<long long*>(Ic) + r0 * nprobe, ix.ws_I.p, (size_t)nr * nprobe, st);
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  ABSB_API_END
}

int absb_ivf_assign(absb_ivf_t h, int64_t n, const float* x, int64_t* list_ids) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (x && list_ids)), ABSB_ERR_INVALID, "bad assign arguments");
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  const int64_t step = 16384;
  ix.ws_I.reserve((size_t)step);
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* xd = stage_in(ix.ws_x, x + r0 * ix.d, (size_t)nr * ix.d, st);
    ix.assign_dev(nr, xd, ix.ws_I.p, st);
    fetch_out(reinterpret_cast<long long*>(list_ids) + r0, ix.ws_I.p, (size_t)nr, st);
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  ABSB_API_END
}

int absb_ivf_search_dev(absb_ivf_t h, int64_t n, const float* q_dev, int k, int nprobe, float* D_dev,
                        int64_t* I_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q_dev && D_dev && I_dev)), ABSB_ERR_INVALID, "bad search arguments");
  DeviceGuard g(h->ix.device);
  h->ix.reset_stats();
  h->ix.search_dev(n, q_dev, k, nprobe, D_dev, reinterpret_cast<long long*>(I_dev), (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_search_preassigned_dev(absb_ivf_t h, int64_t n, const float* q_dev, int k, int nprobe,
                                    const int64_t* coarse_ids_dev, float* D_dev, int64_t* I_dev,
                                    void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n >= 0 && (n == 0 || (q_dev && D_dev && I_dev && coarse_ids_dev)), ABSB_ERR_INVALID, "bad search arguments");
  DeviceGuard g(h->ix.device);
  IvfIndex& ix = h->ix;
  ix.reset_stats();
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t q0 = 0; q0 < n; q0 += kMaxPlanQueries) {
    const int64_t nb = std::min<int64_t>(kMaxPlanQueries, n - q0);
    ix.search_preassigned_dev(nb, q_dev + q0 * ix.d, k, nprobe,
                              reinterpret_cast<const long long*>(coarse_ids_dev) + q0 * nprobe,
                              D_dev + q0 * k, reinterpret_cast<long long*>(I_dev) + q0 * k, st);
  }
  ABSB_API_END
}

static void ivf_search_host(IvfIndex& ix, int64_t n, const float* q, int k, int nprobe,
                            const int64_t* coarse, float* D, int64_t* I) {
  ABSB_CHECK(n >= 0 && (n == 0 || (q && D && I)), ABSB_ERR_INVALID, "bad search arguments");
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  ABSB_CHECK(nprobe >= 1, ABSB_ERR_INVALID, "nprobe=%d", nprobe);
  ABSB_CHECK(ix.trained, ABSB_ERR_STATE, "index is not trained");
  DeviceGuard g(ix.device);
  cudaStream_t st = ix.own_stream;
  ix.reset_stats();
  const int64_t step = 4096;
  ix.ws_D.reserve((size_t)std::min(step, std::max<int64_t>(n, 1)) * k);
  ix.ws_I.reserve((size_t)std::min(step, std::max<int64_t>(n, 1)) * k);
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    const float* qd = stage_in(ix.ws_x, q + r0 * ix.d, (size_t)nr * ix.d, st);
    if (coarse) {
      const long long* cd = stage_in(ix.ws_ids, reinterpret_cast<const long long*>(coarse) + r0 * nprobe,
                                     (size_t)nr * nprobe, st);
      for (int64_t q0 = 0; q0 < nr; q0 += kMaxPlanQueries) {
        const int64_t nb = std::min<int64_t>(kMaxPlanQueries, nr - q0);
        ix.search_preassigned_dev(nb, qd + q0 * ix.d, k, nprobe, cd + q0 * nprobe, ix.ws_D.p + q0 * k,
                                  ix.ws_I.p + q0 * k, st);
      }
    } else {
      ix.search_dev(nr, qd, k, nprobe, ix.ws_D.p, ix.ws_I.p, st);
    }
    fetch_out(D + r0 * k, ix.ws_D.p, (size_t)nr * k, st);
    fetch_out(reinterpret_cast<long long*>(I) + r0 * k, ix.ws_I.p, (size_t)nr * k, st);
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
}

int absb_ivf_search(absb_ivf_t h, int64_t n, const float* q, int k, int nprobe, float* D, int64_t* I) {
  ABSB_API_BEGIN
  NEED(h);
  ivf_search_host(h->ix, n, q, k, nprobe, nullptr, D, I);
  ABSB_API_END
}

int absb_ivf_search_preassigned(absb_ivf_t h, int64_t n, const float* q, int k, int nprobe,
                                const int64_t* coarse_ids, float* D, int64_t* I) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(n == 0 || coarse_ids, ABSB_ERR_INVALID, "null coarse_ids");
  ivf_search_host(h->ix, n, q, k, nprobe, coarse_ids, D, I);
  ABSB_API_END
}

int absb_ivf_list_sizes(absb_ivf_t h, int64_t* sizes) {
  ABSB_API_BEGIN
  NEED(h); NEED(sizes);
  std::copy(h->ix.h_list_size.begin(), h->ix.h_list_size.end(), sizes);
  ABSB_API_END
}

int absb_ivf_get_list(absb_ivf_t h, int64_t list_no, float* codes, int64_t* ids) {
  ABSB_API_BEGIN
  NEED(h);
  IvfIndex& ix = h->ix;
  ABSB_CHECK(list_no >= 0 && list_no < ix.nlist, ABSB_ERR_INVALID, "list %lld outside [0,%d)", (long long)list_no, ix.nlist);
  DeviceGuard g(ix.device);
  ix.get_list(list_no, codes, reinterpret_cast<long long*>(ids));
  ABSB_API_END
}

int absb_ivf_set_shard(absb_ivf_t h, int rank, int world) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(world >= 1 && rank >= 0 && rank < world, ABSB_ERR_INVALID, "shard rank=%d world=%d", rank, world);
  ABSB_CHECK(h->ix.rows_seen == 0, ABSB_ERR_STATE, "set_shard must precede the first add");
  h->ix.shard_rank = rank;
  h->ix.shard_world = world;
  ABSB_API_END
}

int absb_merge_shards_dev(int device, int world, int64_t n, int k, const float* D_all_dev,
                          const int64_t* I_all_dev, int64_t rank_stride_bytes, float* D_dev,
                          int64_t* I_dev, void* stream) {
  ABSB_API_BEGIN
  ABSB_CHECK(world >= 1 && n >= 0 && k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "bad merge arguments");
  ABSB_CHECK(n == 0 || (D_all_dev && I_all_dev && D_dev && I_dev), ABSB_ERR_INVALID, "null argument");
  ABSB_CHECK(rank_stride_bytes % 8 == 0 && (reinterpret_cast<uintptr_t>(I_all_dev) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(D_all_dev) & 3) == 0,
             ABSB_ERR_INVALID, "merge_shards: the int64 ids of every rank must be 8-byte aligned (stride %lld)",
             (long long)rank_stride_bytes);
  DeviceGuard g(device);
  const int64_t ds = rank_stride_bytes ? rank_stride_bytes : n * k * (int64_t)sizeof(float);
  const int64_t is = rank_stride_bytes ? rank_stride_bytes : n * k * (int64_t)sizeof(long long);
  merge_shards(world, n, k, D_all_dev, reinterpret_cast<const long long*>(I_all_dev), ds, is, D_dev,
               reinterpret_cast<long long*>(I_dev), (cudaStream_t)stream);
  ABSB_API_END
}

// ------------------------------------------------------------------ peer exchange ----------
int absb_peer_create(int device, int rank, int world, size_t slot_bytes, absb_peer_t* out) {
  ABSB_API_BEGIN
  NEED(out);
  require_sm100(device);
  DeviceGuard g(device);
  *out = new absb_peer_s(device, rank, world, slot_bytes);
  ABSB_API_END
}

int absb_peer_destroy(absb_peer_t p) {
  ABSB_API_BEGIN
  if (p) {
    DeviceGuard g(p->px.device);
    delete p;
  }
  ABSB_API_END
}

int absb_peer_ipc_handle(absb_peer_t p, void* handle64) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(handle64);
  DeviceGuard g(p->px.device);
  p->px.ipc_handle(handle64);
  ABSB_API_END
}

int absb_peer_local_ptr(absb_peer_t p, void** ptr_dev) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(ptr_dev);
  *ptr_dev = p->px.local;
  ABSB_API_END
}

int absb_peer_connect(absb_peer_t p, const void* handles) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(handles);
  DeviceGuard g(p->px.device);
  p->px.connect_ipc(handles);
  ABSB_API_END
}

int absb_peer_connect_ptrs(absb_peer_t p, void* const* ptrs_dev) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(ptrs_dev);
  DeviceGuard g(p->px.device);
  p->px.connect_ptrs(ptrs_dev);
  ABSB_API_END
}

int absb_peer_allgather_dev(absb_peer_t p, const void* src_dev, size_t bytes, void** gathered_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(src_dev);
  NEED(gathered_dev);
  DeviceGuard g(p->px.device);
  *gathered_dev = p->px.allgather(src_dev, bytes, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_peer_push_dev(absb_peer_t p, const void* src_dev, size_t bytes, void* stream) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(src_dev);
  DeviceGuard g(p->px.device);
  p->px.push(src_dev, bytes, (cudaStream_t)stream);
  ABSB_API_END
}

int absb_peer_wait_dev(absb_peer_t p, void** gathered_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(gathered_dev);
  DeviceGuard g(p->px.device);
  p->px.wait((cudaStream_t)stream);
  *gathered_dev = p->px.local_entry(p->px.epoch);
  ABSB_API_END
}

int absb_peer_status(absb_peer_t p, int* status) {
  ABSB_API_BEGIN
  NEED(p);
  NEED(status);
  DeviceGuard g(p->px.device);
  *status = p->px.read_status();
  ABSB_API_END
}

int absb_ivf_search_push_dev(absb_ivf_t h, absb_peer_t p, int64_t n, const float* q_dev, int k, int nprobe,
                             void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  NEED(p);
  ABSB_CHECK(n >= 1 && q_dev, ABSB_ERR_INVALID, "bad search arguments");
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  ABSB_CHECK(p->px.device == h->ix.device, ABSB_ERR_INVALID, "index and exchange live on different devices");
  const size_t i_bytes = ((size_t)n * k * sizeof(long long) + 15) & ~(size_t)15;
  ABSB_CHECK(i_bytes + (size_t)n * k * sizeof(float) <= p->px.slot_bytes, ABSB_ERR_INVALID,
             "record of %lld x %d results does not fit the exchange slot (%zu bytes)", (long long)n, k, p->px.slot_bytes);
  DeviceGuard g(h->ix.device);
  h->ix.reset_stats();
  SearchPush sp{p->px.begin_push(0, (long long)i_bytes), 0, n};
  h->ix.search_dev(n, q_dev, k, nprobe, nullptr, nullptr, (cudaStream_t)stream, &sp);
  p->px.commit();
  p->px.rec_n = n;
  p->px.rec_k = k;
  ABSB_API_END
}

int absb_ivf_search_preassigned_push_dev(absb_ivf_t h, absb_peer_t p, int64_t n, const float* q_dev, int k, int nprobe,
                                         const int64_t* coarse_ids_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h);
  NEED(p);
  ABSB_CHECK(n >= 1 && q_dev && coarse_ids_dev, ABSB_ERR_INVALID, "bad search arguments");
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  ABSB_CHECK(p->px.device == h->ix.device, ABSB_ERR_INVALID, "index and exchange live on different devices");
  const size_t i_bytes = ((size_t)n * k * sizeof(long long) + 15) & ~(size_t)15;
  ABSB_CHECK(i_bytes + (size_t)n * k * sizeof(float) <= p->px.slot_bytes, ABSB_ERR_INVALID,
             "record of %lld x %d results does not fit the exchange slot (%zu bytes)", (long long)n, k, p->px.slot_bytes);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ix.reset_stats();
  SearchPush sp{p->px.begin_push(0, (long long)i_bytes), 0, n};
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t q0 = 0; q0 < n; q0 += kMaxPlanQueries) {
    const int64_t nb = std::min<int64_t>(kMaxPlanQueries, n - q0);
    SearchPush sub = sp;
    sub.q_base = q0;
    ix.search_preassigned_dev(nb, q_dev + q0 * ix.d, k, nprobe,
                              reinterpret_cast<const long long*>(coarse_ids_dev) + q0 * nprobe, nullptr, nullptr, st, &sub);
  }
  p->px.commit();
  p->px.rec_n = n;
  p->px.rec_k = k;
  ABSB_API_END
}

int absb_peer_push_results_dev(absb_peer_t p, int64_t n, int k, const float* D_dev, const int64_t* I_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(p);
  ABSB_CHECK(n >= 1 && k >= 1 && k <= ABSB_MAX_K && D_dev && I_dev, ABSB_ERR_INVALID, "bad push arguments");
  PeerExchange& px = p->px;
  const size_t i_bytes = ((size_t)n * k * sizeof(long long) + 15) & ~(size_t)15;
  const size_t d_bytes = ((size_t)n * k * sizeof(float) + 15) & ~(size_t)15;
  ABSB_CHECK(i_bytes + d_bytes <= px.slot_bytes, ABSB_ERR_INVALID,
             "record of %lld x %d results does not fit the exchange slot (%zu bytes)", (long long)n, k, px.slot_bytes);
  DeviceGuard g(px.device);
  cudaStream_t st = (cudaStream_t)stream;
  px.staging.reserve(px.slot_bytes);
  ABSB_CUDA(cudaMemcpyAsync(px.staging.p, I_dev, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToDevice, st));
  ABSB_CUDA(cudaMemcpyAsync(px.staging.p + i_bytes, D_dev, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
  px.push(px.staging.p, i_bytes + d_bytes, st);
  px.rec_n = n;
  px.rec_k = k;
  ABSB_API_END
}

int absb_peer_merge_shards_dev(absb_peer_t p, int64_t n, int k, float* D_dev, int64_t* I_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(p);
  ABSB_CHECK(n >= 1 && D_dev && I_dev, ABSB_ERR_INVALID, "bad merge arguments");
  PeerExchange& px = p->px;
  ABSB_CHECK(px.epoch > 0 && px.rec_n == n && px.rec_k == k, ABSB_ERR_STATE,
             "no pushed search record of %lld x %d results to merge", (long long)n, k);
  DeviceGuard g(px.device);
  const size_t i_bytes = ((size_t)n * k * sizeof(long long) + 15) & ~(size_t)15;
  merge_shards_wait(px.world, n, k, px.local_entry(px.epoch), (int64_t)px.slot_bytes, 0, (int64_t)i_bytes,
                    px.local_flags(), px.epoch, px.status_dev, D_dev, reinterpret_cast<long long*>(I_dev),
                    (cudaStream_t)stream);
  ABSB_API_END
}

int absb_ivf_set_two_stage(absb_ivf_t h, int shortlist) {
  ABSB_API_BEGIN
  NEED(h);
  DeviceGuard g(h->ix.device);
  h->ix.set_two_stage(shortlist);
  ABSB_API_END
}

int absb_ivf_two_stage_fallbacks(absb_ivf_t h, int64_t* queries) {
  ABSB_API_BEGIN
  NEED(h);
  NEED(queries);
  DeviceGuard g(h->ix.device);
  *queries = h->ix.two_stage_fallbacks(nullptr);
  ABSB_API_END
}

int absb_ivf_set_tunables(absb_ivf_t h, int scan_chunk, int coarse_impl, int scan_ctas_per_sm) {
  ABSB_API_BEGIN
  NEED(h);
  if (scan_chunk > 0) h->ix.scan_chunk = scan_chunk;
  if (coarse_impl >= 0) {
    ABSB_CHECK(coarse_impl <= 1, ABSB_ERR_INVALID, "coarse_impl=%d", coarse_impl);
    h->ix.coarse_impl = coarse_impl;
  }
  if (scan_ctas_per_sm >= 0) h->ix.scan_ctas_per_sm = scan_ctas_per_sm;
  ABSB_API_END
}

int absb_ivf_set_scan_impl(absb_ivf_t h, int impl, int ring_warps, int ring_depth, int ring_stage_vecs) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(impl >= -1 && impl <= 2, ABSB_ERR_INVALID, "scan impl %d", impl);
  IvfIndex& ix = h->ix;
  if (impl >= 0) {
    ix.scan_impl = impl;
    ix.ring.small = impl == 2;
    ix.ring.l2_evict_first = impl == 2 && getenv("ABSB_SCAN_L2_DEFAULT") == nullptr;
    if (impl == 2) {  // the co-resident shape: ONE CTA of 8 warps x 2 stages x 4 KB per SM
      ix.ring.warps = 8;
      ix.ring.depth = 2;
      ix.ring.stage_vecs = 1;
      ix.scan_ctas_per_sm = 1;
    }
  }
  if (ring_warps > 0) {
    ABSB_CHECK(ring_warps <= 16, ABSB_ERR_INVALID, "ring warps %d", ring_warps);
    ix.ring.warps = ring_warps;
  }
  if (ring_depth > 0) {
    ABSB_CHECK(ring_depth >= 2 && ring_depth <= 7, ABSB_ERR_INVALID, "ring depth %d outside [2,7]", ring_depth);
    ix.ring.depth = ring_depth;
  }
  if (ring_stage_vecs > 0) {
    ABSB_CHECK(ring_stage_vecs == 1 || ring_stage_vecs == 2, ABSB_ERR_INVALID, "ring stage of %d fp32 vectors (1 or 2)",
               ring_stage_vecs);
    ix.ring.stage_vecs = ring_stage_vecs;
  }
  ABSB_API_END
}

int absb_ivf_set_scan_order(absb_ivf_t h, int list_major) {
  ABSB_API_BEGIN
  NEED(h);
  ABSB_CHECK(list_major == 0 || list_major == 1, ABSB_ERR_INVALID, "scan order %d", list_major);
  h->ix.scan_order = list_major;
  ABSB_API_END
}

int absb_ivf_last_stats(absb_ivf_t h, int64_t* vectors_scanned, int64_t* bytes_scanned,
                        int64_t* work_items, int64_t* launches) {
  ABSB_API_BEGIN
  NEED(h);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  ix.fold_stats();
  if (vectors_scanned) *vectors_scanned = ix.stats.vectors;
  if (bytes_scanned) *bytes_scanned = ix.stats.bytes;
  if (work_items) *work_items = ix.stats.items;
  if (launches) *launches = ix.stats.launches;
  ABSB_API_END
}

int absb_ivf_centroid_sums_dev(absb_ivf_t h, int64_t n, const float* x_dev, const int64_t* list_ids_dev,
                               float* sums_dev, float* counts_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(h); NEED(sums_dev); NEED(counts_dev);
  ABSB_CHECK(n >= 0 && (n == 0 || (x_dev && list_ids_dev)), ABSB_ERR_INVALID, "bad centroid_sums arguments");
  DeviceGuard g(h->ix.device);
  h->ix.centroid_sums_dev(n, x_dev, reinterpret_cast<const long long*>(list_ids_dev), sums_dev, counts_dev,
                          (cudaStream_t)stream);
  ABSB_API_END
}

int absb_rand_perm(int64_t n, int64_t seed, int32_t* out) {
  ABSB_API_BEGIN
  ABSB_CHECK(n >= 0 && n < ((int64_t)1 << 31) && (n == 0 || out), ABSB_ERR_INVALID, "bad rand_perm arguments");
  rand_perm_export(n, seed, out);
  ABSB_API_END
}

int absb_kmeans_split_clusters(int d, int64_t k, int64_t n, float* hassign, float* centroids, int64_t* nsplit) {
  ABSB_API_BEGIN
  ABSB_CHECK(d > 0 && k > 0 && n > k && hassign && centroids, ABSB_ERR_INVALID, "bad split_clusters arguments");
  const int64_t ns = split_clusters_export(d, k, n, hassign, centroids);
  if (nsplit) *nsplit = ns;
  ABSB_API_END
}

int absb_ivf_set_profile(absb_ivf_t h, int on) {
  ABSB_API_BEGIN
  NEED(h);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  if (ix.profile) ix.fold_profile();
  ix.profile = on != 0;
  if (on == 2) {
    ix.prof_ms[0] = ix.prof_ms[1] = ix.prof_ms[2] = ix.prof_ms[3] = 0;
    ix.prof_scan_launches = 0;
    ix.prof_scan16_launches = 0;
  }
  ABSB_API_END
}

int absb_ivf_get_profile(absb_ivf_t h, double* scan_ms, double* coarse_gemm_ms, double* other_ms,
                         int64_t* scan_launches) {
  ABSB_API_BEGIN
  NEED(h);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  ix.fold_profile();
  if (scan_ms) *scan_ms = ix.prof_ms[0];
  if (coarse_gemm_ms) *coarse_gemm_ms = ix.prof_ms[1];
  if (other_ms) *other_ms = ix.prof_ms[2];
  if (scan_launches) *scan_launches = ix.prof_scan_launches;
  ABSB_API_END
}

// Timeline of the spans recorded since the profile was switched on (before they are folded into the totals):
// out[i] = {kind, start_ms, stop_ms} relative to `base_event` (a cudaEvent_t recorded by the caller).
static int64_t dump_spans(const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& pool, size_t used,
                          const std::vector<int>& kinds, void* base_event, float* out, int64_t cap) {
  ABSB_CUDA(cudaDeviceSynchronize());
  int64_t n = 0;
  for (size_t i = 0; i < used && n < cap; ++i) {
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, (cudaEvent_t)base_event, pool[i].first) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&t1, (cudaEvent_t)base_event, pool[i].second) != cudaSuccess) continue;
    out[3 * n] = (float)kinds[i];
    out[3 * n + 1] = t0;
    out[3 * n + 2] = t1;
    ++n;
  }
  return n;
}

int absb_ivf_profile_spans(absb_ivf_t h, void* base_event, float* out, int64_t cap, int64_t* n) {
  ABSB_API_BEGIN
  NEED(h); NEED(base_event); NEED(out); NEED(n);
  DeviceGuard g(h->ix.device);
  *n = dump_spans(h->ix.ev_pool, h->ix.ev_used, h->ix.ev_kind, base_event, out, cap);
  ABSB_API_END
}

int absb_ivf_get_profile_scan16(absb_ivf_t h, double* scan16_ms, int64_t* scan16_launches) {
  ABSB_API_BEGIN
  NEED(h);
  IvfIndex& ix = h->ix;
  DeviceGuard g(ix.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  ix.fold_profile();
  if (scan16_ms) *scan16_ms = ix.prof_ms[3];
  if (scan16_launches) *scan16_launches = ix.prof_scan16_launches;
  ABSB_API_END
}

int absb_ivf_time_scan(absb_ivf_t h, int iters, void* stream, float* ms_mean) {
  ABSB_API_BEGIN
  NEED(h); NEED(ms_mean);
  IvfIndex& ix = h->ix;
  ABSB_CHECK(ix.have_last_scan, ABSB_ERR_STATE, "no search has run on this handle yet");
  ABSB_CHECK(iters >= 1, ABSB_ERR_INVALID, "iters=%d", iters);
  DeviceGuard g(ix.device);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  ABSB_CUDA(cudaEventCreate(&e0));
  ABSB_CUDA(cudaEventCreate(&e1));
  float total = 0.f;
  for (int i = 0; i < iters; ++i) {
    ABSB_CUDA(cudaMemsetAsync(ix.last_scan.queue_counter, 0, sizeof(int), st));
    ABSB_CUDA(cudaEventRecord(e0, st));
    ix.run_scan(ix.last_scan, st);
    ABSB_CUDA(cudaEventRecord(e1, st));
    ABSB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ABSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_mean = total / iters;
  ABSB_API_END
}

}  // extern "C"
