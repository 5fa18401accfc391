// ivf.cu — IndexIVFFlat / IndexFlatIP host logic on top of the kernels.
//
// Mirrors what faiss does behind Index.train / Index.add / Index.search for
// index_factory(d, "IVF<nlist>,Flat", METRIC_INNER_PRODUCT) — the calls made by
// `sidecar-search index train|fill|tune` (/root/reference/Makefile:38-39, 24-25, 31-32) and app.py
// (/root/reference/README.md:16,28).  See SURVEY.md §8(a) a4–a8 for the semantics followed here.
#include "ivf.cuh"

#include <cuda_fp16.h>

#include <cub/cub.cuh>

#include "gemm_tc.cuh"
#include <numeric>
#include <random>

namespace absb {

// =========================================================================================
// small kernels
// =========================================================================================
namespace {

__global__ void make_keys_kernel(int64_t n, const long long* __restrict__ list_ids, int nlist,
                                 int rank, int world, unsigned* __restrict__ keys,
                                 int* __restrict__ vals, unsigned* __restrict__ hist /* [nlist+1] */) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const long long l = list_ids[i];
    unsigned key = (unsigned)nlist;  // dropped rows sort last
    if (l >= 0 && l < nlist && (world == 1 || (int)(l % world) == rank)) key = (unsigned)l;
    keys[i] = key;
    vals[i] = (int)i;
    atomicAdd(&hist[key], 1u);
  }
}

// per-list bookkeeping for one add(): new sizes, number of fresh pages
__global__ void list_growth_kernel(int nlist, int P, const long long* __restrict__ old_size,
                                   const unsigned* __restrict__ hist, long long* __restrict__ new_size,
                                   int* __restrict__ fresh_pages, long long* __restrict__ total_pages_per_list) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nlist) return;
  const long long os = old_size[l], ns = os + hist[l];
  new_size[l] = ns;
  const long long op = (os + P - 1) / P, np = (ns + P - 1) / P;
  fresh_pages[l] = (int)(np - op);
  total_pages_per_list[l] = np;
}

__global__ void build_page_table_kernel(int nlist, int P, const long long* __restrict__ old_size,
                                        const long long* __restrict__ old_off,
                                        const int* __restrict__ old_pages,
                                        const long long* __restrict__ new_off,
                                        const int* __restrict__ fresh_off /* exclusive scan */,
                                        const int* __restrict__ fresh_pages, long long first_fresh_page,
                                        int* __restrict__ new_pages) {
  const int l = blockIdx.x;
  const long long op = (old_size[l] + P - 1) / P;
  const long long nb = new_off[l];
  for (long long j = threadIdx.x; j < op; j += blockDim.x) new_pages[nb + j] = old_pages[old_off[l] + j];
  const int f = fresh_pages[l];
  for (int j = threadIdx.x; j < f; j += blockDim.x)
    new_pages[nb + op + j] = (int)(first_fresh_page + fresh_off[l] + j);
}

// One warp per kept row (in list-sorted, insertion-stable order): copy the vector + id to its slot.
__global__ void scatter_rows_kernel(int64_t n_kept, int d4, int P, int slab_shift,
                                    const unsigned* __restrict__ keys_sorted,
                                    const int* __restrict__ perm, const unsigned* __restrict__ call_off,
                                    const long long* __restrict__ old_size,
                                    const long long* __restrict__ pt_off, const int* __restrict__ pt_pages,
                                    float* const* __restrict__ code_slabs,
                                    long long* const* __restrict__ id_slabs,
                                    const float* __restrict__ x, const long long* __restrict__ ids,
                                    long long id0, unsigned short* const* __restrict__ half_slabs,
                                    float* __restrict__ maxima) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float worst_r2 = 0.f, worst_x2 = 0.f;  // lane 0: largest |x - fp16(x)|^2 and |x|^2 of this warp's rows
  for (int64_t p = wid; p < n_kept; p += nw) {
    const unsigned l = keys_sorted[p];
    const int row = perm[p];
    const long long pos = old_size[l] + (long long)(p - call_off[l]);
    const long long page = pt_pages[pt_off[l] + pos / P];
    const int slot = (int)(pos % P);
    const int slab = (int)(page >> slab_shift);
    const long long in_slab = page & ((1ll << slab_shift) - 1);
    float4* dst = reinterpret_cast<float4*>(code_slabs[slab]) + ((size_t)in_slab * P + slot) * d4;
    const float4* src = reinterpret_cast<const float4*>(x) + (size_t)row * d4;
    if (half_slabs == nullptr) {
      for (int j = lane; j < d4; j += 32) dst[j] = __ldcs(src + j);
    } else {
      // fp16 shadow copy for the two-stage scan + the two norms its error bound needs
      uint2* hdst = reinterpret_cast<uint2*>(half_slabs[slab]) + ((size_t)in_slab * P + slot) * d4;
      float r2 = 0.f, x2 = 0.f;
      for (int j = lane; j < d4; j += 32) {
        const float4 v = __ldcs(src + j);
        dst[j] = v;
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const unsigned*>(&h0);
        pk.y = *reinterpret_cast<const unsigned*>(&h1);
        hdst[j] = pk;
        const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
        const float e0 = v.x - b0.x, e1 = v.y - b0.y, e2 = v.z - b1.x, e3 = v.w - b1.y;
        r2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        x2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        x2 += __shfl_xor_sync(0xffffffffu, x2, o);
      }
      // NaN/inf rows (fp16 overflow) make the bound infinite: such an index always takes the fallback
      worst_r2 = (r2 > worst_r2 || !(r2 == r2)) ? (r2 == r2 ? r2 : INFINITY) : worst_r2;
      worst_x2 = (x2 > worst_x2 || !(x2 == x2)) ? (x2 == x2 ? x2 : INFINITY) : worst_x2;
    }
    if (lane == 0) id_slabs[slab][(size_t)in_slab * P + slot] = ids ? ids[row] : id0 + row;
  }
  if (half_slabs != nullptr && lane == 0) {
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<unsigned*>(maxima), __float_as_uint(sqrtf(worst_r2)));
    atomicMax(reinterpret_cast<unsigned*>(maxima) + 1, __float_as_uint(sqrtf(worst_x2)));
  }
}

__global__ void gather_rows_kernel(int64_t n, int d4, const int* __restrict__ rows,
                                   const float* __restrict__ x, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = wid; p < n; p += nw) {
    const float4* src = reinterpret_cast<const float4*>(x) + (size_t)rows[p] * d4;
    float4* dst = reinterpret_cast<float4*>(out) + (size_t)p * d4;
    for (int j = lane; j < d4; j += 32) dst[j] = src[j];
  }
}

// Clustering::compute_centroids: members of centroid c are rows perm[off[c] .. off[c+1]) in
// ascending row order; fp32 running sum in that order, then multiply by 1/count.
__global__ void centroid_mean_kernel(int d, const unsigned* __restrict__ off,
                                     const int* __restrict__ perm, const float* __restrict__ x,
                                     float* __restrict__ centroids, float* __restrict__ hassign,
                                     int normalise = 1) {
  const int c = blockIdx.x;
  const unsigned b = off[c], e = off[c + 1];
  if (threadIdx.x == 0) hassign[c] = (float)(e - b);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float acc = 0.f;
    for (unsigned i = b; i < e; ++i) acc += x[(size_t)perm[i] * d + j];
    if (normalise && e > b) acc *= 1.f / (float)(e - b);
    centroids[(size_t)c * d + j] = acc;
  }
}

__global__ void copy_list_kernel(ListTable lt, long long l, float* __restrict__ codes,
                                 long long* __restrict__ ids) {
  const long long size = lt.list_size[l];
  const int P = lt.page_vecs;
  const int d = lt.d;
  for (long long v = blockIdx.x; v < size; v += gridDim.x) {
    const long long page = lt.pt_pages[lt.pt_off[l] + v / P];
    const int slot = (int)(v % P);
    const int slab = (int)(page >> lt.slab_shift);
    const long long in_slab = page & ((1ll << lt.slab_shift) - 1);
    const float* src = lt.code_slabs[slab] + ((size_t)in_slab * P + slot) * d;
    if (codes)
      for (int j = threadIdx.x; j < d; j += blockDim.x) codes[(size_t)v * d + j] = src[j];
    if (ids && threadIdx.x == 0) ids[v] = lt.id_slabs[slab][(size_t)in_slab * P + slot];
  }
}

// One CTA per page copy; `from`/`to` >= 0 name pool pages, < 0 scratch slots (-1 - v).
__global__ void move_pages_kernel(int64_t n_moves, const PageMove* __restrict__ moves, int P, int d4,
                                  int slab_shift, float* const* __restrict__ code_slabs,
                                  long long* const* __restrict__ id_slabs, float* __restrict__ scratch_codes,
                                  long long* __restrict__ scratch_ids, unsigned short* const* __restrict__ half_slabs,
                                  unsigned short* __restrict__ scratch_half) {
  const long long mask = (1ll << slab_shift) - 1;
  const size_t page_f4 = (size_t)P * d4;
  for (int64_t m = blockIdx.x; m < n_moves; m += gridDim.x) {
    const PageMove mv = moves[m];
    const float4* src_c;
    const long long* src_i;
    float4* dst_c;
    long long* dst_i;
    if (mv.from >= 0) {
      src_c = reinterpret_cast<const float4*>(code_slabs[mv.from >> slab_shift]) + (size_t)(mv.from & mask) * page_f4;
      src_i = id_slabs[mv.from >> slab_shift] + (size_t)(mv.from & mask) * P;
    } else {
      src_c = reinterpret_cast<const float4*>(scratch_codes) + (size_t)(-1 - mv.from) * page_f4;
      src_i = scratch_ids + (size_t)(-1 - mv.from) * P;
    }
    if (mv.to >= 0) {
      dst_c = reinterpret_cast<float4*>(code_slabs[mv.to >> slab_shift]) + (size_t)(mv.to & mask) * page_f4;
      dst_i = id_slabs[mv.to >> slab_shift] + (size_t)(mv.to & mask) * P;
    } else {
      dst_c = reinterpret_cast<float4*>(scratch_codes) + (size_t)(-1 - mv.to) * page_f4;
      dst_i = scratch_ids + (size_t)(-1 - mv.to) * P;
    }
    for (size_t j = threadIdx.x; j < page_f4; j += blockDim.x) __stcs(dst_c + j, __ldcs(src_c + j));
    for (int j = threadIdx.x; j < P; j += blockDim.x) dst_i[j] = src_i[j];
    if (half_slabs != nullptr) {
      const size_t page_h4 = page_f4 / 2;  // 16-byte words of one page of fp16 codes
      const float4* src_h = reinterpret_cast<const float4*>(
          mv.from >= 0 ? half_slabs[mv.from >> slab_shift] + (size_t)(mv.from & mask) * page_h4 * 8
                       : scratch_half + (size_t)(-1 - mv.from) * page_h4 * 8);
      float4* dst_h = reinterpret_cast<float4*>(
          mv.to >= 0 ? half_slabs[mv.to >> slab_shift] + (size_t)(mv.to & mask) * page_h4 * 8
                     : scratch_half + (size_t)(-1 - mv.to) * page_h4 * 8);
      for (size_t j = threadIdx.x; j < page_h4; j += blockDim.x) __stcs(dst_h + j, __ldcs(src_h + j));
    }
  }
}

int grid_for(int64_t work, int threads, int sms) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), (int64_t)sms * 16));
}

}  // namespace

// =========================================================================================
// PagePool
// =========================================================================================
void PagePool::configure(int d_, int page_vecs_) {
  d = d_;
  page_vecs = page_vecs_;
  // slabs of ~256 MB of codes
  const size_t page_bytes = (size_t)page_vecs * d * sizeof(float);
  slab_shift = 0;
  while (((size_t)1 << (slab_shift + 1)) * page_bytes <= ((size_t)256 << 20)) ++slab_shift;
}

void PagePool::ensure_pages(int64_t total_pages, cudaStream_t st) {
  const int64_t per_slab = (int64_t)1 << slab_shift;
  const size_t need = (size_t)ceil_div(total_pages, per_slab);
  if (need <= code_slabs.size()) return;
  while (code_slabs.size() < need) {
    float* c = nullptr;
    long long* i = nullptr;
    ABSB_CUDA(cudaMalloc(&c, (size_t)per_slab * page_vecs * d * sizeof(float)));
    cudaError_t e = cudaMalloc(&i, (size_t)per_slab * page_vecs * sizeof(long long));
    if (e != cudaSuccess) {
      cudaFree(c);
      ABSB_CUDA(e);
    }
    unsigned short* hs = nullptr;
    if (shadow) {
      e = cudaMalloc(&hs, (size_t)per_slab * page_vecs * d * sizeof(unsigned short));
      if (e != cudaSuccess) {
        cudaFree(c);
        cudaFree(i);
        ABSB_CUDA(e);
      }
    }
    code_slabs.push_back(c);
    id_slabs.push_back(i);
    half_slabs.push_back(hs);
  }
  if (code_slabs.size() > table_cap) {
    // the old tables may still be referenced by in-flight kernels on st: sync before freeing
    ABSB_CUDA(cudaStreamSynchronize(st));
    table_cap = std::max<size_t>(64, code_slabs.size() * 2);
    d_code_slabs.alloc_exact(table_cap);
    d_id_slabs.alloc_exact(table_cap);
    d_half_slabs.alloc_exact(table_cap);
  }
  ABSB_CUDA(cudaMemcpyAsync(d_code_slabs.p, code_slabs.data(), code_slabs.size() * sizeof(float*),
                            cudaMemcpyHostToDevice, st));
  ABSB_CUDA(cudaMemcpyAsync(d_id_slabs.p, id_slabs.data(), id_slabs.size() * sizeof(long long*),
                            cudaMemcpyHostToDevice, st));
  ABSB_CUDA(cudaMemcpyAsync(d_half_slabs.p, half_slabs.data(), half_slabs.size() * sizeof(unsigned short*),
                            cudaMemcpyHostToDevice, st));
  ABSB_CUDA(cudaStreamSynchronize(st));  // host vectors may reallocate later
}

void PagePool::release() {
  for (auto p : code_slabs) cudaFree(p);
  for (auto p : id_slabs) cudaFree(p);
  for (auto p : half_slabs)
    if (p) cudaFree(p);
  code_slabs.clear();
  id_slabs.clear();
  half_slabs.clear();
  pages_used = 0;
}

// =========================================================================================
// scores GEMM dispatch
// =========================================================================================
// impl 0: exact-fp32 FFMA GEMM.  impl 1: split-bf16 tcgen05 GEMM (hi/mid/lo terms, fp32 accumulate in
// TMEM) — fp32-faithful scores at tensor-core speed; needs d % 64 == 0 and nlist % 32 == 0, other
// shapes take the FFMA kernel.
void IvfIndex::coarse_scores(int M, const float* q, float* S, cudaStream_t st) {
  Span sp(this, st, 1);
  if (coarse_impl == 1 && d % 64 == 0 && nlist % 32 == 0) {
    if (c3_dirty) {
      centroids3.reserve((size_t)nlist * 3 * d);
      split3_bf16(nlist, d, centroids.p, centroids3.p, st);
      c3_dirty = false;
    }
    ws_q3.reserve((size_t)M * 3 * d);
    split3_bf16(M, d, q, ws_q3.p, st);
    gemm_split3_f32(M, nlist, d, ws_q3.p, centroids3.p, S, nlist, props.sm_count, st);
    stats.launches += 2;
  } else {
    gemm_nt_f32(M, nlist, d, q, d, centroids.p, d, S, nlist, st);
    stats.launches += 1;
  }
}

// =========================================================================================
// IvfIndex
// =========================================================================================
IvfIndex::IvfIndex(int d_, int nlist_, int device_) : d(d_), nlist(nlist_), device(device_) {
  ABSB_CHECK(d > 0 && d % 4 == 0, ABSB_ERR_UNSUPPORTED, "d must be a positive multiple of 4 (d=%d)", d);
  ABSB_CHECK(nlist > 0 && nlist < (1 << 30), ABSB_ERR_INVALID, "nlist=%d", nlist);
  props = device_props(device);
  DeviceGuard g(device);
  ABSB_CUDA(cudaStreamCreate(&own_stream));
  centroids.alloc_exact((size_t)nlist * d);
  list_size.alloc_exact(nlist);
  pt_off.alloc_exact(nlist + 1);
  ws_counters.alloc_exact(2);
  ws_stats.alloc_exact(2);
  ws_counters2.alloc_exact(2);
  ws_stats2.alloc_exact(2);
  ws_maxima.alloc_exact(2);
  ws_nflag.alloc_exact(1);
  ABSB_CUDA(cudaMemsetAsync(ws_maxima.p, 0, 2 * sizeof(float), own_stream));
  ABSB_CUDA(cudaMemsetAsync(ws_nflag.p, 0, sizeof(int), own_stream));
  ABSB_CUDA(cudaMemsetAsync(list_size.p, 0, sizeof(long long) * nlist, own_stream));
  ABSB_CUDA(cudaMemsetAsync(pt_off.p, 0, sizeof(long long) * (nlist + 1), own_stream));
  ABSB_CUDA(cudaStreamSynchronize(own_stream));
  h_list_size.assign(nlist, 0);
  // ~64 KB pages
  int P = (int)std::max<size_t>(1, (64 * 1024) / ((size_t)d * sizeof(float)));
  int p2 = 1;
  while (p2 * 2 <= P) p2 *= 2;
  pool.configure(d, p2);
}

IvfIndex::~IvfIndex() {
  cudaSetDevice(device);
  if (own_stream) {
    cudaStreamSynchronize(own_stream);
    cudaStreamDestroy(own_stream);
  }
  for (auto& e : ev_pool) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
}

IvfIndex::Span::Span(IvfIndex* ix_, cudaStream_t st_, int kind) : ix(ix_), st(st_), on(ix_->profile) {
  if (!on) return;
  if (ix->ev_used == ix->ev_pool.size()) {
    cudaEvent_t a, b;
    ABSB_CUDA(cudaEventCreate(&a));
    ABSB_CUDA(cudaEventCreate(&b));
    ix->ev_pool.emplace_back(a, b);
  }
  auto& pr = ix->ev_pool[ix->ev_used++];
  ix->ev_kind.push_back(kind);
  stop = pr.second;
  cudaEventRecord(pr.first, st);
}

// call with the stream idle
void IvfIndex::fold_profile() {
  for (size_t i = 0; i < ev_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_pool[i].first, ev_pool[i].second) == cudaSuccess) {
      prof_ms[ev_kind[i]] += ms;
      if (ev_kind[i] == 3) prof_ms[0] += ms;  // the fp16 pass is part of the scan total
    }
    if (ev_kind[i] == 0 || ev_kind[i] == 3) ++prof_scan_launches;
    if (ev_kind[i] == 3) ++prof_scan16_launches;
  }
  ev_used = 0;
  ev_kind.clear();
}

ListTable IvfIndex::table() const {
  ListTable lt;
  lt.nlist = nlist;
  lt.d = d;
  lt.page_vecs = pool.page_vecs;
  lt.slab_shift = pool.slab_shift;
  lt.list_size = list_size.p;
  lt.pt_off = pt_off.p;
  lt.pt_pages = pt_pages.p;
  lt.code_slabs = pool.d_code_slabs.p;
  lt.id_slabs = pool.d_id_slabs.p;
  lt.half_slabs = pool.shadow ? pool.d_half_slabs.p : nullptr;
  return lt;
}

void IvfIndex::reset() {
  DeviceGuard g(device);
  ABSB_CUDA(cudaStreamSynchronize(own_stream));
  pool.release();
  pt_pages.release();
  pt_total_pages = 0;
  ABSB_CUDA(cudaMemset(list_size.p, 0, sizeof(long long) * nlist));
  ABSB_CUDA(cudaMemset(pt_off.p, 0, sizeof(long long) * (nlist + 1)));
  ABSB_CUDA(cudaMemset(ws_maxima.p, 0, 2 * sizeof(float)));
  h_list_size.assign(nlist, 0);
  h_items_prefix_desc.clear();
  items_bound_chunk = -1;
  ntotal = 0;
  rows_seen = 0;
  have_last_scan = false;
}

void IvfIndex::set_centroids_dev(const float* c, cudaStream_t st) {
  ABSB_CUDA(cudaMemcpyAsync(centroids.p, c, sizeof(float) * (size_t)nlist * d, cudaMemcpyDeviceToDevice, st));
  trained = true;
  c3_dirty = true;
}

// ---------------------------------------------------------------- coarse / assign ---------
void IvfIndex::coarse_dev(int64_t nq, const float* q, int nprobe, float* Dc, long long* Ic,
                          bool finalize, cudaStream_t st) {
  ABSB_CHECK(trained, ABSB_ERR_STATE, "index is not trained");
  ABSB_CHECK(nprobe >= 1 && nprobe <= ABSB_MAX_K, ABSB_ERR_INVALID, "nprobe=%d outside [1,%d]", nprobe, ABSB_MAX_K);
  // chunk rows so the score matrix stays <= 1 GiB
  const int64_t rows_max = std::max<int64_t>(128, ((int64_t)1 << 28) / nlist);
  for (int64_t r0 = 0; r0 < nq; r0 += rows_max) {
    const int64_t nr = std::min(rows_max, nq - r0);
    ws_scores.reserve((size_t)std::min(rows_max, nq) * nlist);
    coarse_scores((int)nr, q + r0 * d, ws_scores.p, st);
    {
      Span sp(this, st, 2);
      select_rows(ws_scores.p, nlist, nr, nlist, 0, nprobe, Dc + r0 * nprobe, Ic + r0 * nprobe, nprobe,
                  finalize, st);
    }
    stats.launches += 1;
  }
}

void IvfIndex::assign_dev(int64_t n, const float* x, long long* list_ids, cudaStream_t st) {
  ABSB_CHECK(trained, ABSB_ERR_STATE, "index is not trained");
  if (coarse_impl == 1 && d % 64 == 0 && nlist % 32 == 0) {
    // tensor-core path: split-bf16 GEMM with the arg-max fused into the epilogue — the [n, nlist]
    // score matrix (262 KB per row at nlist = 65536) is never written
    if (c3_dirty) {
      centroids3.reserve((size_t)nlist * 3 * d);
      split3_bf16(nlist, d, centroids.p, centroids3.p, st);
      c3_dirty = false;
      stats.launches += 1;
    }
    const int P = argmax_partials_per_row(nlist);
    const int64_t slice = std::max<int64_t>(256, std::min<int64_t>(65536, ((int64_t)1 << 27) / P));
    const int64_t cap = std::min(slice, n);
    ws_q3.reserve((size_t)cap * 3 * d);
    ws_amax.reserve((size_t)cap * P);
    ws_aidx.reserve((size_t)cap * P);
    for (int64_t r0 = 0; r0 < n; r0 += slice) {
      const int64_t nr = std::min(slice, n - r0);
      Span sp(this, st, 1);
      split3_bf16(nr, d, x + r0 * d, ws_q3.p, st);
      gemm_split3_argmax((int)nr, nlist, d, ws_q3.p, centroids3.p, ws_amax.p, ws_aidx.p, list_ids + r0, nullptr,
                         props.sm_count, st);
      stats.launches += 3;
    }
    return;
  }
  ws_coarse_s.reserve((size_t)std::min<int64_t>(n, (int64_t)1 << 22));
  // Dc is scratch: process in slices that fit ws_coarse_s
  const int64_t slice = (int64_t)1 << 22;
  for (int64_t r0 = 0; r0 < n; r0 += slice) {
    const int64_t nr = std::min(slice, n - r0);
    coarse_dev(nr, x + r0 * d, 1, ws_coarse_s.p, list_ids + r0, true, st);
  }
}

// ---------------------------------------------------------------- add ----------------------
void IvfIndex::refresh_host_sizes(cudaStream_t st) {
  ABSB_CUDA(cudaMemcpyAsync(h_list_size.data(), list_size.p, sizeof(long long) * nlist,
                            cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  items_bound_chunk = -1;
}

// Upper bound of the work items one query can produce: the sum of the `nprobe` largest per-list item
// counts, counted by the plan's own list walk (contiguous page runs cut at scan_chunk vectors), so a
// compacted 3,000-vector list costs ~25 entries instead of one per 16-vector page.
int64_t IvfIndex::items_bound_per_query(int nprobe, cudaStream_t st) {
  if (ntotal == 0) return 0;
  if (items_bound_chunk != scan_chunk) {
    ws_list_items.reserve((size_t)nlist);
    launch_count_list_items(table(), scan_chunk, ws_list_items.p, st);
    std::vector<int> cnt((size_t)nlist);
    ABSB_CUDA(cudaMemcpyAsync(cnt.data(), ws_list_items.p, sizeof(int) * nlist, cudaMemcpyDeviceToHost, st));
    ABSB_CUDA(cudaStreamSynchronize(st));
    std::sort(cnt.begin(), cnt.end(), std::greater<int>());
    h_items_prefix_desc.assign((size_t)nlist + 1, 0);
    for (int l = 0; l < nlist; ++l) h_items_prefix_desc[l + 1] = h_items_prefix_desc[l] + cnt[l];
    items_bound_chunk = scan_chunk;
  }
  return h_items_prefix_desc[std::min(nprobe, nlist)];
}

void IvfIndex::add_core_dev(int64_t n, const float* x, const long long* ids,
                            const long long* list_ids, cudaStream_t st) {
  ABSB_CHECK(trained, ABSB_ERR_STATE, "index is not trained");
  if (n == 0) return;
  ABSB_CHECK(n < ((int64_t)1 << 31), ABSB_ERR_INVALID, "add chunk too large");
  const int P = pool.page_vecs;
  const int sms = props.sm_count;

  // scratch carved out of one buffer: keys, keys_sorted, vals, perm, hist, call_off, fresh, ...
  DBuf<unsigned> keys, keys_sorted, hist, call_off;
  DBuf<int> vals, perm, fresh_pages, fresh_off;
  DBuf<long long> new_size, pages_per_list, new_off;
  keys.alloc_exact(n); keys_sorted.alloc_exact(n); vals.alloc_exact(n); perm.alloc_exact(n);
  hist.alloc_exact(nlist + 2); call_off.alloc_exact(nlist + 2);
  fresh_pages.alloc_exact(nlist + 1); fresh_off.alloc_exact(nlist + 1);
  new_size.alloc_exact(nlist); pages_per_list.alloc_exact(nlist + 1); new_off.alloc_exact(nlist + 1);

  ABSB_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned) * (nlist + 2), st));
  make_keys_kernel<<<grid_for(n, 256, sms), 256, 0, st>>>(n, list_ids, nlist, shard_rank, shard_world,
                                                          keys.p, vals.p, hist.p);
  ABSB_CUDA(cudaGetLastError());

  int bits = 1;
  while ((1ll << bits) < (long long)nlist + 1) ++bits;
  size_t tmp_bytes = 0, t2 = 0;
  ABSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, perm.p,
                                            (int)n, 0, bits, st));
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, hist.p, call_off.p, nlist + 2, st));
  tmp_bytes = std::max(tmp_bytes, t2);
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, pages_per_list.p, new_off.p, nlist + 1, st));
  tmp_bytes = std::max(tmp_bytes, t2);
  ws_cub.reserve(tmp_bytes + 16);
  size_t tb = ws_cub.cap;
  ABSB_CUDA(cub::DeviceRadixSort::SortPairs(ws_cub.p, tb, keys.p, keys_sorted.p, vals.p, perm.p, (int)n,
                                            0, bits, st));
  tb = ws_cub.cap;
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(ws_cub.p, tb, hist.p, call_off.p, nlist + 2, st));

  list_growth_kernel<<<(nlist + 255) / 256, 256, 0, st>>>(nlist, P, list_size.p, hist.p, new_size.p,
                                                          fresh_pages.p, pages_per_list.p);
  ABSB_CUDA(cudaGetLastError());
  ABSB_CUDA(cudaMemsetAsync(fresh_pages.p + nlist, 0, sizeof(int), st));
  ABSB_CUDA(cudaMemsetAsync(pages_per_list.p + nlist, 0, sizeof(long long), st));
  tb = ws_cub.cap;
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(ws_cub.p, tb, fresh_pages.p, fresh_off.p, nlist + 1, st));
  tb = ws_cub.cap;
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(ws_cub.p, tb, pages_per_list.p, new_off.p, nlist + 1, st));

  // host needs: rows kept, fresh pages, total pages
  unsigned h_dropped_off = 0;
  int h_fresh_total = 0;
  long long h_total_pages = 0;
  ABSB_CUDA(cudaMemcpyAsync(&h_dropped_off, call_off.p + nlist, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaMemcpyAsync(&h_fresh_total, fresh_off.p + nlist, sizeof(int), cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaMemcpyAsync(&h_total_pages, new_off.p + nlist, sizeof(long long), cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  const int64_t n_kept = h_dropped_off;  // rows with key < nlist come first
  ABSB_CHECK(h_total_pages < ((long long)1 << 31), ABSB_ERR_UNSUPPORTED, "page table overflow");

  pool.ensure_pages(pool.pages_used + h_fresh_total, st);
  DBuf<int> new_pages;
  new_pages.alloc_exact((size_t)std::max<long long>(h_total_pages, 1));
  build_page_table_kernel<<<nlist, 64, 0, st>>>(nlist, P, list_size.p, pt_off.p, pt_pages.p, new_off.p,
                                                fresh_off.p, fresh_pages.p, pool.pages_used, new_pages.p);
  ABSB_CUDA(cudaGetLastError());
  if (n_kept > 0) {
    scatter_rows_kernel<<<grid_for(n_kept * 32, 256, sms), 256, 0, st>>>(
        n_kept, d / 4, P, pool.slab_shift, keys_sorted.p, perm.p, call_off.p, list_size.p, new_off.p,
        new_pages.p, pool.d_code_slabs.p, pool.d_id_slabs.p, x, ids, rows_seen,
        pool.shadow ? pool.d_half_slabs.p : nullptr, ws_maxima.p);
    ABSB_CUDA(cudaGetLastError());
  }
  ABSB_CUDA(cudaMemcpyAsync(list_size.p, new_size.p, sizeof(long long) * nlist, cudaMemcpyDeviceToDevice, st));
  ABSB_CUDA(cudaMemcpyAsync(pt_off.p, new_off.p, sizeof(long long) * (nlist + 1), cudaMemcpyDeviceToDevice, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  pt_pages = std::move(new_pages);
  pt_total_pages = h_total_pages;
  pool.pages_used += h_fresh_total;
  // faiss: ids default to ntotal + i over ALL rows of the call; a shard numbers rows the same way
  // (rows_seen) but counts only the rows it keeps in ntotal.
  ntotal += n_kept;
  rows_seen += n;
  refresh_host_sizes(st);
  have_last_scan = false;
}

void IvfIndex::add_dev(int64_t n, const float* x, const long long* ids, cudaStream_t st) {
  if (n == 0) return;
  const int64_t step = (int64_t)1 << 22;
  for (int64_t r0 = 0; r0 < n; r0 += step) {
    const int64_t nr = std::min(step, n - r0);
    ws_list_ids.reserve((size_t)nr);
    assign_dev(nr, x + r0 * d, ws_list_ids.p, st);
    add_core_dev(nr, x + r0 * d, ids ? ids + r0 : nullptr, ws_list_ids.p, st);
  }
}

void IvfIndex::get_list(int64_t l, float* codes, long long* ids) {
  const int64_t size = h_list_size[l];
  if (size == 0) return;
  cudaStream_t st = own_stream;
  DBuf<float> dc;
  DBuf<long long> di;
  if (codes) dc.alloc_exact((size_t)size * d);
  if (ids) di.alloc_exact((size_t)size);
  copy_list_kernel<<<(unsigned)std::min<int64_t>(size, 4096), 128, 0, st>>>(table(), l, dc.p, di.p);
  ABSB_CUDA(cudaGetLastError());
  if (codes) ABSB_CUDA(cudaMemcpyAsync(codes, dc.p, sizeof(float) * (size_t)size * d, cudaMemcpyDeviceToHost, st));
  if (ids) ABSB_CUDA(cudaMemcpyAsync(ids, di.p, sizeof(long long) * (size_t)size, cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
}

// ---------------------------------------------------------------- two-stage scan -----------
void IvfIndex::set_two_stage(int shortlist) {
  ABSB_CHECK(shortlist == 0 || shortlist == 32 || shortlist == 64 || shortlist == 128, ABSB_ERR_INVALID,
             "shortlist length %d (0 = off, 32, 64 or 128)", shortlist);
  if (shortlist > 0 && !pool.shadow) {
    ABSB_CHECK(d == 1024, ABSB_ERR_UNSUPPORTED, "the two-stage scan is built for d = 1024 (d=%d)", d);
    ABSB_CHECK(ntotal == 0 && pool.pages_used == 0, ABSB_ERR_STATE,
               "enable the two-stage scan before the first add(): the fp16 shadow codes are written by add()");
    pool.shadow = true;
  }
  two_stage_k = shortlist;  // 0 keeps the shadow codes (if any) and returns to the single-pass scan
}

int64_t IvfIndex::two_stage_fallbacks(cudaStream_t st) {
  int h = 0;
  ABSB_CUDA(cudaMemcpyAsync(&h, ws_nflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaMemsetAsync(ws_nflag.p, 0, sizeof(int), st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  return h;
}

// ---------------------------------------------------------------- compact ------------------
// Target layout: list l owns pool pages [pt_off[l], pt_off[l+1]) in order, i.e. pt_pages becomes the
// identity and the plan kernel can merge a whole list into scan_chunk-sized work items.  The pages
// are permuted IN PLACE, window by window, because a 106 GB shard leaves no room for a second copy:
// for the window W = [w, w+S) of target positions
//   phase 1  scratch[i] <- page src[w+i]                  (the S pages that belong in W)
//   phase 2  pages of W still wanted by a later target move into the slots phase 1 just freed
//            outside W (|W minus sources| == |sources minus W|, paired up in order)
//   phase 3  page w+i <- scratch[i]
// After the round every target >= w+S again has its source >= w+S.  Pages already in place are
// skipped.  Traffic: every page is read and written at most twice (+ once if it is displaced).
void plan_page_compaction(std::vector<int> src, int64_t S, std::vector<PageMove>& moves,
                          std::vector<int64_t>& phase_end) {
  const int64_t n = (int64_t)src.size();
  moves.clear();
  phase_end.clear();
  if (n == 0) return;
  ABSB_CHECK(S >= 1, ABSB_ERR_INVALID, "compaction needs at least one scratch page");
  std::vector<int> where((size_t)n, -1);  // where[p] = the target that wants pool page p
  for (int64_t t = 0; t < n; ++t) {
    ABSB_CHECK(src[t] >= 0 && src[t] < n && where[src[t]] < 0, ABSB_ERR_STATE, "page table is not a permutation");
    where[src[t]] = (int)t;
  }
  std::vector<int> displaced, freed;
  for (int64_t w = 0; w < n; w += S) {
    const int64_t e = std::min(n, w + S);
    const size_t before = moves.size();
    for (int64_t t = w; t < e; ++t)
      if (src[t] != t) moves.push_back({src[t], (int)(-1 - (t - w))});
    if (moves.size() == before) continue;  // window already in place
    phase_end.push_back((int64_t)moves.size());
    displaced.clear();
    freed.clear();
    for (int64_t t = w; t < e; ++t) {
      if (where[t] < w || where[t] >= e) displaced.push_back((int)t);  // page t is wanted later
      if (src[t] < w || src[t] >= e) freed.push_back(src[t]);          // its slot was read in phase 1
    }
    ABSB_CHECK(displaced.size() == freed.size(), ABSB_ERR_STATE, "compaction bookkeeping broke");
    for (size_t k = 0; k < displaced.size(); ++k) {
      const int p = displaced[k], f = freed[k], u = where[p];
      moves.push_back({p, f});
      src[u] = f;
      where[f] = u;
    }
    if (!displaced.empty()) phase_end.push_back((int64_t)moves.size());
    for (int64_t t = w; t < e; ++t) {
      if (src[t] != t) moves.push_back({(int)(-1 - (t - w)), (int)t});
      src[t] = (int)t;
      where[t] = (int)t;
    }
    phase_end.push_back((int64_t)moves.size());
  }
}

void IvfIndex::compact(int64_t scratch_pages, cudaStream_t st) {
  const int64_t n = pt_total_pages;
  if (n == 0) return;
  ABSB_CHECK(n == pool.pages_used, ABSB_ERR_STATE, "page table (%lld) and pool (%lld) disagree", (long long)n,
             (long long)pool.pages_used);
  std::vector<int> src((size_t)n);
  ABSB_CUDA(cudaMemcpyAsync(src.data(), pt_pages.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  const int P = pool.page_vecs;
  const size_t page_bytes = (size_t)P * d * (pool.shadow ? 6 : 4) + (size_t)P * sizeof(long long);
  if (scratch_pages <= 0) {
    size_t free_b = 0, total_b = 0;
    ABSB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    scratch_pages = (int64_t)std::min<size_t>(free_b / 2 / page_bytes, (size_t)1 << 16);  // <= 4 GB of fp32 d=1024
  }
  scratch_pages = std::max<int64_t>(1, std::min(scratch_pages, n));
  std::vector<PageMove> moves;
  std::vector<int64_t> phase_end;
  plan_page_compaction(std::move(src), scratch_pages, moves, phase_end);
  if (!moves.empty()) {
    DBuf<float> sc_codes;
    DBuf<long long> sc_ids;
    DBuf<PageMove> d_moves;
    DBuf<unsigned short> sc_half;
    sc_codes.alloc_exact((size_t)scratch_pages * P * d);
    sc_ids.alloc_exact((size_t)scratch_pages * P);
    if (pool.shadow) sc_half.alloc_exact((size_t)scratch_pages * P * d);
    d_moves.alloc_exact(moves.size());
    ABSB_CUDA(cudaMemcpyAsync(d_moves.p, moves.data(), sizeof(PageMove) * moves.size(), cudaMemcpyHostToDevice, st));
    int64_t begin = 0;
    for (int64_t end : phase_end) {
      const int64_t cnt = end - begin;
      if (cnt > 0) {
        const int grid = (int)std::min<int64_t>(cnt, (int64_t)props.sm_count * 8);
        move_pages_kernel<<<grid, 256, 0, st>>>(cnt, d_moves.p + begin, P, d / 4, pool.slab_shift, pool.d_code_slabs.p,
                                                pool.d_id_slabs.p, sc_codes.p, sc_ids.p,
                                                pool.shadow ? pool.d_half_slabs.p : nullptr, sc_half.p);
        ABSB_CUDA(cudaGetLastError());
      }
      begin = end;
    }
    std::vector<int> iota((size_t)n);
    std::iota(iota.begin(), iota.end(), 0);
    ABSB_CUDA(cudaMemcpyAsync(pt_pages.p, iota.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    ABSB_CUDA(cudaStreamSynchronize(st));  // scratch and host vectors die with this frame
  }
  have_last_scan = false;
  items_bound_chunk = -1;
}

// ---------------------------------------------------------------- train --------------------
namespace {
// faiss rand_perm: Fisher–Yates driven by std::mt19937, i2 = i + mt() % (n - i)
std::vector<int> rand_perm_host(int64_t n, int64_t seed) {
  std::vector<int> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  std::mt19937 mt((unsigned)seed);
  for (int64_t i = 0; i + 1 < n; ++i) {
    const int64_t i2 = i + (int64_t)(mt() % (unsigned)(n - i));
    std::swap(perm[i], perm[i2]);
  }
  return perm;
}

// Clustering::split_clusters on the host (needs a sequential RNG); returns number of splits.
int64_t split_clusters_host(int d, int64_t k, int64_t n, std::vector<float>& hassign, float* centroids) {
  const float EPS = 1.f / 1024.f;
  int64_t nsplit = 0;
  std::mt19937 mt(1234u);
  for (int64_t ci = 0; ci < k; ++ci) {
    if (hassign[ci] != 0.f) continue;
    int64_t cj;
    for (cj = 0;; cj = (cj + 1) % k) {
      const float p = (float)((hassign[cj] - 1.0) / (float)(n - k));
      const float r = mt() / (float)mt.max();
      if (r < p) break;
    }
    std::copy(centroids + cj * d, centroids + (cj + 1) * d, centroids + ci * d);
    for (int j = 0; j < d; ++j) {
      if (j % 2 == 0) {
        centroids[ci * d + j] *= 1 + EPS;
        centroids[cj * d + j] *= 1 - EPS;
      } else {
        centroids[ci * d + j] *= 1 - EPS;
        centroids[cj * d + j] *= 1 + EPS;
      }
    }
    hassign[ci] = hassign[cj] / 2;
    hassign[cj] -= hassign[ci];
    ++nsplit;
  }
  return nsplit;
}
}  // namespace

namespace {
// faiss fvec_renorm_L2: x_i *= 1 / sqrt(|x_i|^2) for every row with a non-zero norm; one warp per row.
__global__ void renorm_rows_kernel(int64_t n, int d, float* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  float* xr = x + (size_t)r * d;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) ss = fmaf(xr[j], xr[j], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (ss > 0.f) {
    const float inv = 1.0f / sqrtf(ss);
    for (int j = lane; j < d; j += 32) xr[j] *= inv;
  }
}
}  // namespace

void renorm_rows(int64_t n, int d, float* x, cudaStream_t st) {
  if (n == 0) return;
  renorm_rows_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(n, d, x);
  ABSB_CUDA(cudaGetLastError());
}

void rand_perm_export(int64_t n, int64_t seed, int* out) {
  std::vector<int> p = rand_perm_host(n, seed);
  std::copy(p.begin(), p.end(), out);
}

int64_t split_clusters_export(int d, int64_t k, int64_t n, float* hassign, float* centroids) {
  std::vector<float> h(hassign, hassign + k);
  const int64_t ns = split_clusters_host(d, k, n, h, centroids);
  std::copy(h.begin(), h.end(), hassign);
  return ns;
}

// Per-list fp32 sums (members in ascending row order) and member counts of one slice of rows — the
// local half of Clustering::compute_centroids when the training rows are spread over ranks.
void IvfIndex::centroid_sums_dev(int64_t n, const float* x, const long long* assign, float* sums, float* counts,
                                 cudaStream_t st) {
  const int64_t k = nlist;
  ABSB_CHECK(n < ((int64_t)1 << 31), ABSB_ERR_INVALID, "too many rows in one call");
  const int sms = props.sm_count;
  DBuf<unsigned> keys, keys_sorted, hist, off;
  DBuf<int> vals, perm;
  keys.alloc_exact(std::max<int64_t>(n, 1)); keys_sorted.alloc_exact(std::max<int64_t>(n, 1));
  vals.alloc_exact(std::max<int64_t>(n, 1)); perm.alloc_exact(std::max<int64_t>(n, 1));
  hist.alloc_exact(k + 2); off.alloc_exact(k + 2);
  int bits = 1;
  while ((1ll << bits) < (long long)k + 1) ++bits;
  size_t tmp_bytes = 0, t2 = 0;
  ABSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, perm.p, (int)n, 0, bits, st));
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, hist.p, off.p, (int)k + 2, st));
  ws_cub.reserve(std::max(tmp_bytes, t2) + 16);
  ABSB_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned) * (k + 2), st));
  if (n > 0) {
    make_keys_kernel<<<grid_for(n, 256, sms), 256, 0, st>>>(n, assign, (int)k, 0, 1, keys.p, vals.p, hist.p);
    ABSB_CUDA(cudaGetLastError());
    size_t tb = ws_cub.cap;
    ABSB_CUDA(cub::DeviceRadixSort::SortPairs(ws_cub.p, tb, keys.p, keys_sorted.p, vals.p, perm.p, (int)n, 0, bits, st));
  }
  size_t tb = ws_cub.cap;
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(ws_cub.p, tb, hist.p, off.p, (int)k + 2, st));
  centroid_mean_kernel<<<(unsigned)k, std::min(256, d), 0, st>>>(d, off.p, perm.p, x, sums, counts, 0);
  ABSB_CUDA(cudaGetLastError());
  ABSB_CUDA(cudaStreamSynchronize(st));  // scratch buffers die with this frame
}

void IvfIndex::train_dev(int64_t n, const float* x, cudaStream_t st) {
  const int64_t k = nlist;
  ABSB_CHECK(n >= k, ABSB_ERR_INVALID,
             "Number of training points (%lld) should be at least as large as number of clusters (%lld)",
             (long long)n, (long long)k);
  ABSB_CHECK(n < ((int64_t)1 << 31), ABSB_ERR_INVALID, "training set too large");
  const int sms = props.sm_count;
  const int d4 = d / 4;
  DBuf<float> sample;
  DBuf<int> rows;
  const float* xs = x;
  int64_t ns = n;
  if (n > k * cp.max_points_per_centroid) {
    ns = k * cp.max_points_per_centroid;
    std::vector<int> perm = rand_perm_host(n, cp.seed);
    rows.alloc_exact(ns);
    ABSB_CUDA(cudaMemcpyAsync(rows.p, perm.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    sample.alloc_exact((size_t)ns * d);
    gather_rows_kernel<<<grid_for(ns * 32, 256, sms), 256, 0, st>>>(ns, d4, rows.p, x, sample.p);
    ABSB_CUDA(cudaGetLastError());
    ABSB_CUDA(cudaStreamSynchronize(st));
    xs = sample.p;
  }
  if (ns == k) {
    ABSB_CUDA(cudaMemcpyAsync(centroids.p, xs, sizeof(float) * (size_t)k * d, cudaMemcpyDeviceToDevice, st));
    trained = true;
    c3_dirty = true;
    return;
  }
  {
    std::vector<int> perm = rand_perm_host(ns, cp.seed + 1);
    rows.alloc_exact(k);
    ABSB_CUDA(cudaMemcpyAsync(rows.p, perm.data(), sizeof(int) * k, cudaMemcpyHostToDevice, st));
    gather_rows_kernel<<<grid_for(k * 32, 256, sms), 256, 0, st>>>(k, d4, rows.p, xs, centroids.p);
    ABSB_CUDA(cudaGetLastError());
    if (cp.spherical) renorm_rows(k, d, centroids.p, st);  // post_process_centroids before the first assignment
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  trained = true;  // coarse_dev needs it; centroids are valid from here on
  c3_dirty = true;

  DBuf<long long> assign;
  DBuf<unsigned> keys, keys_sorted, hist, off;
  DBuf<int> vals, perm;
  DBuf<float> hassign_d;
  assign.alloc_exact(ns); keys.alloc_exact(ns); keys_sorted.alloc_exact(ns); vals.alloc_exact(ns);
  perm.alloc_exact(ns); hist.alloc_exact(k + 2); off.alloc_exact(k + 2); hassign_d.alloc_exact(k);
  int bits = 1;
  while ((1ll << bits) < (long long)k + 1) ++bits;
  size_t tmp_bytes = 0, t2 = 0;
  ABSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, perm.p, (int)ns, 0, bits, st));
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, hist.p, off.p, (int)k + 2, st));
  ws_cub.reserve(std::max(tmp_bytes, t2) + 16);
  std::vector<float> hassign(k);
  std::vector<float> hcent;

  for (int it = 0; it < cp.niter; ++it) {
    assign_dev(ns, xs, assign.p, st);
    ABSB_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned) * (k + 2), st));
    make_keys_kernel<<<grid_for(ns, 256, sms), 256, 0, st>>>(ns, assign.p, (int)k, 0, 1, keys.p, vals.p, hist.p);
    ABSB_CUDA(cudaGetLastError());
    size_t tb = ws_cub.cap;
    ABSB_CUDA(cub::DeviceRadixSort::SortPairs(ws_cub.p, tb, keys.p, keys_sorted.p, vals.p, perm.p, (int)ns, 0, bits, st));
    tb = ws_cub.cap;
    ABSB_CUDA(cub::DeviceScan::ExclusiveSum(ws_cub.p, tb, hist.p, off.p, (int)k + 2, st));
    centroid_mean_kernel<<<(unsigned)k, std::min(256, d), 0, st>>>(d, off.p, perm.p, xs, centroids.p, hassign_d.p);
    ABSB_CUDA(cudaGetLastError());
    c3_dirty = true;
    ABSB_CUDA(cudaMemcpyAsync(hassign.data(), hassign_d.p, sizeof(float) * k, cudaMemcpyDeviceToHost, st));
    ABSB_CUDA(cudaStreamSynchronize(st));
    bool any_empty = false;
    for (int64_t c = 0; c < k; ++c) any_empty |= (hassign[c] == 0.f);
    if (any_empty) {
      hcent.resize((size_t)k * d);
      ABSB_CUDA(cudaMemcpy(hcent.data(), centroids.p, sizeof(float) * (size_t)k * d, cudaMemcpyDeviceToHost));
      split_clusters_host(d, k, ns, hassign, hcent.data());
      ABSB_CUDA(cudaMemcpy(centroids.p, hcent.data(), sizeof(float) * (size_t)k * d, cudaMemcpyHostToDevice));
      c3_dirty = true;
    }
    if (cp.spherical) renorm_rows(k, d, centroids.p, st);  // post_process_centroids: after mean + split
  }
}

void IvfIndex::train_host(int64_t n, const float* x) {
  const int64_t k = nlist;
  ABSB_CHECK(n >= k, ABSB_ERR_INVALID,
             "Number of training points (%lld) should be at least as large as number of clusters (%lld)",
             (long long)n, (long long)k);
  // Subsample on the host so that only the sample crosses PCIe, then run the device trainer on it
  // with subsampling disabled (the permutation must not be applied twice).
  cudaStream_t st = own_stream;
  DBuf<float> sample;
  int64_t ns = n;
  if (n > k * cp.max_points_per_centroid) {
    ns = k * cp.max_points_per_centroid;
    std::vector<int> perm = rand_perm_host(n, cp.seed);
    sample.alloc_exact((size_t)ns * d);
    const int64_t step = std::max<int64_t>(1, ((int64_t)64 << 20) / (d * (int64_t)sizeof(float)));
    HBuf<float> stage;
    stage.reserve((size_t)step * d);
    for (int64_t r0 = 0; r0 < ns; r0 += step) {
      const int64_t nr = std::min(step, ns - r0);
      for (int64_t i = 0; i < nr; ++i)
        std::copy(x + (size_t)perm[r0 + i] * d, x + (size_t)(perm[r0 + i] + 1) * d, stage.p + (size_t)i * d);
      ABSB_CUDA(cudaMemcpyAsync(sample.p + (size_t)r0 * d, stage.p, sizeof(float) * (size_t)nr * d,
                                cudaMemcpyHostToDevice, st));
      ABSB_CUDA(cudaStreamSynchronize(st));
    }
  } else {
    sample.alloc_exact((size_t)ns * d);
    ABSB_CUDA(cudaMemcpyAsync(sample.p, x, sizeof(float) * (size_t)ns * d, cudaMemcpyHostToDevice, st));
    ABSB_CUDA(cudaStreamSynchronize(st));
  }
  const int saved = cp.max_points_per_centroid;
  // ns <= k * max_ppc already holds, so train_dev will not subsample again
  train_dev(ns, sample.p, st);
  cp.max_points_per_centroid = saved;
  ABSB_CUDA(cudaStreamSynchronize(st));
}

// ---------------------------------------------------------------- search -------------------
void IvfIndex::fold_stats() {
  if (!stats_pending) return;
  unsigned long long h[2] = {0, 0};
  ABSB_CUDA(cudaMemcpy(h, ws_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
  stats.vectors += (int64_t)h[0];
  stats.items += (int64_t)h[1];
  stats.bytes = stats.vectors * ((int64_t)d * 4 + 8);
  stats_pending = false;
}

void IvfIndex::search_preassigned_dev(int64_t nq, const float* q, int k, int nprobe,
                                      const long long* coarse, float* D, long long* I,
                                      cudaStream_t st, const SearchPush* push) {
  ABSB_CHECK(trained, ABSB_ERR_STATE, "index is not trained");
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  ABSB_CHECK(nprobe >= 1, ABSB_ERR_INVALID, "nprobe=%d", nprobe);
  if (nq == 0) return;
  const int64_t per_q = std::max<int64_t>(1, items_bound_per_query(nprobe, st));
  // bound the partial-result buffers to ~1 GiB per launch
  const int k_part = std::max(k, two_stage_k);  // entries per work item in the partial-result buffers
  int64_t nb_max = std::min<int64_t>(kMaxPlanQueries, std::max<int64_t>(1, ((int64_t)1 << 30) / (per_q * k_part * 12)));
  for (int64_t q0 = 0; q0 < nq; q0 += nb_max) {
    const int nb = (int)std::min(nb_max, nq - q0);
    const int64_t max_items = per_q * nb;
    ABSB_CHECK(max_items < ((int64_t)1 << 31), ABSB_ERR_UNSUPPORTED, "too many scan items");
    ws_items.reserve((size_t)max_items);
    ws_part_s.reserve((size_t)max_items * k_part);
    ws_part_id.reserve((size_t)max_items * k_part);
    ws_q_begin.reserve(kMaxPlanQueries + 1);
    if (stats_pending) fold_stats();  // only one plan's numbers fit in ws_stats
    {
      Span sp(this, st, 2);
      const size_t npairs_max = (size_t)kMaxPlanQueries * nprobe;
      ws_pair_counts.reserve(npairs_max + 1);
      ws_pair_offs.reserve(npairs_max + 1);
      ws_plan_tmp.reserve(plan_scan_tmp_bytes((int)npairs_max) + 16);
      PlanOrderWs ow{};
      if (scan_order == 1) {
        ws_okeys.reserve(npairs_max + 1); ws_okeys_sorted.reserve(npairs_max + 1);
        ws_ovals.reserve(npairs_max + 1); ws_ovals_sorted.reserve(npairs_max + 1);
        ws_ocounts.reserve(npairs_max + 1); ws_oqoffs.reserve(npairs_max + 1);
        ws_order.reserve((size_t)max_items);
        ow = PlanOrderWs{ws_okeys.p, ws_okeys_sorted.p, ws_ovals.p, ws_ovals_sorted.p, ws_ocounts.p, ws_oqoffs.p, ws_order.p};
      }
      launch_plan(table(), coarse + q0 * nprobe, nb, nprobe, scan_chunk, (int)max_items, ws_items.p,
                  ws_q_begin.p, ws_counters.p, ws_counters.p + 1, ws_stats.p, ws_pair_counts.p, ws_pair_offs.p,
                  ws_plan_tmp.p, ws_plan_tmp.cap, scan_order == 1 ? &ow : nullptr, st);
    }
    ScanLaunch a;
    a.Q = q + q0 * d;
    a.d = d;
    a.k = k;
    a.items = ws_items.p;
    a.n_items = ws_counters.p;
    a.queue_counter = ws_counters.p + 1;
    a.order = scan_order == 1 ? ws_order.p : nullptr;
    a.part_s = ws_part_s.p;
    a.part_id = ws_part_id.p;
    a.sm_count = props.sm_count;
    a.ctas_per_sm = scan_ctas_per_sm;
    const bool two_stage = two_stage_k >= k && pool.shadow && d == 1024;
    if (two_stage) {
      // with a push target the stages work on a local (D, I) and the LAST merge stores every row into
      // the peers' buffers (fallback rows merged, proven rows copied): no separate pack + push kernel
      float* Dl = D ? D + q0 * k : nullptr;
      long long* Il = I ? I + q0 * k : nullptr;
      if (push) {
        ws_push_D.reserve((size_t)kMaxPlanQueries * k);
        ws_push_I.reserve((size_t)kMaxPlanQueries * k);
        Dl = ws_push_D.p;
        Il = ws_push_I.p;
      }
      // ---- stage 1: fp16 shadow codes -> shortlist of K approximate candidates per query (ivf_scan16.cu)
      const int K = two_stage_k;
      ws_short_s.reserve((size_t)kMaxPlanQueries * K);
      ws_short_g.reserve((size_t)kMaxPlanQueries * K);
      ws_items2.reserve((size_t)kMaxPlanQueries * K);
      ws_part2_s.reserve((size_t)kMaxPlanQueries * K * k);
      ws_part2_id.reserve((size_t)kMaxPlanQueries * K * k);
      ws_q_begin2.reserve(kMaxPlanQueries + 1);
      ws_flags.reserve(kMaxPlanQueries);
      Scan16Launch s16;
      s16.Q = a.Q;
      s16.K = K;
      s16.items = ws_items.p;
      s16.n_items = ws_counters.p;
      s16.queue_counter = ws_counters.p + 1;
      s16.order = a.order;
      s16.part_s = ws_part_s.p;
      s16.part_g = ws_part_id.p;
      s16.half_slabs = pool.d_half_slabs.p;
      s16.slab_shift = pool.slab_shift;
      s16.page_vecs = pool.page_vecs;
      s16.sm_count = props.sm_count;
      s16.ctas_per_sm = scan_ctas_per_sm;
      {
        Span sp(this, st, 3);
        if (use_ring()) {
          ScanRing r16 = ring;
          r16.stage_vecs = ring.stage_vecs * 2;  // same stage bytes as the fp32 ring
          launch_scan16_ring(s16, r16, st);
        } else {
          launch_scan16(s16, st);
        }
      }
      {
        Span sp(this, st, 2);
        merge_partials(nb, K, ws_q_begin.p, ws_part_s.p, ws_part_id.p, ws_short_s.p, ws_short_g.p, st);
        launch_rescore_items(table(), nb, K, ws_short_g.p, ws_items2.p, ws_q_begin2.p, ws_counters2.p,
                             ws_counters2.p + 1, st);
      }
      // ---- stage 2: exact fp32 scores of the shortlist with the single-pass kernel
      ScanLaunch a2 = a;
      a2.items = ws_items2.p;
      a2.n_items = ws_counters2.p;
      a2.queue_counter = ws_counters2.p + 1;
      a2.order = nullptr;
      a2.part_s = ws_part2_s.p;
      a2.part_id = ws_part2_id.p;
      {
        Span sp(this, st, 0);
        launch_scan(a2, st);
      }
      {
        Span sp(this, st, 2);
        merge_partials(nb, k, ws_q_begin2.p, ws_part2_s.p, ws_part2_id.p, Dl, Il, st);
        launch_two_stage_check(nb, d, k, K, a.Q, Dl, Il, ws_short_s.p, ws_short_g.p, ws_maxima.p,
                               ws_flags.p, ws_nflag.p, st);
        // ---- fallback: the queries the bound could not prove go through the single-pass scan (the plan
        // emits no work for the others, so this costs a few empty launches when everything was proven)
        launch_plan(table(), coarse + q0 * nprobe, nb, nprobe, scan_chunk, (int)max_items, ws_items.p, ws_q_begin.p,
                    ws_counters.p, ws_counters.p + 1, ws_stats2.p, ws_pair_counts.p, ws_pair_offs.p, ws_plan_tmp.p,
                    ws_plan_tmp.cap, nullptr, st, ws_flags.p);
      }
      ScanLaunch a3 = a;
      a3.order = nullptr;
      {
        Span sp(this, st, 0);
        run_scan(a3, st);
      }
      {
        Span sp(this, st, 2);
        if (push) {
          PeerPush pp = push->pp;
          pp.q_off = push->q_base + q0;
          if (push->q_base + q0 + nb != push->nq_total) pp.epoch = 0;  // record not complete yet
          merge_partials_push(nb, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, pp, st, ws_flags.p, Dl, Il);
        } else {
          merge_partials(nb, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, Dl, Il, st, ws_flags.p);
        }
      }
      have_last_scan = false;
      stats_pending = true;
      stats.launches += (scan_order == 1 ? 11 : 5) + 12;
      if (q0 + nb_max < nq) fold_stats();
      continue;
    }
    {
      Span sp(this, st, 0);
      run_scan(a, st);
    }
    {
      Span sp(this, st, 2);
      if (push) {
        PeerPush pp = push->pp;
        pp.q_off = push->q_base + q0;
        if (push->q_base + q0 + nb != push->nq_total) pp.epoch = 0;  // record not complete yet
        merge_partials_push(nb, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, pp, st);
      } else {
        merge_partials(nb, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, D + q0 * k, I + q0 * k, st);
      }
    }
    last_scan = a;
    have_last_scan = true;
    stats_pending = true;
    stats.launches += scan_order == 1 ? 11 : 5;
    if (q0 + nb_max < nq) fold_stats();
  }
}

void IvfIndex::search_dev(int64_t nq, const float* q, int k, int nprobe, float* D, long long* I,
                          cudaStream_t st, const SearchPush* push) {
  ABSB_CHECK(trained, ABSB_ERR_STATE, "index is not trained");
  const int np = std::min(nprobe, nlist);
  ABSB_CHECK(np >= 1 && np <= ABSB_MAX_K, ABSB_ERR_INVALID, "nprobe=%d outside [1,%d]", nprobe, ABSB_MAX_K);
  const int64_t step = kMaxPlanQueries;
  ws_coarse_s.reserve((size_t)step * np);
  ws_coarse_i.reserve((size_t)step * np);
  for (int64_t q0 = 0; q0 < nq; q0 += step) {
    const int64_t nb = std::min(step, nq - q0);
    coarse_dev(nb, q + q0 * d, np, ws_coarse_s.p, ws_coarse_i.p, true, st);
    if (push) {
      SearchPush sub = *push;
      sub.q_base = push->q_base + q0;
      search_preassigned_dev(nb, q + q0 * d, k, np, ws_coarse_i.p, nullptr, nullptr, st, &sub);
    } else {
      search_preassigned_dev(nb, q + q0 * d, k, np, ws_coarse_i.p, D + q0 * k, I + q0 * k, st);
    }
  }
}

// =========================================================================================
// FlatIndex
// =========================================================================================
FlatIndex::FlatIndex(int d_, int device_) : d(d_), device(device_) {
  ABSB_CHECK(d > 0 && d % 4 == 0, ABSB_ERR_UNSUPPORTED, "d must be a positive multiple of 4 (d=%d)", d);
  props = device_props(device);
  DeviceGuard g(device);
  ABSB_CUDA(cudaStreamCreate(&own_stream));
}

FlatIndex::~FlatIndex() {
  cudaSetDevice(device);
  if (own_stream) {
    cudaStreamSynchronize(own_stream);
    cudaStreamDestroy(own_stream);
  }
}

void FlatIndex::add_dev(int64_t n, const float* x, cudaStream_t st) {
  if (n == 0) return;
  if ((size_t)(ntotal + n) * d > xb.cap) {
    DBuf<float> bigger;
    bigger.alloc_exact(std::max<size_t>((size_t)(ntotal + n) * d, xb.cap * 2));
    if (ntotal)
      ABSB_CUDA(cudaMemcpyAsync(bigger.p, xb.p, sizeof(float) * (size_t)ntotal * d, cudaMemcpyDeviceToDevice, st));
    ABSB_CUDA(cudaStreamSynchronize(st));
    xb = std::move(bigger);
  }
  ABSB_CUDA(cudaMemcpyAsync(xb.p + (size_t)ntotal * d, x, sizeof(float) * (size_t)n * d, cudaMemcpyDeviceToDevice, st));
  ntotal += n;
}

void FlatIndex::search_dev(int64_t nq, const float* q, int k, float* D, long long* I, cudaStream_t st) {
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  if (nq == 0) return;
  const int64_t col_chunk = 65536;
  const int nchunks = (int)std::max<int64_t>(1, ceil_div(ntotal, col_chunk));
  const int64_t row_step = 1024;
  ws_scores.reserve((size_t)row_step * std::min<int64_t>(col_chunk, std::max<int64_t>((ntotal + 31) & ~(int64_t)31, 1)));
  ws_part_s.reserve((size_t)row_step * nchunks * k);
  ws_part_id.reserve((size_t)row_step * nchunks * k);
  ws_q_begin.reserve(row_step + 1);
  std::vector<int> qb(row_step + 1);
  for (int i = 0; i <= row_step; ++i) qb[i] = i * nchunks;
  ABSB_CUDA(cudaMemcpyAsync(ws_q_begin.p, qb.data(), sizeof(int) * (row_step + 1), cudaMemcpyHostToDevice, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  for (int64_t r0 = 0; r0 < nq; r0 += row_step) {
    const int64_t nr = std::min(row_step, nq - r0);
    if (ntotal == 0) {
      // every slot missing: merge of zero candidates writes the padding
      std::vector<int> zero(nr + 1, 0);
      ABSB_CUDA(cudaMemcpyAsync(ws_q_begin.p, zero.data(), sizeof(int) * (nr + 1), cudaMemcpyHostToDevice, st));
      ABSB_CUDA(cudaStreamSynchronize(st));
      merge_partials((int)nr, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, D + r0 * k, I + r0 * k, st);
      continue;
    }
    // The all-pairs inner products run on tcgen05 as a split-bf16 GEMM (x = hi + mid + lo in bf16, six products
    // accumulated in fp32: fp32-faithful scores, exact on the lattice corpus) whenever the shapes allow; the
    // database chunk is split on the fly, so no second copy of the corpus is kept.  d % 64 != 0 falls back to
    // the fp32 FFMA GEMM.
    const bool tensor = d % 64 == 0;
    if (tensor) {
      ws_q3.reserve((size_t)row_step * 3 * d);
      split3_bf16(nr, d, q + r0 * d, ws_q3.p, st);
    }
    for (int c = 0; c < nchunks; ++c) {
      const int64_t c0 = (int64_t)c * col_chunk;
      const int nc = (int)std::min(col_chunk, ntotal - c0);
      if (tensor) {
        const int ncp = (nc + 31) & ~31;  // the GEMM wants N % 32 == 0: zero rows pad the last chunk
        ws_x3.reserve((size_t)std::min<int64_t>(col_chunk, (ntotal + 31) & ~(int64_t)31) * 3 * d);
        split3_bf16(nc, d, xb.p + (size_t)c0 * d, ws_x3.p, st);
        if (ncp > nc)
          ABSB_CUDA(cudaMemsetAsync(ws_x3.p + (size_t)nc * 3 * d, 0, sizeof(uint16_t) * (size_t)(ncp - nc) * 3 * d, st));
        gemm_split3_f32((int)nr, ncp, d, ws_q3.p, ws_x3.p, ws_scores.p, ncp, props.sm_count, st);
        select_rows(ws_scores.p, ncp, nr, nc, c0, k, ws_part_s.p + (size_t)c * k, ws_part_id.p + (size_t)c * k,
                    (int64_t)nchunks * k, false, st);
      } else {
        gemm_nt_f32((int)nr, nc, d, q + r0 * d, d, xb.p + (size_t)c0 * d, d, ws_scores.p, nc, st);
        select_rows(ws_scores.p, nc, nr, nc, c0, k, ws_part_s.p + (size_t)c * k, ws_part_id.p + (size_t)c * k,
                    (int64_t)nchunks * k, false, st);
      }
    }
    merge_partials((int)nr, k, ws_q_begin.p, ws_part_s.p, ws_part_id.p, D + r0 * k, I + r0 * k, st);
  }
}

}  // namespace absb
