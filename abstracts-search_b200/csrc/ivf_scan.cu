// ivf_scan.cu — the fine scan of IndexIVFFlat::search: for every (query, probed list) pair, the
// inner product of the query with every vector of the list, keeping the k best.
//
// Stands in for faiss's IVFFlatScanner::scan_codes + heap_replace_top under
// IndexIVF::search_preassigned — what `sidecar-search index tune` (/root/reference/Makefile:31-32)
// and app.py's query loop (/root/reference/README.md:16,28) spend their time in.
//
// B200 design (HBM-bound: 2 flop per 4 bytes):
//   * Inverted lists live in fixed-size pages (P vectors, ~64 KB) inside large slabs; a list is a
//     sequence of pages in insertion order.  A *plan* kernel turns the coarse result
//     [nq, nprobe] into a flat queue of work items, each a run of physically contiguous vectors
//     (<= chunk) of one list for one query.
//   * The *scan* kernel is persistent (grid = SMs x resident CTAs): every WARP pulls items off a
//     global atomic queue (next index prefetched while the current item streams).  A warp reads
//     one 4 KB vector as 8 x 128-bit fully coalesced streaming loads per lane (ld.global.cs),
//     U vectors in flight, FMAs against the query held in registers, butterfly-reduces, and
//     tests the score against the warp-uniform k-th-best threshold of a register-resident
//     WarpTopK.  No shared memory, no block barrier, no intermediate score array.
//   * Each item emits k partial results; merge_partials (dense.cu) reduces them per query.
// Algorithmic bytes per launch = sum over items of len * (4 d + 8).
#include <cub/cub.cuh>

#include "common.cuh"
#include "ivf_scan.cuh"
#include "topk.cuh"

namespace absb {

namespace {

// ------------------------------------------------------------------ plan --------------------
// count (one thread per (query, probe) pair) -> exclusive scan (cub) -> emit: three small grid-wide
// launches instead of one serial CTA.
__device__ __forceinline__ int walk_list(const ListTable& lt, long long l, int chunk, int q,
                                         ScanItem* out /* nullptr = count only */,
                                         long long* vectors) {
  if (l < 0 || l >= lt.nlist) return 0;
  const long long size = lt.list_size[l];
  if (size == 0) return 0;
  const long long pb = lt.pt_off[l];
  const int P = lt.page_vecs;
  const long long npages = (size + P - 1) / P;
  int n = 0;
  long long run_first = -1, run_len = 0;  // run of contiguous pages, length in vectors
  auto flush = [&]() {
    if (run_len == 0) return;
    if (out) {
      const int slab = (int)(run_first >> lt.slab_shift);
      const long long in_slab = run_first & ((1ll << lt.slab_shift) - 1);
      ScanItem it;
      it.codes = lt.code_slabs[slab] + (size_t)in_slab * P * lt.d;
      it.ids = lt.id_slabs[slab] + (size_t)in_slab * P;
      it.len = (int)run_len;
      it.q = q;
      it.pad = run_first * P;
      out[n] = it;
    }
    ++n;
    run_len = 0;
  };
  for (long long i = 0; i < npages; ++i) {
    const long long pg = lt.pt_pages[pb + i];
    const long long nv = (i == npages - 1) ? size - i * P : P;
    const bool contiguous = run_len > 0 && pg == run_first + run_len / P &&
                            (pg >> lt.slab_shift) == (run_first >> lt.slab_shift) &&
                            run_len + nv <= chunk;
    if (!contiguous) {
      flush();
      run_first = pg;
    }
    run_len += nv;
  }
  flush();
  if (vectors) *vectors += size;
  return n;
}

// Pass 1: one thread per (query, probe) pair counts the work items of that list.
__global__ void plan_count_kernel(ListTable lt, const long long* __restrict__ coarse, int npairs, int chunk,
                                  int* __restrict__ counts /* [npairs + 1] */,
                                  unsigned* __restrict__ keys /* [npairs] list number, or nullptr */,
                                  int* __restrict__ vals /* [npairs] pair index */,
                                  unsigned long long* __restrict__ stats, int nprobe,
                                  const unsigned char* __restrict__ active) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  long long vec = 0;
  if (i < npairs) {
    const long long l = coarse[i];
    const bool on = active == nullptr || active[i / nprobe] != 0;
    counts[i] = on ? walk_list(lt, l, chunk, 0, nullptr, &vec) : 0;
    if (keys) {
      keys[i] = (l >= 0 && l < lt.nlist) ? (unsigned)l : (unsigned)lt.nlist;
      vals[i] = i;
    }
  }
  if (i == npairs) counts[i] = 0;
  // vectors scanned = sum of probed list sizes (statistics only)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vec += __shfl_xor_sync(kFullMask, vec, o);
  if ((threadIdx.x & 31) == 0 && vec) atomicAdd(&stats[0], (unsigned long long)vec);
}

// Pass 3 (after an exclusive scan of the counts): every pair writes its items at its offset.
__global__ void plan_emit_kernel(ListTable lt, const long long* __restrict__ coarse, int nq, int nprobe, int chunk,
                                 int max_items, const int* __restrict__ offs /* [npairs + 1] */,
                                 ScanItem* __restrict__ items, int* __restrict__ q_begin /* [nq + 1] */,
                                 int* __restrict__ n_items, int* __restrict__ queue_counter,
                                 unsigned long long* __restrict__ stats) {
  const int npairs = nq * nprobe;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = offs[npairs];
  if (i == 0) {
    *n_items = total <= max_items ? total : -1;  // -1: capacity bug, the scan does nothing
    *queue_counter = 0;
    stats[1] = (unsigned long long)total;
  }
  if (i <= nq) q_begin[i] = offs[min(i * nprobe, npairs)];
  if (i >= npairs || total > max_items) return;
  const int off = offs[i];
  if (offs[i + 1] > off) walk_list(lt, coarse[i], chunk, i / nprobe, items + off, nullptr);
}

// Work items per list under the current chunk length (the exact count the plan would emit for it):
// the host sums the nprobe largest to size the partial-result buffers of a search.
__global__ void count_list_items_kernel(ListTable lt, int chunk, int* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < lt.nlist) out[l] = walk_list(lt, l, chunk, 0, nullptr, nullptr);
}

// ------------------------------------------------------------------ scan --------------------
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

__device__ __forceinline__ ScanItem load_item(const ScanItem* p) {
  // 32-byte item, two 16-byte loads (all lanes the same address: one broadcast transaction)
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  ScanItem it;
  it.codes = reinterpret_cast<const float*>(((unsigned long long)(unsigned)a.y << 32) | (unsigned)a.x);
  it.ids = reinterpret_cast<const long long*>(((unsigned long long)(unsigned)a.w << 32) | (unsigned)a.z);
  it.len = b.x;
  it.q = b.y;
  return it;
}

// Fast path: d = 128 * D4, query in registers, U vectors in flight per warp.
template <int SLOTS, int D4, int U>
__global__ __launch_bounds__(kScanThreads) void ivf_scan_kernel(
    const float* __restrict__ Q, const ScanItem* __restrict__ items, const int* __restrict__ n_items_ptr,
    int* __restrict__ queue_counter, const int* __restrict__ order, int k, float* __restrict__ part_s,
    long long* __restrict__ part_id) {
  constexpr int d4 = D4 * 32;  // float4 per vector
  const int lane = threadIdx.x & 31;
  const int n_items = *n_items_ptr;
  int pos = 0;
  if (lane == 0) pos = atomicAdd(queue_counter, 1);
  pos = __shfl_sync(kFullMask, pos, 0);
  while (pos < n_items) {
    int next = 0;
    if (lane == 0) next = atomicAdd(queue_counter, 1);  // latency hidden behind this item's scan
    const int item = order ? __ldg(order + pos) : pos;  // queue position -> item (list-major order)
    const ScanItem it = load_item(items + item);
    const float4* qp = reinterpret_cast<const float4*>(Q) + (size_t)it.q * d4 + lane;
    float4 qv[D4];
#pragma unroll
    for (int j = 0; j < D4; ++j) qv[j] = __ldg(qp + j * 32);

    WarpTopK<SLOTS> tk;
    tk.init(k, lane);
    const float4* cp = reinterpret_cast<const float4*>(it.codes) + lane;
    int v = 0;
    for (; v + U <= it.len; v += U) {
      float4 x[U][D4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < D4; ++j) x[u][j] = __ldcs(cp + (size_t)(v + u) * d4 + j * 32);
      float acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = 0.f;
#pragma unroll
        for (int j = 0; j < D4; ++j) acc[u] = dot4(x[u][j], qv[j], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = warp_sum(acc[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (tk.may_enter(acc[u])) {
          const long long id = __ldg(it.ids + v + u);
          if (tk.admits(acc[u], id)) tk.insert(acc[u], id);
        }
      }
    }
    for (; v < it.len; ++v) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < D4; ++j) acc = dot4(__ldcs(cp + (size_t)v * d4 + j * 32), qv[j], acc);
      acc = warp_sum(acc);
      if (tk.may_enter(acc)) {
        const long long id = __ldg(it.ids + v);
        if (tk.admits(acc, id)) tk.insert(acc, id);
      }
    }
    tk.store(part_s + (size_t)item * k, part_id + (size_t)item * k);
    pos = __shfl_sync(kFullMask, next, 0);
  }
}

// Generic path: any d with d % 4 == 0; the query sits in shared memory (one copy per warp).
template <int SLOTS>
__global__ __launch_bounds__(kScanThreads) void ivf_scan_generic_kernel(
    const float* __restrict__ Q, int d, const ScanItem* __restrict__ items,
    const int* __restrict__ n_items_ptr, int* __restrict__ queue_counter, const int* __restrict__ order, int k,
    float* __restrict__ part_s, long long* __restrict__ part_id) {
  extern __shared__ __align__(16) float sm_q[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = d / 4;
  float4* myq = reinterpret_cast<float4*>(sm_q) + (size_t)warp * d4;
  const int n_items = *n_items_ptr;
  int pos = 0;
  if (lane == 0) pos = atomicAdd(queue_counter, 1);
  pos = __shfl_sync(kFullMask, pos, 0);
  while (pos < n_items) {
    int next = 0;
    if (lane == 0) next = atomicAdd(queue_counter, 1);
    const int item = order ? __ldg(order + pos) : pos;
    const ScanItem it = load_item(items + item);
    const float4* qp = reinterpret_cast<const float4*>(Q) + (size_t)it.q * d4;
    __syncwarp();
    for (int j = lane; j < d4; j += 32) myq[j] = __ldg(qp + j);
    __syncwarp();
    WarpTopK<SLOTS> tk;
    tk.init(k, lane);
    const float4* cp = reinterpret_cast<const float4*>(it.codes);
    for (int v = 0; v < it.len; ++v) {
      float acc = 0.f;
      for (int j = lane; j < d4; j += 32) acc = dot4(__ldcs(cp + (size_t)v * d4 + j), myq[j], acc);
      acc = warp_sum(acc);
      if (tk.may_enter(acc)) {
        const long long id = __ldg(it.ids + v);
        if (tk.admits(acc, id)) tk.insert(acc, id);
      }
    }
    tk.store(part_s + (size_t)item * k, part_id + (size_t)item * k);
    pos = __shfl_sync(kFullMask, next, 0);
  }
}

template <typename Kern>
int resident_ctas(Kern kern, size_t smem) {
  int n = 0;
  ABSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kScanThreads, smem));
  return n < 1 ? 1 : n;
}

// List-major queue order: pairs sorted by list number, so that the probes of different queries into the
// same inverted list are scanned at the same time and the later ones hit L2 instead of HBM.
__global__ void plan_gather_counts_kernel(int npairs, const int* __restrict__ pair_sorted, const int* __restrict__ counts,
                                          int* __restrict__ counts_sorted /* [npairs + 1] */) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < npairs) counts_sorted[s] = counts[pair_sorted[s]];
  if (s == npairs) counts_sorted[s] = 0;
}

__global__ void plan_order_kernel(int npairs, int max_items, const int* __restrict__ pair_sorted,
                                  const int* __restrict__ offs /* item offsets, pair order */,
                                  const int* __restrict__ qoffs /* queue offsets, sorted order */,
                                  int* __restrict__ order) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= npairs || offs[npairs] > max_items) return;
  const int pi = pair_sorted[s];
  const int b = offs[pi], n = offs[pi + 1] - b, q = qoffs[s];
  for (int c = 0; c < n; ++c) order[q + c] = b + c;
}

}  // namespace

void launch_plan(const ListTable& lt, const long long* coarse, int nq, int nprobe, int chunk,
                 int max_items, ScanItem* items, int* q_begin, int* n_items, int* queue_counter,
                 unsigned long long* stats, int* pair_counts, int* pair_offs, void* scan_tmp,
                 size_t scan_tmp_bytes, const PlanOrderWs* ow, cudaStream_t st, const unsigned char* active) {
  ABSB_CHECK(nq >= 1 && nq <= kMaxPlanQueries, ABSB_ERR_INVALID, "plan: nq=%d", nq);
  const int npairs = nq * nprobe;
  ABSB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), st));
  const int threads = 256;
  const int pair_blocks = (npairs + 1 + threads - 1) / threads;
  plan_count_kernel<<<pair_blocks, threads, 0, st>>>(lt, coarse, npairs, chunk, pair_counts, ow ? ow->keys : nullptr,
                                                     ow ? ow->vals : nullptr, stats, nprobe, active);
  ABSB_CUDA(cudaGetLastError());
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, pair_counts, pair_offs, npairs + 1, st));
  if (ow) {
    int bits = 1;
    while ((1ll << bits) < (long long)lt.nlist + 1) ++bits;
    size_t tb = scan_tmp_bytes;
    ABSB_CUDA(cub::DeviceRadixSort::SortPairs(scan_tmp, tb, ow->keys, ow->keys_sorted, ow->vals, ow->vals_sorted, npairs, 0,
                                              bits, st));
    plan_gather_counts_kernel<<<pair_blocks, threads, 0, st>>>(npairs, ow->vals_sorted, pair_counts, ow->counts_sorted);
    ABSB_CUDA(cudaGetLastError());
    ABSB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, ow->counts_sorted, ow->qoffs, npairs + 1, st));
    plan_order_kernel<<<pair_blocks, threads, 0, st>>>(npairs, max_items, ow->vals_sorted, pair_offs, ow->qoffs, ow->order);
    ABSB_CUDA(cudaGetLastError());
  }
  const int work = std::max(npairs, nq + 1);
  plan_emit_kernel<<<(work + threads - 1) / threads, threads, 0, st>>>(lt, coarse, nq, nprobe, chunk, max_items,
                                                                      pair_offs, items, q_begin, n_items,
                                                                      queue_counter, stats);
  ABSB_CUDA(cudaGetLastError());
}

void launch_count_list_items(const ListTable& lt, int chunk, int* out, cudaStream_t st) {
  count_list_items_kernel<<<(lt.nlist + 255) / 256, 256, 0, st>>>(lt, chunk, out);
  ABSB_CUDA(cudaGetLastError());
}

size_t plan_scan_tmp_bytes(int max_pairs) {
  size_t bytes = 0, b2 = 0;
  ABSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int*)nullptr, (int*)nullptr, max_pairs + 1));
  ABSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr,
                                            max_pairs, 0, 32));
  return std::max(bytes, b2);
}

void launch_scan(const ScanLaunch& a, cudaStream_t st) {
  const int sms = a.sm_count;
  if (a.d == 1024) {
    constexpr int D4 = 8;
    ABSB_DISPATCH_SLOTS(a.k, {
      auto kern = ivf_scan_kernel<SLOTS, D4, kScanUnroll>;
      const int per_sm = a.ctas_per_sm > 0 ? a.ctas_per_sm : resident_ctas(kern, 0);
      kern<<<sms * per_sm, kScanThreads, 0, st>>>(a.Q, a.items, a.n_items, a.queue_counter, a.order, a.k,
                                                  a.part_s, a.part_id);
    });
  } else {
    ABSB_CHECK(a.d % 4 == 0, ABSB_ERR_UNSUPPORTED, "IVF scan needs d %% 4 == 0 (d=%d)", a.d);
    const size_t smem = (size_t)(kScanThreads / 32) * a.d * sizeof(float);
    ABSB_CHECK(smem <= 200 * 1024, ABSB_ERR_UNSUPPORTED, "d=%d too large for the generic scan", a.d);
    ABSB_DISPATCH_SLOTS(a.k, {
      auto kern = ivf_scan_generic_kernel<SLOTS>;
      if (smem > 48 * 1024)
        ABSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int per_sm = a.ctas_per_sm > 0 ? a.ctas_per_sm : resident_ctas(kern, smem);
      kern<<<sms * per_sm, kScanThreads, smem, st>>>(a.Q, a.d, a.items, a.n_items, a.queue_counter, a.order,
                                                     a.k, a.part_s, a.part_id);
    });
  }
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb
