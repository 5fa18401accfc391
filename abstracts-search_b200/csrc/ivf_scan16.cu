// ivf_scan16.cu — two-stage fine scan: the same IndexIVFFlat::search result (faiss
// IVFFlatScanner::scan_codes + k-best, /root/reference/Makefile:31-32, README.md:16,28) from half the
// HBM traffic.
//
//   stage 1  scan an fp16 SHADOW copy of the list codes (2 KB instead of 4 KB per vector) with the
//            work queue, warp top-k and merge of the single-pass scan, keeping a shortlist of the K
//            best APPROXIMATE scores per query (K = 32 / 64 / 128 >= k) as global slot numbers;
//   stage 2  re-score the shortlist from the fp32 codes with the unmodified fp32 scan kernel (one
//            single-vector work item per candidate: the scores are bit-identical to the single-pass
//            scan's) and merge to the exact top-k;
//   check    per query, prove that nothing outside the shortlist can reach rank k:
//                computed exact score of any vector  <=  its computed fp16-pass score + B(q),
//                B(q) = |q| * (r_max + 2 * 40 u * (x_max + r_max)),   u = 2^-24,
//            r_max = max |x - fp16(x)|_2 and x_max = max |x|_2 over everything ever added (tracked by
//            the scatter kernel), 40 u covering the fp32 rounding of either pass (32 chained FMAs per
//            lane + 5 butterfly levels, Cauchy-Schwarz on sum |q_i x_i|).  If the shortlist is not
//            full every probed vector is in it; otherwise the query is proven when
//            a_min + B(q) < T (a_min = K-th approximate score, T = k-th exact score).  Queries that are
//            not proven are flagged and re-done by the single-pass fp32 scan (ivf.cu), so the result
//            is exact unconditionally.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ivf_scan.cuh"
#include "topk.cuh"

namespace absb {

namespace {

constexpr int kD = 1024;
constexpr int kH4 = 4;  // 128-bit loads (8 halves) per lane per vector

__device__ __forceinline__ float warp_sum16(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

__device__ __forceinline__ float dot8(const uint4 x, const float4 qa, const float4 qb, float acc) {
  const __half2* h = reinterpret_cast<const __half2*>(&x);
  const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
  const float2 f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
  acc = fmaf(f0.x, qa.x, acc);
  acc = fmaf(f0.y, qa.y, acc);
  acc = fmaf(f1.x, qa.z, acc);
  acc = fmaf(f1.y, qa.w, acc);
  acc = fmaf(f2.x, qb.x, acc);
  acc = fmaf(f2.y, qb.y, acc);
  acc = fmaf(f3.x, qb.z, acc);
  acc = fmaf(f3.y, qb.w, acc);
  return acc;
}

struct Item16 {
  long long g0;
  int len, q;
};

__device__ __forceinline__ Item16 load_item16(const ScanItem* p) {
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);  // len, q, pad.lo, pad.hi
  Item16 it;
  it.len = b.x;
  it.q = b.y;
  it.g0 = (long long)(((unsigned long long)(unsigned)b.w << 32) | (unsigned)b.z);
  return it;
}

// Stage 1.  Same persistent warp-per-item structure as ivf_scan_kernel; a vector is 4 x 128-bit loads
// per lane, lane l owning elements (jj * 32 + l) * 8 .. + 7.
template <int SLOTS, int U>
__global__ __launch_bounds__(kScanThreads) void ivf_scan16_kernel(
    const float* __restrict__ Q, const ScanItem* __restrict__ items, const int* __restrict__ n_items_ptr,
    int* __restrict__ queue_counter, const int* __restrict__ order, int K, float* __restrict__ part_s,
    long long* __restrict__ part_g, const unsigned short* const* __restrict__ half_slabs, int slab_shift, int P) {
  const int lane = threadIdx.x & 31;
  const int n_items = *n_items_ptr;
  const long long slab_mask = (1ll << slab_shift) - 1;
  int pos = 0;
  if (lane == 0) pos = atomicAdd(queue_counter, 1);
  pos = __shfl_sync(kFullMask, pos, 0);
  while (pos < n_items) {
    int next = 0;
    if (lane == 0) next = atomicAdd(queue_counter, 1);
    const int item = order ? __ldg(order + pos) : pos;
    const Item16 it = load_item16(items + item);
    const float4* qp = reinterpret_cast<const float4*>(Q) + (size_t)it.q * (kD / 4);
    float4 qa[kH4], qb[kH4];
#pragma unroll
    for (int jj = 0; jj < kH4; ++jj) {
      qa[jj] = __ldg(qp + (jj * 32 + lane) * 2);
      qb[jj] = __ldg(qp + (jj * 32 + lane) * 2 + 1);
    }
    const long long page = it.g0 / P;
    const uint4* hp = reinterpret_cast<const uint4*>(half_slabs[page >> slab_shift] +
                                                    (size_t)((page & slab_mask) * P) * kD) + lane;
    WarpTopK<SLOTS> tk;
    tk.init(K, lane);
    int v = 0;
    for (; v + U <= it.len; v += U) {
      uint4 x[U][kH4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int jj = 0; jj < kH4; ++jj) x[u][jj] = __ldcs(hp + (size_t)(v + u) * (kD / 8) + jj * 32);
      float acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = 0.f;
#pragma unroll
        for (int jj = 0; jj < kH4; ++jj) acc[u] = dot8(x[u][jj], qa[jj], qb[jj], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = warp_sum16(acc[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long g = it.g0 + v + u;
        if (tk.may_enter(acc[u]) && tk.admits(acc[u], g)) tk.insert(acc[u], g);
      }
    }
    for (; v < it.len; ++v) {
      float acc = 0.f;
#pragma unroll
      for (int jj = 0; jj < kH4; ++jj) acc = dot8(__ldcs(hp + (size_t)v * (kD / 8) + jj * 32), qa[jj], qb[jj], acc);
      acc = warp_sum16(acc);
      const long long g = it.g0 + v;
      if (tk.may_enter(acc) && tk.admits(acc, g)) tk.insert(acc, g);
    }
    tk.store(part_s + (size_t)item * K, part_g + (size_t)item * K);
    pos = __shfl_sync(kFullMask, next, 0);
  }
}

// Stage 2 work items: candidate (q, r) -> one single-vector run of the fp32 codes.
__global__ void rescore_items_kernel(ListTable lt, int nq, int K, const long long* __restrict__ G,
                                     ScanItem* __restrict__ items, int* __restrict__ q_begin,
                                     int* __restrict__ n_items, int* __restrict__ queue_counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nq * K;
  if (i == 0) {
    *n_items = total;
    *queue_counter = 0;
  }
  if (i <= nq) q_begin[i] = i * K;
  if (i >= total) return;
  const long long g = G[i];
  ScanItem it;
  it.codes = nullptr;
  it.ids = nullptr;
  it.len = 0;
  it.q = i / K;
  it.pad = 0;
  if (g >= 0) {
    const int P = lt.page_vecs;
    const long long page = g / P;
    const int slot = (int)(g - page * P);
    const int slab = (int)(page >> lt.slab_shift);
    const long long in_slab = page & ((1ll << lt.slab_shift) - 1);
    it.codes = lt.code_slabs[slab] + ((size_t)in_slab * P + slot) * lt.d;
    it.ids = lt.id_slabs[slab] + (size_t)in_slab * P + slot;
    it.len = 1;
    it.pad = g;
  }
  items[i] = it;
}

// One warp per query.
__global__ void two_stage_check_kernel(int nq, int d, int k, int K, const float* __restrict__ Q,
                                       const float* __restrict__ D, const long long* __restrict__ I,
                                       const float* __restrict__ Dp, const long long* __restrict__ G,
                                       const float* __restrict__ maxima, unsigned char* __restrict__ flags,
                                       int* __restrict__ n_flagged) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = Q[(size_t)q * d + j];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum16(ss);
  if (lane != 0) return;
  bool proven;
  if (G[(size_t)q * K + K - 1] < 0) {
    proven = true;  // shortlist not full: it holds every probed vector
  } else if (I[(size_t)q * k + k - 1] < 0) {
    proven = false;  // cannot happen with K >= k; be safe
  } else {
    const float qn = __fmul_ru(__fsqrt_ru(ss), 1.0001f);
    const float r_max = __fmul_ru(maxima[0], 1.0001f), x_max = __fmul_ru(maxima[1], 1.0001f);
    const float round_slack = __fmul_ru(80.f * 5.9604645e-8f, __fadd_ru(x_max, r_max));
    const float B = __fmul_ru(qn, __fadd_ru(r_max, round_slack));
    const float a_min = Dp[(size_t)q * K + K - 1];
    const float T = D[(size_t)q * k + k - 1];
    proven = __fadd_ru(a_min, B) < T && B == B && a_min == a_min;  // NaN anywhere: not proven
  }
  flags[q] = proven ? 0 : 1;
  if (!proven) atomicAdd(n_flagged, 1);
}

template <typename Kern>
int resident16(Kern kern) {
  int n = 0;
  ABSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kScanThreads, 0));
  return n < 1 ? 1 : n;
}

}  // namespace

void launch_scan16(const Scan16Launch& a, cudaStream_t st) {
  ABSB_CHECK(a.K == 32 || a.K == 64 || a.K == 128, ABSB_ERR_INVALID, "shortlist length %d (32, 64 or 128)", a.K);
  ABSB_DISPATCH_SLOTS(a.K, {
    auto kern = ivf_scan16_kernel<SLOTS, kScanUnroll>;
    const int per_sm = a.ctas_per_sm > 0 ? a.ctas_per_sm : resident16(kern);
    kern<<<a.sm_count * per_sm, kScanThreads, 0, st>>>(a.Q, a.items, a.n_items, a.queue_counter, a.order, a.K,
                                                       a.part_s, a.part_g, a.half_slabs, a.slab_shift, a.page_vecs);
  });
  ABSB_CUDA(cudaGetLastError());
}

void launch_rescore_items(const ListTable& lt, int nq, int K, const long long* G, ScanItem* items, int* q_begin,
                          int* n_items, int* queue_counter, cudaStream_t st) {
  const int work = std::max(nq * K, nq + 1);
  rescore_items_kernel<<<(work + 255) / 256, 256, 0, st>>>(lt, nq, K, G, items, q_begin, n_items, queue_counter);
  ABSB_CUDA(cudaGetLastError());
}

void launch_two_stage_check(int nq, int d, int k, int K, const float* Q, const float* D, const long long* I,
                            const float* Dp, const long long* G, const float* maxima, unsigned char* flags,
                            int* n_flagged, cudaStream_t st) {
  two_stage_check_kernel<<<(nq + 3) / 4, 128, 0, st>>>(nq, d, k, K, Q, D, I, Dp, G, maxima, flags, n_flagged);
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb
