// ivf_scan_ring.cu — the fine scan of IndexIVFFlat::search with the list vectors STAGED IN SHARED MEMORY
// by the bulk-copy engine (faiss IVFFlatScanner::scan_codes + heap under IndexIVF::search_preassigned:
// /root/reference/Makefile:31-32, README.md:16,28).
//
// Why a second scan kernel: ivf_scan_kernel / ivf_scan16_kernel hold the vectors in flight in REGISTERS
// (2 x 4 KB per warp), so HBM bandwidth needs ~24 resident warps per SM and the kernel cannot share an SM
// with the persistent tcgen05 GEMM CTAs of the encoder (profiles/r01r_pipeline_experiment.md).  Here the
// bytes in flight live in shared memory:
//
//   * every warp owns a ring of `depth` stages of STAGE_VECS vectors (4-8 KB each) and is its own
//     producer: lane 0 issues ONE cp.async.bulk (TMA engine, no register staging) per stage, completion
//     on a per-stage mbarrier; the producer cursor runs depth-1 stages ahead of the consumer cursor and
//     ACROSS work-item boundaries (it pulls the next item off the queue while the current one is still
//     being consumed), so the ring never drains between items;
//   * the consumer half of the warp waits on the stage's mbarrier, reads the vectors with conflict-free
//     128-bit shared loads in exactly the lane -> element mapping of the register kernels (lane l owns
//     float4 j*32 + l, j = 0..7), runs the same 32-FMA chain per lane and the same xor butterfly, and
//     feeds the same WarpTopK — scores and results are bit-identical to ivf_scan_kernel /
//     ivf_scan16_kernel (tests compare them);
//   * 4 warps x 3 stages x 4 KB = 48 KB per CTA already keeps 32 KB per SM in flight from ONE small CTA
//     (128 threads, < 80 registers): it fits beside a GEMM CTA built with fewer pipeline stages, which is
//     what lets the HBM-bound scan of batch i run under the tensor-bound encode of batch i+1.  Alone on
//     the GPU it runs with deeper rings / more CTAs per SM.
//
// Work distribution, work items, partial results and the merge are unchanged (ivf_scan.cu).
#include <cuda_fp16.h>

#include "common.cuh"
#include "ivf_scan.cuh"
#include "tc.cuh"
#include "topk.cuh"

namespace absb {

namespace {

constexpr int kD = 1024;
constexpr int kFifo = 8;  // item indices the producer cursor may be ahead of the consumer (>= depth + 1)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// Same copy with an L2 eviction-priority hint (evict_first): list codes are streamed once per batch, so they
// should not displace the encoder's weights and activations from L2 when the scan shares the GPU with it.
__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                               uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float dot4r(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

__device__ __forceinline__ float dot8r(const uint4 x, const float4 qa, const float4 qb, float acc) {
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

// Transposed butterfly: reduces the per-lane partial sums of U = 1, 2 or 4 vectors at once.  Afterwards lane l
// holds the complete inner product of vector l / (32 / U) (every lane of that group the same value).  For each
// vector the additions form exactly the xor-butterfly tree of warp_sum (levels 16, 8, 4, 2, 1; fp32 addition is
// commutative), so the value is bit-identical to the register kernels' `acc += shfl_xor(acc, o)` chain — with
// 5 / 5 / 6 shuffles per stage instead of 5 per vector.
template <int U>
__device__ __forceinline__ float warp_sum_transposed(const float (&acc)[U], int lane) {
  float r;
  if (U == 1) {
    r = acc[0] + __shfl_xor_sync(kFullMask, acc[0], 16);
    r += __shfl_xor_sync(kFullMask, r, 8);
  } else if (U == 2) {
    const bool hi = lane & 16;
    const float keep = hi ? acc[1] : acc[0], send = hi ? acc[0] : acc[1];
    r = keep + __shfl_xor_sync(kFullMask, send, 16);
    r += __shfl_xor_sync(kFullMask, r, 8);
  } else {
    const bool hi = lane & 16;
    const float k0 = hi ? acc[2] : acc[0], s0 = hi ? acc[0] : acc[2];
    const float k1 = hi ? acc[3] : acc[1], s1 = hi ? acc[1] : acc[3];
    const float r0 = k0 + __shfl_xor_sync(kFullMask, s0, 16);
    const float r1 = k1 + __shfl_xor_sync(kFullMask, s1, 16);
    const bool hi8 = lane & 8;
    const float keep = hi8 ? r1 : r0, send = hi8 ? r0 : r1;
    r = keep + __shfl_xor_sync(kFullMask, send, 8);
  }
  r += __shfl_xor_sync(kFullMask, r, 4);
  r += __shfl_xor_sync(kFullMask, r, 2);
  r += __shfl_xor_sync(kFullMask, r, 1);
  return r;
}

struct RingItem {
  const char* src;       // first byte of the item's codes (fp32 codes or fp16 shadow codes)
  const long long* ids;  // fp32 scan only
  long long g0;          // global slot number of the first vector (fp16 scan: candidate ids)
  int len, q;
};

template <bool HALF>
__device__ __forceinline__ RingItem load_ring_item(const ScanItem* p, const unsigned short* const* half_slabs,
                                                   int slab_shift, int P) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  RingItem it;
  it.len = b.x;
  it.q = b.y;
  it.g0 = (long long)(((unsigned long long)(unsigned)b.w << 32) | (unsigned)b.z);
  it.ids = reinterpret_cast<const long long*>(((unsigned long long)(unsigned)a.w << 32) | (unsigned)a.z);
  if (HALF) {
    const long long page = it.g0 / P;
    const long long in_slab = page & ((1ll << slab_shift) - 1);
    // the item starts at a page boundary or inside a page: slot offset = g0 - page * P
    it.src = reinterpret_cast<const char*>(half_slabs[page >> slab_shift]) +
             ((size_t)in_slab * P + (size_t)(it.g0 - page * P)) * (kD * 2);
  } else {
    it.src = reinterpret_cast<const char*>(((unsigned long long)(unsigned)a.y << 32) | (unsigned)a.x);
  }
  return it;
}

// HALF = false: fp32 codes, k results per item with the vectors' int64 ids   (ivf_scan_kernel's contract)
// HALF = true : fp16 shadow codes, K approximate candidates per item as global slot numbers (ivf_scan16_kernel's)
template <int SLOTS, bool HALF, int SV>
__device__ __forceinline__ void ivf_scan_ring_body(const float* __restrict__ Q, const ScanItem* __restrict__ items,
                                                   const int* __restrict__ n_items_ptr, int* __restrict__ queue_counter,
                                                   const int* __restrict__ order, int k, float* __restrict__ part_s,
                                                   long long* __restrict__ part_id,
                                                   const unsigned short* const* __restrict__ half_slabs, int slab_shift,
                                                   int P, int depth_and_hint) {
  const int depth = depth_and_hint & 0xff;
  const bool l2_evict_first = (depth_and_hint >> 8) & 1;
  uint64_t l2_policy = 0;
  if (l2_evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy));
  constexpr int VB = HALF ? kD * 2 : kD * 4;  // bytes per vector
  constexpr int SB = SV * VB;                 // bytes per stage
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* my_ring = ring_smem + (size_t)warp * depth * SB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_smem + (size_t)nwarps * depth * SB) + warp * depth;
  int* fifo = reinterpret_cast<int*>(ring_smem + (size_t)nwarps * depth * SB + (size_t)nwarps * depth * 8) + warp * kFifo;

  if (lane == 0) {
    for (int s = 0; s < depth; ++s) tc::mbar_init(bars + s, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();

  const int n_items = *n_items_ptr;

  // ---- producer cursor (warp-uniform registers; lane 0 issues) ----
  const char* p_src = nullptr;
  int p_left = 0;       // vectors of the producer's current item not yet requested
  bool p_done = false;  // queue exhausted
  int fifo_head = 0, fifo_tail = 0;
  int in_flight = 0;  // stages requested and not yet consumed

  auto produce = [&](int s) {
    while (p_left == 0) {
      if (p_done) return;
      int pos = 0;
      if (lane == 0) pos = atomicAdd(queue_counter, 1);
      pos = __shfl_sync(kFullMask, pos, 0);
      if (pos >= n_items) {
        p_done = true;
        return;
      }
      const int item = order ? __ldg(order + pos) : pos;
      const RingItem it = load_ring_item<HALF>(items + item, half_slabs, slab_shift, P);
      if (it.len <= 0) {
        // nothing to stream (a shortlist slot without a candidate): its partial result is all sentinels
        for (int r = lane; r < k; r += 32) {
          part_s[(size_t)item * k + r] = -INFINITY;
          part_id[(size_t)item * k + r] = kIdSentinel;
        }
        continue;
      }
      if (lane == 0) fifo[fifo_head & (kFifo - 1)] = item;
      ++fifo_head;
      p_src = it.src;
      p_left = it.len;
    }
    const int nv = p_left < SV ? p_left : SV;
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(bars + s, (uint32_t)(nv * VB));
      if (l2_evict_first) bulk_load_hint(my_ring + (size_t)s * SB, p_src, (uint32_t)(nv * VB), bars + s, l2_policy);
      else bulk_load(my_ring + (size_t)s * SB, p_src, (uint32_t)(nv * VB), bars + s);
    }
    p_src += (size_t)nv * VB;
    p_left -= nv;
    ++in_flight;
  };

  for (int s = 0; s < depth; ++s) produce(s);
  __syncwarp();  // fifo entries written by lane 0 are visible to the warp

  // ---- consumer ----
  int stage = 0;
  uint32_t parity = 0;
  int c_left = 0, c_v = 0, c_item = 0;
  RingItem cit{};
  float4 qv[8];
  WarpTopK<SLOTS> tk;
  tk.init(k, lane);
  // Software pipeline inside the warp: the per-lane partial sums of a stage are reduced (5-6 dependent
  // shuffles) and offered to the top-k while the NEXT stage's shared-memory loads and FMA chains issue — one
  // basic block holds both, so the shuffle latency hides behind FMA issue instead of adding to it.
  float pend[SV];
#pragma unroll
  for (int u = 0; u < SV; ++u) pend[u] = 0.f;
  int pend_nv = 0, pend_v = 0;
  const int my_u = SV == 1 ? 0 : (SV == 2 ? lane >> 4 : lane >> 3);
  // lane l holds the score of vector u(l) = l / (32 / SV) of the stage that started at vector v0 of the item;
  // one ballot finds the (rare) candidates, offered in list order
  auto offer = [&](float sc, int nv_, int v0) {
    unsigned m = __ballot_sync(kFullMask, my_u < nv_ && tk.may_enter(sc));
    while (m) {
      const int src = __ffs(m) - 1;  // first lane of the lowest candidate vector: ascending u = list order
      const int u = SV == 1 ? 0 : (SV == 2 ? src >> 4 : src >> 3);
      m &= ~(SV == 1 ? 0xffffffffu : (SV == 2 ? 0xffffu << (u * 16) : 0xffu << (u * 8)));
      const float cs = __shfl_sync(kFullMask, sc, src);
      if (HALF) {
        const long long g = cit.g0 + v0 + u;
        if (tk.admits(cs, g)) tk.insert(cs, g);
      } else if (tk.may_enter(cs)) {  // the threshold may have risen since the ballot
        const long long id = __ldg(cit.ids + v0 + u);
        if (tk.admits(cs, id)) tk.insert(cs, id);
      }
    }
  };
  while (in_flight > 0) {
    if (c_left == 0) {  // (nothing is pending here: an item's last stage is finished before the next item starts)
      c_item = fifo[fifo_tail & (kFifo - 1)];
      ++fifo_tail;
      cit = load_ring_item<HALF>(items + c_item, half_slabs, slab_shift, P);
      c_left = cit.len;
      c_v = 0;
      const float4* qp = reinterpret_cast<const float4*>(Q) + (size_t)cit.q * (kD / 4);
      if (HALF) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          qv[2 * jj] = __ldg(qp + (jj * 32 + lane) * 2);
          qv[2 * jj + 1] = __ldg(qp + (jj * 32 + lane) * 2 + 1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) qv[j] = __ldg(qp + j * 32 + lane);
      }
      tk.init(k, lane);
    }
    const int nv = c_left < SV ? c_left : SV;
    tc::mbar_wait(bars + stage, parity);
    const unsigned char* sp = my_ring + (size_t)stage * SB;
    float acc[SV];
    uint4 x[SV][HALF ? 4 : 8];
#pragma unroll
    for (int u = 0; u < SV; ++u)
#pragma unroll
      for (int j = 0; j < (HALF ? 4 : 8); ++j)
        if (u < nv) x[u][j] = *reinterpret_cast<const uint4*>(sp + (size_t)u * VB + (size_t)(j * 32 + lane) * 16);
    // the stage's bytes are on their way to registers: hand the buffer back to the bulk-copy engine.  The
    // refill is issued by lane 0 after the warp barrier, i.e. after every lane's shared-memory reads of this
    // stage have been performed (the ordering an mbarrier-based consumer release of a TMA pipeline relies on).
    __syncwarp();
    --in_flight;
    produce(stage);
    __syncwarp();
    // previous stage: reduction (shuffle chain) — independent of this stage's FMA chains below
    const float sc_prev = warp_sum_transposed<SV>(pend, lane);
#pragma unroll
    for (int u = 0; u < SV; ++u) {
      acc[u] = 0.f;
      if (u < nv) {
        if (HALF) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[u] = dot8r(x[u][jj], qv[2 * jj], qv[2 * jj + 1], acc[u]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[u] = dot4r(*reinterpret_cast<const float4*>(&x[u][j]), qv[j], acc[u]);
        }
      }
    }
    offer(sc_prev, pend_nv, pend_v);
#pragma unroll
    for (int u = 0; u < SV; ++u) pend[u] = acc[u];
    pend_nv = nv;
    pend_v = c_v;
    c_v += nv;
    c_left -= nv;
    if (c_left == 0) {  // the item ends with this stage: finish it now
      offer(warp_sum_transposed<SV>(pend, lane), pend_nv, pend_v);
      pend_nv = 0;
      tk.store(part_s + (size_t)c_item * k, part_id + (size_t)c_item * k);
    }
    if (++stage == depth) {
      stage = 0;
      parity ^= 1;
    }
  }
}

#define ABSB_RING_ARGS                                                                                            \
  const float *__restrict__ Q, const ScanItem *__restrict__ items, const int *__restrict__ n_items_ptr,           \
      int *__restrict__ queue_counter, const int *__restrict__ order, int k, float *__restrict__ part_s,          \
      long long *__restrict__ part_id, const unsigned short *const *__restrict__ half_slabs, int slab_shift, int P, \
      int depth_and_hint

// Full-size variant: whatever registers the compiler wants (120-165), several CTAs per SM when alone on the GPU.
template <int SLOTS, bool HALF, int SV>
__global__ void ivf_scan_ring_kernel(ABSB_RING_ARGS) {
  ivf_scan_ring_body<SLOTS, HALF, SV>(Q, items, n_items_ptr, queue_counter, order, k, part_s, part_id, half_slabs,
                                      slab_shift, P, depth_and_hint);
}

// Co-resident variant: capped at 96 registers so that ONE CTA of 8 warps (24,576 registers, 64 KB of ring) fits
// next to a tcgen05 GEMM CTA of the encoder (384 threads x 104 registers, <= 161 KB) on the same SM.
template <int SLOTS, bool HALF, int SV>
__global__ __maxnreg__(96) void ivf_scan_ring_small_kernel(ABSB_RING_ARGS) {
  ivf_scan_ring_body<SLOTS, HALF, SV>(Q, items, n_items_ptr, queue_counter, order, k, part_s, part_id, half_slabs,
                                      slab_shift, P, depth_and_hint);
}

template <typename Kern>
void launch_ring(Kern kern, const ScanRing& r, int sm_count, int ctas_per_sm, int stage_bytes, cudaStream_t st,
                 const float* Q, const ScanItem* items, const int* n_items, int* queue_counter, const int* order, int k,
                 float* part_s, long long* part_id, const unsigned short* const* half_slabs, int slab_shift, int P) {
  const int warps = r.warps, depth = r.depth;
  ABSB_CHECK(warps >= 1 && warps <= 16 && depth >= 2 && depth + 1 <= kFifo, ABSB_ERR_INVALID,
             "scan ring: %d warps x %d stages (1-16 warps, 2-%d stages)", warps, depth, kFifo - 1);
  const size_t smem = (size_t)warps * depth * stage_bytes + (size_t)warps * depth * 8 + (size_t)warps * kFifo * 4;
  ABSB_CHECK(smem <= 227 * 1024, ABSB_ERR_INVALID, "scan ring of %zu bytes exceeds the shared memory of an SM", smem);
  ABSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prefer_max_shared(kern);
  int per_sm = ctas_per_sm;
  if (per_sm <= 0) {
    ABSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  kern<<<sm_count * per_sm, warps * 32, smem, st>>>(Q, items, n_items, queue_counter, order, k, part_s, part_id,
                                                    half_slabs, slab_shift, P, depth | (r.l2_evict_first ? 0x100 : 0));
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace

void launch_scan_ring(const ScanLaunch& a, const ScanRing& r, cudaStream_t st) {
  ABSB_CHECK(a.d == kD, ABSB_ERR_UNSUPPORTED, "the shared-memory ring scan is built for d = %d (d=%d)", kD, a.d);
  ABSB_DISPATCH_SLOTS(a.k, {
    if (r.small)
      launch_ring(ivf_scan_ring_small_kernel<SLOTS, false, 1>, r, a.sm_count, a.ctas_per_sm, 1 * kD * 4, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.k, a.part_s, a.part_id, nullptr, 0, 1);
    else if (r.stage_vecs == 1)
      launch_ring(ivf_scan_ring_kernel<SLOTS, false, 1>, r, a.sm_count, a.ctas_per_sm, 1 * kD * 4, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.k, a.part_s, a.part_id, nullptr, 0, 1);
    else
      launch_ring(ivf_scan_ring_kernel<SLOTS, false, 2>, r, a.sm_count, a.ctas_per_sm, 2 * kD * 4, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.k, a.part_s, a.part_id, nullptr, 0, 1);
  });
}

void launch_scan16_ring(const Scan16Launch& a, const ScanRing& r, cudaStream_t st) {
  ABSB_CHECK(a.K == 32 || a.K == 64 || a.K == 128, ABSB_ERR_INVALID, "shortlist length %d (32, 64 or 128)", a.K);
  ABSB_DISPATCH_SLOTS(a.K, {
    if (r.small)
      launch_ring(ivf_scan_ring_small_kernel<SLOTS, true, 2>, r, a.sm_count, a.ctas_per_sm, 2 * kD * 2, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.K, a.part_s, a.part_g, a.half_slabs, a.slab_shift, a.page_vecs);
    else if (r.stage_vecs <= 2)
      launch_ring(ivf_scan_ring_kernel<SLOTS, true, 2>, r, a.sm_count, a.ctas_per_sm, 2 * kD * 2, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.K, a.part_s, a.part_g, a.half_slabs, a.slab_shift, a.page_vecs);
    else
      launch_ring(ivf_scan_ring_kernel<SLOTS, true, 4>, r, a.sm_count, a.ctas_per_sm, 4 * kD * 2, st, a.Q, a.items,
                  a.n_items, a.queue_counter, a.order, a.K, a.part_s, a.part_g, a.half_slabs, a.slab_shift, a.page_vecs);
  });
}

}  // namespace absb
