// dense.cu — exact-fp32 dense pieces of the search path:
//   * gemm_nt_f32:  S[M,N] = A[M,K] · B[N,K]^T  (both operands K-major, like faiss's sgemm call in
//     IndexFlat / the IVF coarse quantiser — reached from Index.search, Makefile:31-32)
//   * select_rows:  per-row top-k of a score matrix (quantizer.search / IndexFlatIP::search)
//   * merge_partials: per-query merge of partial k-best lists into the final (D, I)
//
// The fp32 FFMA GEMM is the ranking-faithful baseline implementation of the coarse step
// (coarse_impl = 0).  The tcgen05 split-bf16 GEMM in gemm_tc.cu replaces it when enabled.
#include "common.cuh"
#include "peer.cuh"
#include "topk.cuh"

namespace absb {

// ------------------------------------------------------------------------------------------
// fp32 NT GEMM, 128x128x16 tiles, 256 threads, 8x8 micro-tile per thread, register prefetch.
// ------------------------------------------------------------------------------------------
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__global__ __launch_bounds__(256) void gemm_nt_f32_kernel(int M, int N, int K,
                                                          const float* __restrict__ A, int lda,
                                                          const float* __restrict__ B, int ldb,
                                                          float* __restrict__ C, int ldc) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // global->smem mapping: 128 rows x 16 k = 512 float4; thread loads rows r and r+64 at k-quad kq
  const int lr = tid / 4, kq = (tid % 4) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      const int k = k0 + kq;
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k < K) ra[h] = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k);
      if (n0 + r < N && k < K) rb[h] = *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * ldb + k);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[buf][kq + 0][r] = ra[h].x; As[buf][kq + 1][r] = ra[h].y;
      As[buf][kq + 2][r] = ra[h].z; As[buf][kq + 3][r] = ra[h].w;
      Bs[buf][kq + 0][r] = rb[h].x; Bs[buf][kq + 1][r] = rb[h].y;
      Bs[buf][kq + 2][r] = rb[h].z; Bs[buf][kq + 3][r] = rb[h].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float* c = C + (size_t)m * ldc + n;
      if (n + 3 < N && (ldc % 4 == 0)) {
        *reinterpret_cast<float4*>(c) = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) c[j] = acc[i][jh * 4 + j];
      }
    }
  }
}

}  // namespace

void gemm_nt_f32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                 int ldc, cudaStream_t st) {
  if (M == 0 || N == 0) return;
  ABSB_CHECK(K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, ABSB_ERR_INVALID,
             "gemm_nt_f32 needs K, lda, ldb multiples of 4 (K=%d)", K);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_nt_f32_kernel<<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc);
  ABSB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// Row-wise top-k of a score matrix.  One CTA (8 warps) per row; every warp keeps its own
// register-resident k-best over a strided slice, warp 0 merges the 8 lists through shared memory.
// ids are column + id_offset.  Output is rank-major; sentinels are kept unless `finalize`.
// ------------------------------------------------------------------------------------------
namespace {

constexpr int kSelectWarps = 8;

template <int SLOTS, bool VEC>
__global__ __launch_bounds__(kSelectWarps * 32) void select_rows_kernel(
    const float* __restrict__ S, int64_t ld, int ncols, long long id_offset, int k,
    float* __restrict__ out_s, long long* __restrict__ out_id, int64_t out_ld, int finalize) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm_s = reinterpret_cast<float*>(smem_raw);
  long long* sm_id = reinterpret_cast<long long*>(smem_raw + sizeof(float) * kSelectWarps * ((k + 1) & ~1));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = blockIdx.x;
  const float* srow = S + row * ld;

  WarpTopK<SLOTS> tk;
  tk.init(k, lane);
  if (VEC) {
    // 128-bit loads, four columns per lane; one warp vote rejects the whole 128-column group when no
    // score reaches the current k-th best (the common case once the threshold has settled).  The
    // container's total order (score desc, id asc) makes the result independent of the offer order.
    const int ngroups = ncols / 128;
    // four 128-column groups per step: four independent 128-bit loads in flight per lane and ONE vote for all
    // of them (once the threshold has settled almost every step is rejected by that vote)
    int g = warp;
    for (; g + 3 * kSelectWarps < ngroups; g += 4 * kSelectWarps) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(srow + (g + u * kSelectWarps) * 128 + lane * 4));
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
      if (__any_sync(kFullMask, tk.may_enter(mx))) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = (g + u * kSelectWarps) * 128 + lane * 4;
          const float m1 = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
          if (__any_sync(kFullMask, tk.may_enter(m1))) {
            tk.offer_lanes(v[u].x, id_offset + c, true);
            tk.offer_lanes(v[u].y, id_offset + c + 1, true);
            tk.offer_lanes(v[u].z, id_offset + c + 2, true);
            tk.offer_lanes(v[u].w, id_offset + c + 3, true);
          }
        }
      }
    }
    for (; g < ngroups; g += kSelectWarps) {
      const int c = g * 128 + lane * 4;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(srow + c));
      const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (__any_sync(kFullMask, tk.may_enter(mx))) {
        tk.offer_lanes(v.x, id_offset + c, true);
        tk.offer_lanes(v.y, id_offset + c + 1, true);
        tk.offer_lanes(v.z, id_offset + c + 2, true);
        tk.offer_lanes(v.w, id_offset + c + 3, true);
      }
    }
    for (int c0 = ngroups * 128 + warp * 32; c0 < ncols; c0 += kSelectWarps * 32) {
      const int c = c0 + lane;
      const bool valid = c < ncols;
      const float v = valid ? __ldcs(srow + c) : 0.f;
      tk.offer_lanes(v, id_offset + c, valid);
    }
  } else {
    for (int c0 = warp * 32; c0 < ncols; c0 += kSelectWarps * 32) {
      const int c = c0 + lane;
      const bool valid = c < ncols;
      const float v = valid ? __ldcs(srow + c) : 0.f;
      tk.offer_lanes(v, id_offset + c, valid);
    }
  }
  const int kk = (k + 1) & ~1;
  tk.store(sm_s + warp * kk, sm_id + warp * k);
  __syncthreads();
  if (warp == 0) {
    // merge the other warps' lists into warp 0's container
    for (int w = 1; w < kSelectWarps; ++w)
      for (int r0 = 0; r0 < k; r0 += 32) {
        const int r = r0 + lane;
        const bool valid = r < k;
        const float v = valid ? sm_s[w * kk + r] : 0.f;
        const long long id = valid ? sm_id[w * k + r] : 0;
        tk.offer_lanes(v, id, valid && id != kIdSentinel);
      }
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const int r = lane * SLOTS + i;
      if (r < k) {
        float s = tk.s[i];
        long long id = tk.id[i];
        if (finalize && id == kIdSentinel) { s = -3.4028234663852886e38f; id = -1; }
        out_s[row * out_ld + r] = s;
        out_id[row * out_ld + r] = id;
      }
    }
  }
}

// Offers the candidates [b, e) of a query to the warp's k-best in their stored order, 32 per vote; the loads of four
// votes are issued together so the merge waits for memory once per 128 candidates instead of once per 32.
template <int SLOTS>
__device__ __forceinline__ void offer_span(WarpTopK<SLOTS>& tk, const float* __restrict__ part_s,
                                           const long long* __restrict__ part_id, int64_t b, int64_t e, int lane) {
  constexpr int U = 4;
  for (int64_t c0 = b; c0 < e; c0 += 32 * U) {
    float v[U];
    long long id[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = c0 + 32 * u + lane;
      const bool valid = c < e;
      v[u] = valid ? part_s[c] : 0.f;
      id[u] = valid ? part_id[c] : kIdSentinel;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + 32 * u >= e) break;  // warp-uniform
      tk.offer_lanes(v[u], id[u], id[u] != kIdSentinel);
    }
  }
}

// One warp per query merges `cnt` blocks of k candidates laid out contiguously.
template <int SLOTS>
__global__ __launch_bounds__(128) void merge_partials_kernel(
    int nq, int k, const int* __restrict__ q_begin, const float* __restrict__ part_s,
    const long long* __restrict__ part_id, float* __restrict__ D, long long* __restrict__ I,
    const unsigned char* __restrict__ active) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  if (active != nullptr && active[q] == 0) return;
  WarpTopK<SLOTS> tk;
  tk.init(k, lane);
  offer_span(tk, part_s, part_id, (int64_t)q_begin[q] * k, (int64_t)q_begin[q + 1] * k, lane);
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    const int r = lane * SLOTS + i;
    if (r < k) {
      float s = tk.s[i];
      long long id = tk.id[i];
      if (id == kIdSentinel) { s = -3.4028234663852886e38f; id = -1; }
      D[(int64_t)q * k + r] = s;
      I[(int64_t)q * k + r] = id;
    }
  }
}

// merge_partials fused with the shard exchange (peer.cuh): the merged k-best of every query is
// stored straight into slot [rank] of EVERY rank's ring entry over NVLink instead of a local
// (D, I); the last CTA of the launch that completes the record raises this rank's flag everywhere.
// With `active` (the two-stage scan's fallback flags) only flagged queries are merged from the
// partials; the others push the row the earlier stages left in (Dbase, Ibase).
template <int SLOTS>
__global__ __launch_bounds__(128) void merge_partials_push_kernel(
    int nq, int k, const int* __restrict__ q_begin, const float* __restrict__ part_s,
    const long long* __restrict__ part_id, PeerPush pp, const unsigned char* __restrict__ active,
    const float* __restrict__ Dbase, const long long* __restrict__ Ibase) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q < nq) {
    WarpTopK<SLOTS> tk;
    tk.init(k, lane);
    const bool merge = active == nullptr || active[q] != 0;
    if (merge) {
      offer_span(tk, part_s, part_id, (int64_t)q_begin[q] * k, (int64_t)q_begin[q + 1] * k, lane);
    }
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const int r = lane * SLOTS + i;
      if (r < k) {
        float s = tk.s[i];
        long long id = tk.id[i];
        if (merge) {
          if (id == kIdSentinel) { s = -3.4028234663852886e38f; id = -1; }
        } else {
          s = Dbase[(int64_t)q * k + r];
          id = Ibase[(int64_t)q * k + r];
        }
        const int64_t o = (pp.q_off + q) * k + r;
        for (int w = 0; w < pp.world; ++w) {
          char* rec = pp.slot_ptrs[w];
          reinterpret_cast<long long*>(rec + pp.i_off)[o] = id;
          reinterpret_cast<float*>(rec + pp.d_off)[o] = s;
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(pp.done_counter, 1u) == gridDim.x - 1) {
      *pp.done_counter = 0;
      if (pp.epoch) {
        __threadfence_system();
        for (int w = 0; w < pp.world; ++w)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flag_ptrs[w]), "l"(pp.epoch) : "memory");
      }
    }
  }
}

// Shard merge: candidates of query q live at [w, q, :] for w < world.
template <int SLOTS>
__global__ __launch_bounds__(128) void merge_shards_kernel(int world, int64_t nq, int k,
                                                           const float* __restrict__ D_all,
                                                           const long long* __restrict__ I_all,
                                                           int64_t d_stride /* bytes per rank */,
                                                           int64_t i_stride /* bytes per rank */,
                                                           float* __restrict__ D,
                                                           long long* __restrict__ I,
                                                           const unsigned long long* __restrict__ flags,
                                                           unsigned long long epoch, int* __restrict__ status) {
  __shared__ int timed_out;
  if (flags != nullptr) {
    // fused with the peer exchange: acquire the arrival flag of every rank before reading its record
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if ((int)threadIdx.x < world) {
      unsigned long long t0, t1, v;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) {  // a dead peer must not hang the GPU
          *reinterpret_cast<volatile int*>(status) = 1;  // mapped host memory: the host sees it without a sync
          __threadfence_system();
          timed_out = 1;
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
    if (timed_out) {
      // never hand back a merge of stale records: every slot reads "missing" (faiss: -FLT_MAX, -1)
      const int64_t q0 = (int64_t)blockIdx.x * (blockDim.x >> 5);
      for (int64_t i = threadIdx.x; i < (int64_t)(blockDim.x >> 5) * k; i += blockDim.x) {
        const int64_t o = q0 * k + i;
        if (o < nq * k) {
          D[o] = -3.4028234663852886e38f;
          I[o] = -1;
        }
      }
      return;
    }
  }
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  WarpTopK<SLOTS> tk;
  tk.init(k, lane);
  for (int w = 0; w < world; ++w) {
    const float* Dw = reinterpret_cast<const float*>(reinterpret_cast<const char*>(D_all) + w * d_stride);
    const long long* Iw = reinterpret_cast<const long long*>(reinterpret_cast<const char*>(I_all) + w * i_stride);
    const int64_t base = q * k;
    for (int r0 = 0; r0 < k; r0 += 32) {
      const int r = r0 + lane;
      const bool valid = r < k;
      const float v = valid ? Dw[base + r] : 0.f;
      const long long id = valid ? Iw[base + r] : -1;
      tk.offer_lanes(v, id, valid && id >= 0);
    }
  }
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    const int r = lane * SLOTS + i;
    if (r < k) {
      float s = tk.s[i];
      long long id = tk.id[i];
      if (id == kIdSentinel) { s = -3.4028234663852886e38f; id = -1; }
      D[q * k + r] = s;
      I[q * k + r] = id;
    }
  }
}

}  // namespace

void select_rows(const float* S, int64_t ld, int64_t nrows, int ncols, long long id_offset, int k,
                 float* out_s, long long* out_id, int64_t out_ld, bool finalize, cudaStream_t st) {
  if (nrows == 0) return;
  ABSB_CHECK(k >= 1 && k <= ABSB_MAX_K, ABSB_ERR_INVALID, "k=%d outside [1,%d]", k, ABSB_MAX_K);
  const size_t smem = sizeof(float) * kSelectWarps * ((k + 1) & ~1) + sizeof(long long) * kSelectWarps * k;
  const bool vec = ld % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0;
  if (vec) {
    ABSB_DISPATCH_SLOTS(k, (select_rows_kernel<SLOTS, true><<<(unsigned)nrows, kSelectWarps * 32, smem, st>>>(
                               S, ld, ncols, id_offset, k, out_s, out_id, out_ld, finalize ? 1 : 0)));
  } else {
    ABSB_DISPATCH_SLOTS(k, (select_rows_kernel<SLOTS, false><<<(unsigned)nrows, kSelectWarps * 32, smem, st>>>(
                               S, ld, ncols, id_offset, k, out_s, out_id, out_ld, finalize ? 1 : 0)));
  }
  ABSB_CUDA(cudaGetLastError());
}

void merge_partials(int nq, int k, const int* q_begin, const float* part_s, const long long* part_id,
                    float* D, long long* I, cudaStream_t st, const unsigned char* active) {
  if (nq == 0) return;
  ABSB_DISPATCH_SLOTS(k, (merge_partials_kernel<SLOTS><<<(nq + 3) / 4, 128, 0, st>>>(
                             nq, k, q_begin, part_s, part_id, D, I, active)));
  ABSB_CUDA(cudaGetLastError());
}

void merge_shards(int world, int64_t nq, int k, const float* D_all, const long long* I_all,
                  int64_t d_stride, int64_t i_stride, float* D, long long* I, cudaStream_t st) {
  if (nq == 0) return;
  ABSB_DISPATCH_SLOTS(k, (merge_shards_kernel<SLOTS><<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(
                             world, nq, k, D_all, I_all, d_stride, i_stride, D, I, nullptr, 0ull, nullptr)));
  ABSB_CUDA(cudaGetLastError());
}

void merge_partials_push(int nq, int k, const int* q_begin, const float* part_s, const long long* part_id,
                         const PeerPush& pp, cudaStream_t st, const unsigned char* active, const float* Dbase,
                         const long long* Ibase) {
  if (nq == 0) return;
  ABSB_CHECK(active == nullptr || (Dbase && Ibase), ABSB_ERR_INVALID, "masked push needs the base rows");
  ABSB_DISPATCH_SLOTS(k, (merge_partials_push_kernel<SLOTS><<<(nq + 3) / 4, 128, 0, st>>>(nq, k, q_begin, part_s,
                                                                                         part_id, pp, active, Dbase, Ibase)));
  ABSB_CUDA(cudaGetLastError());
}

void merge_shards_wait(int world, int64_t nq, int k, const char* entry, int64_t slot_bytes, int64_t i_off,
                       int64_t d_off, const unsigned long long* flags, unsigned long long epoch, int* status,
                       float* D, long long* I, cudaStream_t st) {
  if (nq == 0) return;
  ABSB_CHECK(world <= 128, ABSB_ERR_INVALID, "world=%d", world);
  ABSB_DISPATCH_SLOTS(k, (merge_shards_kernel<SLOTS><<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(
                             world, nq, k, reinterpret_cast<const float*>(entry + d_off),
                             reinterpret_cast<const long long*>(entry + i_off), slot_bytes, slot_bytes, D, I, flags,
                             epoch, status)));
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb
