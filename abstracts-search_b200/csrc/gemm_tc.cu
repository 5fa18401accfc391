// gemm_tc.cu — bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05), the one
// dense contraction on the abstracts-search path:
//   * every Linear of the stella/Qwen2 encoder behind SentenceTransformer.encode()
//     (/root/reference/Makefile:65, README.md:28,60 — SURVEY §2c E3/E6/E7/E8), with the bias,
//     SwiGLU and residual-add epilogues fused;
//   * the IVF coarse quantiser  <q, centroid>  (faiss quantizer.search under Index.search/add,
//     /root/reference/Makefile:25,32) as a split-bf16 GEMM that keeps fp32-faithful scores.
//
// C[M,N] = A[M,K] * B[N,K]^T, both operands K-major (row-major activations, nn.Linear weights).
//
// Kernel anatomy (persistent, one CTA per SM, 256 threads; NCTA = 2 pairs the two SMs of a TPC on one
// 256 x BN tile with cta_group::2 so that each SM stages only half of the B tile):
//   warp 0   TMA producer: cp.async.bulk.tensor 128x64 (A) and (BN/NCTA)x64 (B) bf16 boxes,
//            SWIZZLE_128B, into a 4-7 stage shared-memory ring; completion bytes of both CTAs of a
//            pair are credited to the leader's `full` mbarrier;
//   warp 1   MMA issuer (leader CTA): one thread issues tcgen05.mma.cta_group::{1,2}.kind::f16
//            (M=128*NCTA, N=BN, K=16) x4 per stage into a TMEM accumulator, tcgen05.commit releases
//            the stage (`empty`, multicast to both CTAs) and finally signals `tmem_full`;
//   warp 2   allocates / frees the 512 TMEM columns (double-buffered accumulator);
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused epilogue straight from
//            registers (each thread owns one output row: 64-128 contiguous bytes per burst; the
//            residual add goes through a swizzled shared-memory box and ONE cp.reduce.async.bulk.tensor
//            per 32 x 32 block), then release the accumulator (`tmem_empty`,
//            remote arrive from the peer CTA) so the next tile's MMAs overlap.
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc.cuh"

namespace absb {

namespace {

constexpr int BM = 128;  // rows of A per CTA (UMMA M = 128 * NCTA)
constexpr int BK = 64;   // bf16 elements = 128 bytes = one swizzle span
constexpr int kThreads = 384;  // TMA, MMA, TMEM-alloc, spare + 8 epilogue warps
constexpr int kEpiWarp0 = 4;
constexpr int kTmemCols = 512;   // two accumulator stages at column 0 and 256
constexpr int kAccStride = 256;
// Residual-add epilogue: false = red.global.add.v4.f32 straight from the tcgen05.ld registers, true = bulk
// tensor reduction (cp.reduce.async.bulk.tensor) of 32 x 32 boxes staged in shared memory.  Measured on
// B200 (tools/gemm_bench.py, T = 16384): O-proj 99 us (red) vs 91 us (bulk reduction), FFN-down 356 vs 364 us;
// a second staging buffer per warp changed nothing.
#ifndef ABSB_TMA_REDUCE
#define ABSB_TMA_REDUCE 1
#endif
constexpr bool kTmaReduce = ABSB_TMA_REDUCE != 0;

template <int BN, int NCTA, int EPI = EPI_BF16_BIAS>
struct SmemLayout {
  static constexpr int kBRows = BN / NCTA;  // B rows this CTA stages (a pair splits the N tile)
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // residual-add epilogue: one 32 x 32 fp32 box (4 KB, SWIZZLE_128B) per epilogue warp, the source of
  // its cp.reduce.async.bulk.tensor
  static constexpr int kEpiBytes = (EPI == EPI_F32_ADD && kTmaReduce) ? 8 * 4096 : 0;
  // The ring depth is a LAUNCH parameter (KernelParams::stages): kMaxStages fills the SM when the GEMM has it
  // to itself; a smaller budget (absb_gemm_set_smem_budget) leaves room for a co-resident scan CTA.
  static constexpr int kMaxStages = (216 * 1024 - kEpiBytes) / kStageBytes;
  static constexpr int epi_off(int stages) { return stages * kStageBytes; }
  static constexpr int bar_off(int stages) { return epi_off(stages) + kEpiBytes; }
  // full[stages], empty[stages], tmem_full[2], tmem_empty[2], tmem ptr
  static constexpr int total(int stages) { return bar_off(stages) + (2 * stages + 4) * 8 + 16; }
  static constexpr int dyn(int stages) { return total(stages) + 1024; }  // slack for manual 1024-byte alignment
  static int stages_for(int budget_bytes) {
    int s = kMaxStages;
    while (s > 2 && dyn(s) > budget_bytes) --s;
    return s;
  }
  static_assert(kBBytes % 1024 == 0, "SWIZZLE_128B tiles must stay 1024-byte aligned");
  static_assert(dyn(kMaxStages) <= 227 * 1024, "shared memory budget");
};

int g_gemm_smem_budget = 227 * 1024;
inline int gemm_smem_budget() { return g_gemm_smem_budget; }

struct KernelParams {
  int M, N, K;
  int tiles_m, tiles_n;  // tiles_m counts (128 * NCTA)-row blocks
  int n_fast;            // raster: 1 = consecutive tiles walk N (A tile shared), 0 = walk M (B tile shared)
  int stages;            // depth of the shared-memory operand ring (<= SmemLayout::kMaxStages)
  int num_kb;            // k-blocks per tile over all segments
  int seg_kb;            // k-blocks per segment
  int a_off[kMaxGemmSegs];
  int b_off[kMaxGemmSegs];
  void* out;
  int64_t ldc;
  const float* bias;
  const float2* rope_cs;  // EPI_BF16_BIAS_ROPE only
  int rope_S, rope_cols, rope_ld;
  int* arg_idx;  // EPI_ARGMAX only: [M, ldc] next to out = float [M, ldc]
  // Split K (residual-add epilogue only): a tile's k-blocks are cut into `ksplit` slices of `kb_per_slice`; the
  // work unit of the persistent loop is (slice, tile), slice-major.  Every slice adds its partial sum into the
  // output; `split_ctr` holds one turn counter per (tile, CTA, epilogue warp) so that the slices of a tile add
  // in slice order — the result is a fixed sum, independent of timing.
  int ksplit, kb_per_slice;
  int* split_ctr;
};

struct WorkUnit {
  int tile, slice, kb0, kb1;
};
__device__ __forceinline__ WorkUnit work_unit(const KernelParams& p, int u, int num_tiles) {
  WorkUnit w;
  w.slice = u / num_tiles;  // ksplit == 1: always 0
  w.tile = u - w.slice * num_tiles;
  w.kb0 = w.slice * p.kb_per_slice;
  w.kb1 = min(p.num_kb, w.kb0 + p.kb_per_slice);
  return w;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* ptr, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

__device__ __forceinline__ void tile_coords(const KernelParams& p, int tile, int& tm, int& tn) {
  if (p.n_fast) {
    tm = tile / p.tiles_n;
    tn = tile - tm * p.tiles_n;
  } else {
    tn = tile / p.tiles_m;
    tm = tile - tn * p.tiles_m;
  }
}

// ABSB_GEMM_MAXNREG (e.g. -DABSB_GEMM_MAXNREG=104): register cap for the SM-sharing experiment of
// QueryPipeline(coresident=True) — 384 x 104 leaves room for the 8-warp scan CTA of ivf_scan_ring.cu on the
// same SM (profiles/r02_overlap_timeline.md: measured, no net gain, so the default build keeps the
// compiler's own allocation and the epilogues their prefetch registers).
#ifdef ABSB_GEMM_MAXNREG
#define ABSB_GEMM_BOUNDS __maxnreg__(ABSB_GEMM_MAXNREG)
#else
#define ABSB_GEMM_BOUNDS __launch_bounds__(kThreads, 1)
#endif
// CL = 2 ("quad"): a cluster of TWO CTA pairs stacked in M computes a 512 x BN super tile.  The pairs share the B
// tile: every CTA fetches a QUARTER of it and multicasts the piece to its twin in the other pair, so the B operand
// crosses L2 -> SM once per cluster instead of once per pair (-25% operand bytes per flop at BN = 256) — the big
// GEMMs are bound by exactly that traffic (64 B per clock per SM at full tensor rate), most visibly at M = 2048.
template <int BN, int EPI, int NCTA, int CL = 1>
__global__ ABSB_GEMM_BOUNDS void gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmC,
                                                                   const KernelParams p) {
  using L = SmemLayout<BN, NCTA, EPI>;
  const int kStages = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off(kStages));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(CL == 1 || NCTA == 2, "a quad is two CTA pairs");
  const int num_tiles = p.tiles_m * p.tiles_n;  // tiles_m counts rows of (128 * NCTA * CL)-row super tiles
  const int crank = NCTA == 1 ? 0 : (int)tc::cluster_ctarank();
  const int pair = CL == 1 ? 0 : (crank >> 1);  // which pair of the quad
  const int cta_rank = crank & 1;               // rank inside the pair (0 = leader)
  const int worker = blockIdx.x / (NCTA * CL), num_workers = gridDim.x / (NCTA * CL);  // a worker = CTA, pair or quad
  const int num_units = num_tiles * p.ksplit;

  pdl_trigger();  // the next kernel of the stream may set itself up as soon as an SM has room for it
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      tc::mbar_init(full_bar + i, 1);
      tc::mbar_init(empty_bar + i, CL);  // a quad's stage holds bytes issued by BOTH pairs: both MMAs must release it
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tmem_full + i, 1);
      tc::mbar_init(tmem_empty + i, 8 * NCTA);  // one arrive per epilogue warp of every CTA of the pair
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc<NCTA>(tmem_slot, kTmemCols);
    tc::tmem_relinquish<NCTA>();
  }
  tc::tcgen05_fence_before();
  if constexpr (NCTA == 1) __syncthreads();
  else tc::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, tensor-map prefetch) overlapped the tail of the previous kernel; from
  // here on global memory is touched
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) ========
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = worker; u < num_units; u += num_workers) {
        const WorkUnit wu = work_unit(p, u, num_tiles);
        int tm, tn;
        tile_coords(p, wu.tile, tm, tn);
        tm = tm * CL + pair;
        const int row_a = (tm * NCTA + cta_rank) * BM;
        const int row_b = tn * BN + cta_rank * L::kBRows;
        for (int kb = wu.kb0; kb < wu.kb1; ++kb) {
          const int seg = kb / p.seg_kb, within = kb - seg * p.seg_kb;
          tc::mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          if constexpr (NCTA == 1) {
            tc::mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes);
            tc::tma_load_2d(sa, &tmA, full_bar + stage, p.a_off[seg] + within * BK, row_a);
            tc::tma_load_2d(sa + L::kABytes, &tmB, full_bar + stage, p.b_off[seg] + within * BK, row_b);
          } else {
            // both CTAs' bytes are credited to the leader's barrier; only the leader arrives on it
            if (cta_rank == 0) tc::mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes * NCTA);
            tc::tma_load_2d_2cta(sa, &tmA, full_bar + stage, p.a_off[seg] + within * BK, row_a);
            if constexpr (CL == 1) {
              tc::tma_load_2d_2cta(sa + L::kABytes, &tmB, full_bar + stage, p.b_off[seg] + within * BK, row_b);
            } else {
              // quarter `pair` of this CTA's half of B, delivered to this CTA and to its twin in the other pair
              constexpr int kQRows = L::kBRows / 2;
              tc::tma_load_2d_2cta_mc(sa + L::kABytes + pair * kQRows * 128, &tmB, full_bar + stage,
                                      p.b_off[seg] + within * BK, row_b + pair * kQRows,
                                      (uint16_t)((1u << cta_rank) | (1u << (2 + cta_rank))));
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM * NCTA, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = worker; u < num_units; u += num_workers) {
        const WorkUnit wu = work_unit(p, u, num_tiles);
        tc::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
        for (int kb = wu.kb0; kb < wu.kb1; ++kb) {
          tc::mbar_wait(full_bar + stage, phase);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * L::kStageBytes);
          const uint64_t da = tc::make_kmajor_sw128_desc(sa);
          const uint64_t db = tc::make_kmajor_sw128_desc(sa + L::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes along K inside the 128-byte swizzle span = +2 in the (addr >> 4) field
            tc::umma_bf16<NCTA>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                (kb != wu.kb0 || k != 0) ? 1u : 0u);
          }
          if constexpr (CL == 1) tc::umma_commit<NCTA>(empty_bar + stage);
          else tc::umma_commit_2cta_mask(empty_bar + stage, 0xF);  // the stage is shared by the four CTAs
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (CL == 1) tc::umma_commit<NCTA>(tmem_full + acc);
        else tc::umma_commit_2cta_mask(tmem_full + acc, (uint16_t)(3u << (2 * pair)));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (every CTA drains its own 128 accumulator rows) ===============
    // 8 warps: warp w reads TMEM lane quarter w % 4 (hardware restriction) and column half (w - 4) / 4.
    // Each thread owns one output row: 32 consecutive columns per tcgen05.ld, written as 64-128
    // contiguous bytes.  The residual add is a bulk tensor reduction (every element is added exactly once,
    // so the result does not depend on timing).
    const int q = warp & 3;
    const int chalf = (warp - kEpiWarp0) >> 2;
    const uint32_t tmem_empty_addr0 = NCTA == 1 ? 0u : tc::map_to_cta(tc::smem_u32(tmem_empty), (uint32_t)(2 * pair));
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = worker; u < num_units; u += num_workers) {
      const WorkUnit wu = work_unit(p, u, num_tiles);
      int tm, tn;
      tile_coords(p, wu.tile, tm, tn);
      tm = tm * CL + pair;
      const int row = (tm * NCTA + cta_rank) * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      int* turn = nullptr;
      if constexpr (EPI == EPI_F32_ADD && kTmaReduce) {
        if (p.ksplit > 1) {
          // split K: this warp's boxes of the tile are added in slice order.  The earlier slice belongs to a unit
          // with a smaller index — running on another worker or finished — so the wait cannot deadlock.
          turn = p.split_ctr + (((size_t)wu.tile * CL + pair) * NCTA + cta_rank) * 8 + (warp - kEpiWarp0);
          if (lane == 0) {
            while (ld_acquire_gpu(turn) != wu.slice) __nanosleep(40);
            asm volatile("fence.proxy.async;" ::: "memory");  // the reductions below run in the async proxy
          }
          __syncwarp();
        }
      }
      tc::mbar_wait(tmem_full + acc, acc_phase);
      tc::tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);

      if constexpr (EPI == EPI_SWIGLU_BF16) {
        // tile columns [0, BN/2) are gate rows, [BN/2, BN) the matching up rows
        constexpr int H = BN / 2;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        const int ocol0 = tn * H;
        const int ncols = p.N / 2;
#pragma unroll 1
        for (int c = chalf * (H / 2); c < (chalf + 1) * (H / 2); c += 32) {
          uint32_t g[32], u[32];
          tc::tmem_ld_32x32(t_base + c, g);
          tc::tmem_ld_32x32(t_base + H + c, u);
          tc::tmem_ld_wait();
          if (row_ok && ocol0 + c < ncols) {
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * p.ldc + ocol0 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float g0 = __uint_as_float(g[j * 8 + e * 2]), g1 = __uint_as_float(g[j * 8 + e * 2 + 1]);
                const float u0 = __uint_as_float(u[j * 8 + e * 2]), u1 = __uint_as_float(u[j * 8 + e * 2 + 1]);
                w[e] = pack_bf16x2(silu(g0) * u0, silu(g1) * u1);
              }
              dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      } else if constexpr (EPI == EPI_ARGMAX) {
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        const int col0 = tn * BN;
#pragma unroll 1
        for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(t_base + c, v);
          tc::tmem_ld_wait();
          const int col = col0 + c;
          if (col < p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float f = __uint_as_float(v[j]);
              if (f > best) {  // ascending columns + strict compare: ties keep the smallest column
                best = f;
                bidx = col + j;
              }
            }
          }
        }
        if (row_ok) {
          reinterpret_cast<float*>(p.out)[(size_t)row * p.ldc + tn * 2 + chalf] = best;
          p.arg_idx[(size_t)row * p.ldc + tn * 2 + chalf] = bidx;
        }
      } else if constexpr (EPI == EPI_BF16_BIAS_ROPE) {
        // this warp's column half is exactly one 128-wide head: rotate-half pairs (j, j + 64)
        static_assert(BN == 256, "the RoPE epilogue maps one head to one epilogue warp pair");
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        const int hc = tn * BN + chalf * 128;  // first column of the head
        const bool rotate = hc < p.rope_cols;
        // pair-major table: consecutive lanes (= consecutive positions) read consecutive float2
        const float2* cs = p.rope_cs + (row_ok ? row % p.rope_S : 0);
        // 16 rotate-half pairs per step: the TMEM loads are issued first, then the global loads of this step's
        // bias and (cos, sin) entries — all in flight together — and only then the wait; before, the table and
        // bias loads started after the TMEM wait and their L2 latency was paid once per 8 columns.
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
          uint32_t v1[16], v2[16];
          tc::tmem_ld_32x16(t_base + chalf * 128 + c, v1);
          tc::tmem_ld_32x16(t_base + chalf * 128 + 64 + c, v2);
          float ba[16], bb[16];
          float2 t[16];
          const bool live = row_ok && hc < p.N;
          if (live && p.bias) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(p.bias + hc + c + e));
              const float4 y = __ldg(reinterpret_cast<const float4*>(p.bias + hc + 64 + c + e));
              ba[e] = x.x; ba[e + 1] = x.y; ba[e + 2] = x.z; ba[e + 3] = x.w;
              bb[e] = y.x; bb[e + 1] = y.y; bb[e + 2] = y.z; bb[e + 3] = y.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) ba[e] = bb[e] = 0.f;
          }
          if (live && rotate) {
#pragma unroll
            for (int e = 0; e < 16; ++e) t[e] = __ldg(cs + (size_t)(c + e) * p.rope_ld);  // (cos, sin) of pair c + e
          }
          tc::tmem_ld_wait();
          if (live) {
            float a[16], b[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              a[e] = __uint_as_float(v1[e]) + ba[e];
              b[e] = __uint_as_float(v2[e]) + bb[e];
            }
            if (rotate) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float x0 = a[e], y0 = b[e];
                a[e] = x0 * t[e].x - y0 * t[e].y;
                b[e] = y0 * t[e].x + x0 * t[e].y;
              }
            }
            uint4* dst1 = reinterpret_cast<uint4*>(out + (size_t)row * p.ldc + hc + c);
            uint4* dst2 = reinterpret_cast<uint4*>(out + (size_t)row * p.ldc + hc + 64 + c);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              dst1[j] = make_uint4(pack_bf16x2(a[j * 8], a[j * 8 + 1]), pack_bf16x2(a[j * 8 + 2], a[j * 8 + 3]),
                                   pack_bf16x2(a[j * 8 + 4], a[j * 8 + 5]), pack_bf16x2(a[j * 8 + 6], a[j * 8 + 7]));
              dst2[j] = make_uint4(pack_bf16x2(b[j * 8], b[j * 8 + 1]), pack_bf16x2(b[j * 8 + 2], b[j * 8 + 3]),
                                   pack_bf16x2(b[j * 8 + 4], b[j * 8 + 5]), pack_bf16x2(b[j * 8 + 6], b[j * 8 + 7]));
            }
          }
        }
      } else {
        const int col0 = tn * BN;
#pragma unroll 1
        for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(t_base + c, v);
          tc::tmem_ld_wait();
          const int col = col0 + c;
          if constexpr (EPI == EPI_F32_ADD && !kTmaReduce) {
            // fire-and-forget vector reductions: every element is added exactly once, so the result does
            // not depend on timing
            if (row_ok && col < p.N) {
              float* dst = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldc + col;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j * 4),
                             "f"(__uint_as_float(v[j * 4])), "f"(__uint_as_float(v[j * 4 + 1])),
                             "f"(__uint_as_float(v[j * 4 + 2])), "f"(__uint_as_float(v[j * 4 + 3]))
                             : "memory");
            }
          } else if constexpr (EPI == EPI_F32_ADD) {
            // h += acc as ONE bulk tensor reduction per 32 x 32 box: registers -> swizzled shared-memory
            // box -> cp.reduce.async.bulk.tensor (.add, fp32).  Rows past M are clipped by the tensor map;
            // every element is added exactly once, so the result does not depend on timing.
            if (col < p.N) {  // warp-uniform
              uint8_t* stg = smem + L::epi_off(kStages) + (warp - kEpiWarp0) * 4096;
              if (lane == 0) tc::bulk_wait_read_all();  // the previous box has left this buffer
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
              tc::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tc::tma_reduce_add_2d(&tmC, stg, col, (tm * NCTA + cta_rank) * BM + q * 32);
                tc::bulk_commit_group();
              }
            }
          } else if (row_ok && col < p.N) {
            if constexpr (EPI == EPI_BF16_BIAS) {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
              uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * p.ldc + col);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]);
                if (p.bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j * 8));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j * 8 + 4));
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                dst[j] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                    pack_bf16x2(f[6], f[7]));
              }
            } else {
              float* dst = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldc + col;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 f = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]),
                                       __uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3]));
                if (p.bias) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j * 4));
                  f.x += b.x; f.y += b.y; f.z += b.z; f.w += b.w;
                }
                reinterpret_cast<float4*>(dst)[j] = f;
              }
            }
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp of the leader CTA
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        // relaxed: only TMEM reads (already waited for and fenced) are handed over, not the tile's global stores
        if (NCTA == 1 || cta_rank == 0) tc::mbar_arrive_relaxed(tmem_empty + acc);
        else tc::mbar_arrive_cluster_relaxed(tmem_empty_addr0 + (uint32_t)(acc * 8));
      }
      if constexpr (EPI == EPI_F32_ADD && kTmaReduce) {
        if (turn != nullptr && lane == 0) {
          tc::bulk_wait_all();  // this slice's reductions have been performed ...
          __threadfence();
          st_release_gpu(turn, wu.slice + 1 == p.ksplit ? 0 : wu.slice + 1);  // ... the next slice may add (last: reset)
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if constexpr (EPI == EPI_F32_ADD && kTmaReduce) {
      if (lane == 0) tc::bulk_wait_all();  // reductions performed before the CTA (and its smem) goes away
    }
  }

  tc::tcgen05_fence_before();
  if constexpr (NCTA == 1) __syncthreads();
  else tc::cluster_sync();  // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 2) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<NCTA>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ tensor maps -------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  ABSB_CHECK(fn != nullptr, ABSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

// Tensor maps are cached per (buffer, shape, box): a forward launches ~115 GEMMs on the same few dozen
// (pointer, shape) pairs, and cuTensorMapEncodeTiled costs more host time than the launch itself.
struct TmapKey {
  const void* base;
  int64_t rows, cols, ld;
  int box_rows, kind;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && kind == o.kind;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.rows * 0xBF58476D1CE4E5B9ull + (uint64_t)k.cols * 0x94D049BB133111EBull + (uint64_t)k.ld * 31 + (uint64_t)k.box_rows * 7 + k.kind;
    return (size_t)(h ^ (h >> 29));
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

template <typename Make>
CUtensorMap cached_tmap(const TmapKey& key, Make make) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) return it->second;
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();  // descriptors are cheap to rebuild; keep the table bounded
  const CUtensorMap m = make();
  g_tmap_cache.emplace(key, m);
  return m;
}

CUtensorMap make_tmap_uncached(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
// bf16 matrix [rows, cols] with row stride ld (elements); box = box_rows x 64 elements, 128B swizzle
CUtensorMap make_tmap(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return cached_tmap(TmapKey{base, rows, cols, ld, box_rows, 0}, [&] { return make_tmap_uncached(base, rows, cols, ld, box_rows); });
}
CUtensorMap make_tmap_uncached(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  ABSB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, ABSB_ERR_INVALID,
             "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", (long long)ld);
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABSB_CHECK(r == CUDA_SUCCESS, ABSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return m;
}

CUtensorMap make_tmap_f32_box32_uncached(const void* base, int64_t rows, int64_t cols, int64_t ld);
// fp32 matrix [rows, cols] with row pitch ld (elements): 32 x 32 boxes, SWIZZLE_128B (the residual stream)
CUtensorMap make_tmap_f32_box32(const void* base, int64_t rows, int64_t cols, int64_t ld) {
  return cached_tmap(TmapKey{base, rows, cols, ld, 32, 1}, [&] { return make_tmap_f32_box32_uncached(base, rows, cols, ld); });
}
CUtensorMap make_tmap_f32_box32_uncached(const void* base, int64_t rows, int64_t cols, int64_t ld) {
  ABSB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0, ABSB_ERR_INVALID,
             "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", (long long)ld);
  CUtensorMap m;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABSB_CHECK(r == CUDA_SUCCESS, ABSB_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with %d", (int)r);
  return m;
}

// ---- split K for the residual-add GEMMs -----------------------------------------------------------------------
// FFN-down at M = 2048 (the per-GPU token count of the 8-GPU query step) is 64 tiles of 140 k-blocks on 74 CTA
// pairs: one wave, 10 pairs idle, 140 k-block times.  Cut into 8 slices it is 512 units of 17.5 k-blocks = 6.9
// waves = 122.5 k-block times.  The slices of a tile add into the fp32 residual in slice order (turn counters in
// `split_ctr`), so the sum is fixed.  MEASURED on B200 (profiles/r02y_gemm_split_k.log): slower at every shape that
// matters — FFN-down M = 2048: 47.5 us unsplit, 49.1 / 51.8 / 62.2 us with 2 / 4 / 8 slices — because the
// residual-add epilogue of a 256 x 192 tile costs ~9 us (it is what bounds the O-proj, whose main loop is 5 us) and
// every slice pays it, while an unsplit FFN-down hides it under 27 us of MMAs.  Hence: 0 (default) and 1 = never
// split, n = force n slices (tests, micro-benchmark).
int g_gemm_ksplit = 0;
constexpr int kSplitCounters = 1 << 16;

int choose_ksplit(int tiles, int slots, int num_kb, bool allowed) {
  (void)tiles;
  (void)slots;
  if (!allowed || num_kb < 2 || g_gemm_ksplit <= 1) return 1;
  int ks = std::min(g_gemm_ksplit, num_kb);
  while (ks > 1 && (int64_t)(ks - 1) * ceil_div(num_kb, ks) >= num_kb) --ks;  // no empty slice
  return ks;
}

// zero-initialised turn counters, one buffer per (device, stream): every split launch leaves them at zero
int* split_counters(cudaStream_t st) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, int*> bufs;
  int dev = 0;
  ABSB_CUDA(cudaGetDevice(&dev));
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(st) * 64u + (uint64_t)dev;
  std::lock_guard<std::mutex> lock(mu);
  auto it = bufs.find(key);
  if (it != bufs.end()) return it->second;
  int* ptr = nullptr;
  ABSB_CUDA(cudaMalloc(&ptr, kSplitCounters * sizeof(int)));
  ABSB_CUDA(cudaMemset(ptr, 0, kSplitCounters * sizeof(int)));
  ABSB_CUDA(cudaDeviceSynchronize());
  bufs.emplace(key, ptr);
  return ptr;
}

template <int BN, int EPI, int NCTA, int CL = 1>
void launch(const void* A, int64_t a_rows, int64_t a_cols, int64_t lda, const void* B, int64_t b_rows, int64_t b_cols,
            int64_t ldb, KernelParams p, int sms, cudaStream_t st) {
  using L = SmemLayout<BN, NCTA, EPI>;
  auto kern = gemm_bf16_tc_kernel<BN, EPI, NCTA, CL>;
  static bool configured = false;  // per instantiation
  static int max_clusters = 0;     // quads: clusters of 4 CTAs that fit the GPU at once (GPC layout dependent)
  if (!configured) {
    ABSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::dyn(L::kMaxStages)));
    prefer_max_shared(kern);
    configured = true;
  }
  p.stages = L::stages_for(gemm_smem_budget());
  p.tiles_m = (int)ceil_div(p.M, BM * NCTA * CL);
  p.tiles_n = (int)ceil_div(p.N, BN);
  // Few N tiles: walk N first so the CTAs running concurrently share A tiles (A is the big operand of the
  // down/O projections and would otherwise be re-read from HBM once per N tile).  Many N tiles: walk M first
  // so that they share the weight tile.
  p.n_fast = p.tiles_n <= 16 ? 1 : 0;
  const CUtensorMap tmA = make_tmap(A, a_rows, a_cols, lda, BM);
  const CUtensorMap tmB = make_tmap(B, b_rows, b_cols, ldb, L::kBRows / CL);  // a quad fetches B in quarters
  const CUtensorMap tmC = EPI == EPI_F32_ADD ? make_tmap_f32_box32(p.out, p.M, p.N, p.ldc) : tmA;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::dyn(p.stages);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA * CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  int slots = sms / (NCTA * CL);
  if (CL > 1) {
    if (max_clusters == 0) {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3((unsigned)(sms / 4 * 4));
      q.attrs = attr;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n < 1) n = sms / 4 - 1;
      max_clusters = n;
    }
    slots = std::min(slots, max_clusters);
  }
  const int tiles = p.tiles_m * p.tiles_n;
  p.ksplit = 1;
  p.kb_per_slice = p.num_kb;
  p.split_ctr = nullptr;
  if constexpr (EPI == EPI_F32_ADD && kTmaReduce) {
    const int ks = choose_ksplit(tiles, slots, p.num_kb, p.seg_kb == p.num_kb && (int64_t)tiles * NCTA * CL * 8 <= kSplitCounters);
    if (ks > 1) {
      p.ksplit = ks;
      p.kb_per_slice = (int)ceil_div(p.num_kb, ks);
      p.split_ctr = split_counters(st);
    }
  }
  const int workers = std::max(1, std::min(tiles * p.ksplit, slots));
  cfg.gridDim = dim3((unsigned)(workers * NCTA * CL));
  attr[0].val.clusterDim.x = NCTA * CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  ABSB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p));
}

template <int BN, int NCTA, int CL = 1>
void launch_epi(int epi, const void* A, int64_t a_rows, int64_t a_cols, int64_t lda, const void* B, int64_t b_rows,
                int64_t b_cols, int64_t ldb, const KernelParams& p, int sms, cudaStream_t st) {
  switch (epi) {
    case EPI_BF16_BIAS: launch<BN, EPI_BF16_BIAS, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st); break;
    case EPI_F32_BIAS: launch<BN, EPI_F32_BIAS, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st); break;
    case EPI_F32_ADD: launch<BN, EPI_F32_ADD, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st); break;
    case EPI_SWIGLU_BF16:
      if constexpr (BN == 256) {
        launch<BN, EPI_SWIGLU_BF16, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st);
        break;
      }
      fail(ABSB_ERR_INVALID, "SwiGLU epilogue needs 256-column tiles");
    case EPI_ARGMAX:
      if constexpr (BN == 256) {
        launch<BN, EPI_ARGMAX, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st);
        break;
      }
      fail(ABSB_ERR_INVALID, "arg-max epilogue needs 256-column tiles");
    case EPI_BF16_BIAS_ROPE:
      if constexpr (BN == 256) {
        launch<BN, EPI_BF16_BIAS_ROPE, NCTA, CL>(A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, p, sms, st);
        break;
      }
      fail(ABSB_ERR_INVALID, "RoPE epilogue needs 256-column tiles");
    default: fail(ABSB_ERR_INVALID, "unknown epilogue %d", epi);
  }
}

}  // namespace

CUtensorMap make_tmap_bf16(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_tmap(base, rows, cols, ld, box_rows);
}

namespace {

int g_gemm_variant = 0;  // 0 auto; 1 = 1 CTA x BN 256; 2 = CTA pair x BN 256; 3 = CTA pair x BN 192;
                         // 4 = quad (two pairs sharing B by multicast) x BN 256; 5 = quad x BN 192

}  // namespace

void gemm_set_variant(int v) { g_gemm_variant = v; }
void gemm_set_ksplit(int n) { g_gemm_ksplit = n; }
void gemm_set_smem_budget(int bytes) { g_gemm_smem_budget = bytes; }
bool gemm_coresident_mode() { return g_gemm_smem_budget < 200 * 1024; }

void gemm_bf16_tc_ex(int epi, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                     int64_t ldc, const float* bias, const GemmSegs* segs, int sms, cudaStream_t st, const GemmRope* rope,
                     int* arg_idx);

void gemm_bf16_tc(int epi, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                  int64_t ldc, const float* bias, const GemmSegs* segs, int sms, cudaStream_t st, const GemmRope* rope) {
  gemm_bf16_tc_ex(epi, M, N, K, A, lda, B, ldb, out, ldc, bias, segs, sms, st, rope, nullptr);
}

void gemm_bf16_tc_ex(int epi, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                     int64_t ldc, const float* bias, const GemmSegs* segs, int sms, cudaStream_t st, const GemmRope* rope,
                     int* arg_idx) {
  if (M == 0 || N == 0) return;
  ABSB_CHECK(epi != EPI_BF16_BIAS_ROPE || (rope && rope->cs && rope->S >= 1 && rope->ld >= rope->S && rope->cols % 128 == 0 && N % 128 == 0),
             ABSB_ERR_INVALID, "RoPE epilogue needs a table, S >= 1 and 128-wide heads");
  ABSB_CHECK(K > 0 && K % 8 == 0, ABSB_ERR_INVALID, "tcgen05 GEMM needs K %% 8 == 0 (K=%d)", K);
  ABSB_CHECK(N % 32 == 0, ABSB_ERR_INVALID, "tcgen05 GEMM needs N %% 32 == 0 (N=%d)", N);
  ABSB_CHECK(epi != EPI_SWIGLU_BF16 || N % 256 == 0, ABSB_ERR_INVALID, "SwiGLU epilogue needs N %% 256 == 0 (N=%d)", N);
  KernelParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  int64_t a_cols = K, b_cols = K;
  if (segs && segs->nseg > 0) {
    ABSB_CHECK(segs->nseg <= kMaxGemmSegs, ABSB_ERR_INVALID, "too many GEMM segments");
    p.seg_kb = (int)ceil_div(K, BK);
    p.num_kb = p.seg_kb * segs->nseg;
    for (int i = 0; i < segs->nseg; ++i) {
      p.a_off[i] = segs->a_off[i];
      p.b_off[i] = segs->b_off[i];
      ABSB_CHECK(segs->a_off[i] % 8 == 0 && segs->b_off[i] % 8 == 0, ABSB_ERR_INVALID, "segment offsets must be multiples of 8");
    }
    a_cols = segs->a_cols;
    b_cols = segs->b_cols;
    // a segment may not read into its neighbour: K must fill whole k-blocks
    ABSB_CHECK(K % BK == 0, ABSB_ERR_INVALID, "segmented GEMM needs K %% 64 == 0");
  } else {
    p.seg_kb = (int)ceil_div(K, BK);
    p.num_kb = p.seg_kb;
    p.a_off[0] = p.b_off[0] = 0;
  }
  p.out = out;
  p.ldc = ldc;
  p.bias = bias;
  if (rope) {
    p.rope_cs = rope->cs;
    p.rope_S = rope->S;
    p.rope_cols = rope->cols;
    p.rope_ld = rope->ld;
  }
  p.arg_idx = arg_idx;
  ABSB_CHECK(epi != EPI_ARGMAX || arg_idx, ABSB_ERR_INVALID, "arg-max epilogue needs an index buffer");
  const bool needs256 = epi == EPI_SWIGLU_BF16 || epi == EPI_BF16_BIAS_ROPE || epi == EPI_ARGMAX;

  // Tile shape: CTA pairs (cta_group::2, 256-row tiles) whenever there is more than one 128-row block;
  // 192-column tiles when they cut the work into fewer, fuller waves (N = 1536: 8 x 192 instead of 6 x 256).
  // A 192-column tile keeps the tensor pipe ~0.85 as busy as a 256-column one (72% vs 90% active in
  // profiles/r01m_ncu_summary.md), so its column count is charged at 1/0.85: measured on N = 1536,
  // K = 8960 the model then picks 192 at M = 2048 / 4096 and 256 at M = 8192 / 16384, which is the
  // faster variant in all four cases (profiles/r01n_gemm_residual_variants.log).
  int variant = g_gemm_variant;
  if (variant == 0) {
    if (M <= BM) {
      variant = 1;
    } else {
      variant = 2;
      if (!needs256 && N % 192 == 0) {
        const int64_t workers = std::max(1, sms / 2);
        const int64_t tm = ceil_div(M, 2 * BM);
        const int64_t cost256 = ceil_div(tm * ceil_div(N, 256), workers) * 256;
        const int64_t cost192 = ceil_div(tm * ceil_div(N, 192), workers) * 192;
        // (A looser threshold for long-K GEMMs — 192-column tiles for FFN-down at M = 16384, 7 waves instead of 6 x
        // 256 — looked better in one micro-benchmark session and lost in five others; in situ, interleaved bench
        // runs: 41.07 ms per step against 40.33 ms with 256-column tiles, profiles/r02ae_ffn_down_tiles.txt.)
        if (cost192 * 100 < cost256 * 85) variant = 3;
      }
    }
  }
  if (variant == 3 && needs256) variant = 2;
  if (variant == 5 && needs256) variant = 4;
  switch (variant) {
    case 1: launch_epi<256, 1>(epi, A, M, a_cols, lda, B, N, b_cols, ldb, p, sms, st); break;
    case 2: launch_epi<256, 2>(epi, A, M, a_cols, lda, B, N, b_cols, ldb, p, sms, st); break;
    case 3: launch_epi<192, 2>(epi, A, M, a_cols, lda, B, N, b_cols, ldb, p, sms, st); break;
    case 4: launch_epi<256, 2, 2>(epi, A, M, a_cols, lda, B, N, b_cols, ldb, p, sms, st); break;
    case 5: launch_epi<192, 2, 2>(epi, A, M, a_cols, lda, B, N, b_cols, ldb, p, sms, st); break;
    default: fail(ABSB_ERR_INVALID, "unknown GEMM variant %d", variant);
  }
}

// ------------------------------------------------------------------ split-bf16 fp32 GEMM ----
// x = hi + mid + lo with three bf16 terms (8 + 8 + 8 significant bits): the six products below
// carry every term down to 2^-24 relative, accumulated in fp32 inside TMEM, smallest first.
namespace {
__global__ void split3_kernel(int64_t rows, int K, const float* __restrict__ x, __nv_bfloat16* __restrict__ out) {
  const int64_t total = rows * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K;
    const int c = (int)(i - r * K);
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    __nv_bfloat16* o = out + r * 3 * K;
    o[c] = h;
    o[K + c] = m;
    o[2 * K + c] = l;
  }
}
}  // namespace

void split3_bf16(int64_t rows, int K, const float* x, void* out, cudaStream_t st) {
  if (rows == 0) return;
  const int64_t total = rows * K;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  split3_kernel<<<blocks, 256, 0, st>>>(rows, K, x, reinterpret_cast<__nv_bfloat16*>(out));
  ABSB_CUDA(cudaGetLastError());
}

void gemm_split3_f32(int M, int N, int K, const void* A3, const void* B3, float* S, int64_t lds, int sms,
                     cudaStream_t st) {
  GemmSegs segs{};
  // (a part, b part): 0 = hi, 1 = mid, 2 = lo; smallest contributions first
  const int pa[6] = {1, 0, 2, 0, 1, 0};
  const int pb[6] = {1, 2, 0, 1, 0, 0};
  segs.nseg = 6;
  for (int i = 0; i < 6; ++i) {
    segs.a_off[i] = pa[i] * K;
    segs.b_off[i] = pb[i] * K;
  }
  segs.a_cols = segs.b_cols = 3 * (int64_t)K;
  gemm_bf16_tc(EPI_F32_BIAS, M, N, K, A3, 3 * (int64_t)K, B3, 3 * (int64_t)K, S, lds, nullptr, &segs, sms, st);
}

// ------------------------------------------------------------------ fused arg-max -----------
namespace {
// One warp per row: reduce the per-span partials with the order (score desc, column asc).
__global__ void argmax_reduce_kernel(int M, int P, const float* __restrict__ pmax, const int* __restrict__ pidx,
                                     long long* __restrict__ out_idx, float* __restrict__ out_score) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int i = lane; i < P; i += 32) {
    const float f = pmax[(size_t)row * P + i];
    const int c = pidx[(size_t)row * P + i];
    if (f > best || (f == best && c < bidx)) {
      best = f;
      bidx = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float f = __shfl_xor_sync(0xffffffffu, best, o);
    const int c = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (f > best || (f == best && c < bidx)) {
      best = f;
      bidx = c;
    }
  }
  if (lane == 0) {
    out_idx[row] = bidx == 0x7fffffff ? -1 : bidx;
    if (out_score) out_score[row] = best;
  }
}
}  // namespace

int argmax_partials_per_row(int N) { return 2 * (int)ceil_div(N, 256); }

void gemm_split3_argmax(int M, int N, int K, const void* A3, const void* B3, float* ws_max, int* ws_idx,
                        long long* out_idx, float* out_score, int sms, cudaStream_t st) {
  if (M == 0) return;
  GemmSegs segs{};
  const int pa[6] = {1, 0, 2, 0, 1, 0};
  const int pb[6] = {1, 2, 0, 1, 0, 0};
  segs.nseg = 6;
  for (int i = 0; i < 6; ++i) {
    segs.a_off[i] = pa[i] * K;
    segs.b_off[i] = pb[i] * K;
  }
  segs.a_cols = segs.b_cols = 3 * (int64_t)K;
  const int P = argmax_partials_per_row(N);
  gemm_bf16_tc_ex(EPI_ARGMAX, M, N, K, A3, 3 * (int64_t)K, B3, 3 * (int64_t)K, ws_max, P, nullptr, &segs, sms, st,
                  nullptr, ws_idx);
  argmax_reduce_kernel<<<(unsigned)ceil_div((int64_t)M * 32, 256), 256, 0, st>>>(M, P, ws_max, ws_idx, out_idx, out_score);
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb

extern "C" int absb_gemm_set_variant(int variant) {
  ABSB_API_BEGIN
  ABSB_CHECK(variant >= 0 && variant <= 5, ABSB_ERR_INVALID, "GEMM variant %d outside [0,5]", variant);
  absb::gemm_set_variant(variant);
  ABSB_API_END
}

extern "C" int absb_gemm_set_ksplit(int slices) {
  ABSB_API_BEGIN
  ABSB_CHECK(slices >= 0 && slices <= 64, ABSB_ERR_INVALID, "GEMM k-slices %d outside [0,64]", slices);
  absb::gemm_set_ksplit(slices);
  ABSB_API_END
}

extern "C" int absb_gemm_set_smem_budget(int bytes) {
  ABSB_API_BEGIN
  ABSB_CHECK(bytes == 0 || (bytes >= 96 * 1024 && bytes <= 227 * 1024), ABSB_ERR_INVALID,
             "GEMM shared-memory budget %d outside [96 KB, 227 KB] (0 = all of it)", bytes);
  absb::gemm_set_smem_budget(bytes == 0 ? 227 * 1024 : bytes);
  ABSB_API_END
}

extern "C" int absb_gemm_bf16_epi_dev(int device, int epi, int M, int N, int K, const void* A_dev, const void* B_dev,
                                      void* out_dev, int64_t ldc, const float* bias_dev, void* stream) {
  ABSB_API_BEGIN
  using namespace absb;
  ABSB_CHECK(M >= 0 && N >= 0 && K > 0 && A_dev && B_dev && out_dev, ABSB_ERR_INVALID, "bad GEMM arguments");
  ABSB_CHECK(epi == EPI_BF16_BIAS || epi == EPI_F32_BIAS || epi == EPI_F32_ADD || epi == EPI_SWIGLU_BF16, ABSB_ERR_INVALID,
             "epilogue %d is not available through this entry", epi);
  DeviceGuard g(device);
  const DeviceProps pr = device_props(device);
  ABSB_CHECK(pr.cc_major == 10, ABSB_ERR_UNSUPPORTED, "tcgen05 GEMM needs an sm_100 device");
  gemm_bf16_tc(epi, M, N, K, A_dev, K, B_dev, K, out_dev, ldc, bias_dev, nullptr, pr.sm_count, (cudaStream_t)stream);
  ABSB_API_END
}

extern "C" int absb_gemm_bf16_dev(int device, int M, int N, int K, const void* A_dev, const void* B_dev, float* C_dev,
                                  void* stream) {
  ABSB_API_BEGIN
  using namespace absb;
  ABSB_CHECK(M >= 0 && N >= 0 && K > 0 && A_dev && B_dev && C_dev, ABSB_ERR_INVALID, "bad GEMM arguments");
  DeviceGuard g(device);
  const DeviceProps pr = device_props(device);
  ABSB_CHECK(pr.cc_major == 10, ABSB_ERR_UNSUPPORTED, "tcgen05 GEMM needs an sm_100 device");
  gemm_bf16_tc(EPI_F32_BIAS, M, N, K, A_dev, K, B_dev, K, C_dev, N, nullptr, nullptr, pr.sm_count, (cudaStream_t)stream);
  ABSB_API_END
}
