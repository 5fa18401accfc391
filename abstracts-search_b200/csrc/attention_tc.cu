// attention_tc.cu — bidirectional GQA attention of the stella/Qwen2 encoder on the 5th-generation
// tensor cores (the bulk-encode batches of `sidecar-search build -b 32`, /root/reference/Makefile:65, and
// app.py's queries, README.md:28): sequences of up to 256 tokens in one key block (below),
// 257-512 tokens (stella's max_seq_length) in two key blocks (attention_tc_long_kernel).
//
// One CTA = one 128-row tile of queries of one (sequence, kv head): the rows are 128 consecutive
// positions of one q head (S >= 128) or the whole sequences of several q heads that share the kv
// head (S < 128: 128 / roundup(S, 8) heads per tile), so K and V are staged once per tile.
//
//   TMA      Q [128 x 128], K [NK x 128] (K-major), V [NK x 128] (MN-major B operand) -> smem
//   MMA 1    S[128 x NK] = Q K^T          tcgen05.mma SS, fp32 accumulator in TMEM cols [0, NK)
//   softmax  thread = row: tcgen05.ld, scale, key-padding (+ causal) mask, max, exp2, sum; the
//            probabilities go back to TMEM as packed bf16 (tcgen05.st) over the columns just read
//   MMA 2    O[128 x 128] = P V           tcgen05.mma TS (A = P from TMEM), TMEM cols [128, 256)
//   epilogue tcgen05.ld O, * 1/sum, bf16, row-contiguous global stores
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc.cuh"

namespace absb {

namespace {

constexpr int kHD = 128;
constexpr int kThreads = 128;
constexpr int kTmemCols = 256;
constexpr int kOCol = 128;

struct AttnParams {
  int S, nh, nkv, causal;
  int NK;            // keys staged = roundup(S, 16)
  int RB;            // rows of one head sub-block inside the 128-row tile
  int heads_per_blk; // q heads per tile (S < 128) or 1
  int blks_per_head; // tiles per q head (S >= 128) or 1
  float scale_log2;
  const int* mask;   // [T] 1 = token
  __nv_bfloat16* out;
  int ldo;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ __launch_bounds__(kThreads) void attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmKV,
                                                                const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NK = p.NK;
  const uint32_t kv_half = (uint32_t)NK * 128u;  // bytes of one 64-column half of K or V
  uint8_t* sQ = smem;                            // [2][128][128 B]
  uint8_t* sK = sQ + 2 * 128 * 128;              // [2][NK][128 B]
  uint8_t* sV = sK + 2 * kv_half;                // [2][NK][128 B]
  float* sBias = reinterpret_cast<float*>(sV + 2 * kv_half);  // [NK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);  // qk_full, v_full, s_done, o_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int blk = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int group = p.nh / p.nkv;
  const int S = p.S;
  const int64_t tok0 = (int64_t)b * S;
  // tile -> first head (within the group), first position, heads in this tile
  int h0, p0, nheads;
  if (p.blks_per_head > 1 || p.heads_per_blk == 1) {
    h0 = blk / p.blks_per_head;
    p0 = (blk % p.blks_per_head) * 128;
    nheads = 1;
  } else {
    h0 = blk * p.heads_per_blk;
    p0 = 0;
    nheads = min(p.heads_per_blk, group - h0);
  }

  pdl_trigger();
  if (tid == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmKV);
    for (int i = 0; i < 4; ++i) tc::mbar_init(bars + i, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc<1>(tmem_slot, kTmemCols);
    tc::tmem_relinquish<1>();
  }
  pdl_wait();  // global memory from here on
  // key bias: 0 for real tokens of this sequence, -inf for padding and for the rows past S
  for (int k = tid; k < NK; k += kThreads) sBias[k] = (k < S && p.mask[tok0 + k] != 0) ? 0.f : -INFINITY;
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    // ---- TMA: Q sub-blocks + K on one barrier, V on another ----
    const uint32_t q_bytes = (uint32_t)nheads * 2u * (uint32_t)p.RB * 128u;
    tc::mbar_arrive_expect_tx(bars + 0, q_bytes + 2 * kv_half);
    for (int h = 0; h < nheads; ++h)
      for (int half = 0; half < 2; ++half)
        tc::tma_load_2d(sQ + half * 16384 + h * p.RB * 128, &tmQ, bars + 0,
                        (kvh * group + h0 + h) * kHD + half * 64, (int)(tok0 + p0));
    for (int half = 0; half < 2; ++half)
      tc::tma_load_2d(sK + half * kv_half, &tmKV, bars + 0, (p.nh + kvh) * kHD + half * 64, (int)tok0);
    tc::mbar_arrive_expect_tx(bars + 1, 2 * kv_half);
    for (int half = 0; half < 2; ++half)
      tc::tma_load_2d(sV + half * kv_half, &tmKV, bars + 1, (p.nh + p.nkv + kvh) * kHD + half * 64, (int)tok0);
    // ---- MMA 1: S = Q K^T ----
    tc::mbar_wait(bars + 0, 0);
    tc::tcgen05_fence_after();
    const uint32_t idesc1 = tc::make_idesc_bf16_f32(128, NK);
#pragma unroll
    for (int j = 0; j < kHD / 16; ++j) {
      const uint32_t off = (uint32_t)(j >> 2), within = (uint32_t)(j & 3) * 32u;
      const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(sQ) + off * 16384u + within);
      const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(sK) + off * kv_half + within);
      tc::umma_bf16<1>(tmem_base, da, db, idesc1, j != 0 ? 1u : 0u);
    }
    tc::umma_commit<1>(bars + 2);
  }
  __syncwarp();

  // ---- softmax: thread = tile row = TMEM lane ----
  const int sub = tid / p.RB;
  const int pos = p0 + tid - sub * p.RB;
  const bool row_ok = sub < nheads && pos < S;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  tc::mbar_wait(bars + 2, 0);
  tc::tcgen05_fence_after();
  float m = -INFINITY;
  for (int c = 0; c < NK; c += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32(t_row + c, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float s = __uint_as_float(v[j]) * p.scale_log2 + ((c + j < NK) ? sBias[c + j] : -INFINITY);
      if (p.causal && c + j > pos) s = -INFINITY;
      m = fmaxf(m, s);
    }
  }
  const float m_safe = (m == -INFINITY || !(m == m)) ? 0.f : m;
  float sum = 0.f;
  for (int c = 0; c < NK; c += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32(t_row + c, v);
    tc::tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float s0 = __uint_as_float(v[2 * j]) * p.scale_log2 + ((c + 2 * j < NK) ? sBias[c + 2 * j] : -INFINITY);
      float s1 = __uint_as_float(v[2 * j + 1]) * p.scale_log2 + ((c + 2 * j + 1 < NK) ? sBias[c + 2 * j + 1] : -INFINITY);
      if (p.causal) {
        if (c + 2 * j > pos) s0 = -INFINITY;
        if (c + 2 * j + 1 > pos) s1 = -INFINITY;
      }
      const float e0 = exp2f(s0 - m_safe), e1 = exp2f(s1 - m_safe);
      sum += e0 + e1;
      pk[j] = pack2(e0, e1);
    }
    // P overwrites the (already consumed) low columns of S: keys [c, c+32) -> columns [c/2, c/2+16)
    tc::tmem_st_32x16(t_row + (c >> 1), pk);
  }
  tc::tmem_st_wait();
  tc::tcgen05_fence_before();
  __syncthreads();

  if (tid == 0) {
    // ---- MMA 2: O = P V (A from TMEM, B = V MN-major) ----
    tc::tcgen05_fence_after();
    tc::mbar_wait(bars + 1, 0);
    tc::tcgen05_fence_after();
    const uint32_t idesc2 = tc::make_idesc_bf16_f32(128, kHD, 1);
    for (int j = 0; j < NK / 16; ++j) {
      const uint64_t db = tc::make_mnmajor_sw128_desc(tc::smem_u32(sV) + (uint32_t)j * 2048u, kv_half);
      tc::umma_bf16_ts(tmem_base + kOCol, tmem_base + (uint32_t)(j * 8), db, idesc2, j != 0 ? 1u : 0u);
    }
    tc::umma_commit<1>(bars + 3);
  }
  __syncwarp();

  // ---- epilogue ----
  tc::mbar_wait(bars + 3, 0);
  tc::tcgen05_fence_after();
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  __nv_bfloat16* orow = p.out + (size_t)(tok0 + pos) * p.ldo + (size_t)(kvh * group + h0 + sub) * kHD;
#pragma unroll 1
  for (int c = 0; c < kHD; c += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32(t_row + kOCol + c, v);
    tc::tmem_ld_wait();
    if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack2(__uint_as_float(v[j * 8]) * inv, __uint_as_float(v[j * 8 + 1]) * inv),
                            pack2(__uint_as_float(v[j * 8 + 2]) * inv, __uint_as_float(v[j * 8 + 3]) * inv),
                            pack2(__uint_as_float(v[j * 8 + 4]) * inv, __uint_as_float(v[j * 8 + 5]) * inv),
                            pack2(__uint_as_float(v[j * 8 + 6]) * inv, __uint_as_float(v[j * 8 + 7]) * inv));
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}


// ================================================================================================
// 256 < S <= 512 (stella's max_seq_length).  The 512-key score tile takes all 512 TMEM columns and K + V
// (256 KB) do not fit shared memory together, so the tile is laid out around that:
//
//   TMA      Q [128 x 128], K_0 [256 x 128], K_1 [<= 256 x 128]                         (160 KB)
//   MMA 1    S_0 = Q K_0^T -> TMEM cols [0, 256),  S_1 = Q K_1^T -> cols [256, 512)      (tcgen05.mma SS)
//   TMA      V_0, V_1 into the K buffers as soon as both products are complete — under the softmax
//   softmax  TWO warpgroups, one per key block, thread = row: row maximum of the block -> shared memory ->
//            maximum over BOTH blocks, so every probability is exp2(s - m) with the same m and the two P V
//            products accumulate into ONE accumulator without any rescaling; P_b goes back to TMEM as packed
//            bf16 over the low half of S_b
//   MMA 2    O = P_0 V_0 + P_1 V_1 -> TMEM cols [128, 256) (the consumed upper half of S_0)  (tcgen05.mma TS)
//   epilogue both warpgroups, 64 columns each: tcgen05.ld O, * 1 / (l_0 + l_1), bf16, row-contiguous stores
//
// One CTA = 128 consecutive positions of one q head; K / V of a (sequence, kv head) are re-read from L2 by
// its 6 x 4 tiles.
// ================================================================================================
constexpr int kLongKB = 256;      // keys per block
constexpr int kLongThreads = 256;  // two softmax warpgroups

__global__ __launch_bounds__(kLongThreads) void attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                         const __grid_constant__ CUtensorMap tmKV,
                                                                         const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t kv_half = (uint32_t)kLongKB * 128u;  // bytes of one 64-column half of a K or V block
  constexpr uint32_t kv_blk = 2 * kv_half;                // one block of K (later V): 64 KB
  uint8_t* sQ = smem;                                     // [2][128][128 B]
  uint8_t* sKV = sQ + 2 * 128 * 128;                      // [2 blocks][2 halves][256][128 B]
  uint32_t* kvalid = reinterpret_cast<uint32_t*>(sKV + 2 * kv_blk);  // [16] + pad: key-valid bits of the 512 keys
  float* sMax = reinterpret_cast<float*>(kvalid + 512);   // [2][128]
  float* sSum = sMax + 256;                               // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 256);  // qk_full, s_done, v_full, o_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int wg = tid >> 7, row = tid & 127;
  const int blk = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int group = p.nh / p.nkv;
  const int S = p.S;
  const int64_t tok0 = (int64_t)b * S;
  const int h0 = blk / p.blks_per_head;
  const int p0 = (blk % p.blks_per_head) * 128;
  const int nk1 = ((S - kLongKB + 15) / 16) * 16;  // keys of block 1 that the MMAs touch (16 .. 256)

  pdl_trigger();
  if (tid == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmKV);
    for (int i = 0; i < 4; ++i) tc::mbar_init(bars + i, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc<1>(tmem_slot, 512);
    tc::tmem_relinquish<1>();
  }
  pdl_wait();  // global memory from here on
  // key-valid bits: real token of this sequence (padding, rows past S and the next sequence's rows are masked)
  for (int w = warp; w < 2 * kLongKB / 32; w += kLongThreads / 32) {
    const int k = w * 32 + (tid & 31);
    const unsigned bits = __ballot_sync(0xffffffffu, k < S && p.mask[tok0 + k] != 0);
    if ((tid & 31) == 0) kvalid[w] = bits;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int pos = p0 + row;
  const bool row_ok = pos < S;
  const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int k_col = (p.nh + kvh) * kHD, v_col = (p.nh + p.nkv + kvh) * kHD;

  if (tid == 0) {
    // rows past the end of the tensor are zero-filled by TMA, rows of the next sequence are finite numbers: both
    // are masked out by the key-valid bits
    tc::mbar_arrive_expect_tx(bars + 0, 2u * 16384u + 2 * kv_blk);
    for (int half = 0; half < 2; ++half)
      tc::tma_load_2d(sQ + half * 16384, &tmQ, bars + 0, (kvh * group + h0) * kHD + half * 64, (int)(tok0 + p0));
    for (int kb = 0; kb < 2; ++kb)
      for (int half = 0; half < 2; ++half)
        tc::tma_load_2d(sKV + kb * kv_blk + half * kv_half, &tmKV, bars + 0, k_col + half * 64, (int)tok0 + kb * kLongKB);
    // ---- S_0 = Q K_0^T, S_1 = Q K_1^T ----
    tc::mbar_wait(bars + 0, 0);
    tc::tcgen05_fence_after();
    for (int kb = 0; kb < 2; ++kb) {
      const uint32_t idesc1 = tc::make_idesc_bf16_f32(128, kb == 0 ? kLongKB : nk1);
#pragma unroll
      for (int j = 0; j < kHD / 16; ++j) {
        const uint32_t off = (uint32_t)(j >> 2), within = (uint32_t)(j & 3) * 32u;
        const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(sQ) + off * 16384u + within);
        const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(sKV) + kb * kv_blk + off * kv_half + within);
        tc::umma_bf16<1>(tmem_base + (uint32_t)(kb * 256), da, db, idesc1, j != 0 ? 1u : 0u);
      }
    }
    tc::umma_commit<1>(bars + 1);
    // both products complete -> the K buffers are free: V_0, V_1 land under the softmax
    tc::mbar_wait(bars + 1, 0);
    tc::mbar_arrive_expect_tx(bars + 2, 2 * kv_blk);
    for (int kb = 0; kb < 2; ++kb)
      for (int half = 0; half < 2; ++half)
        tc::tma_load_2d(sKV + kb * kv_blk + half * kv_half, &tmKV, bars + 2, v_col + half * 64, (int)tok0 + kb * kLongKB);
  }
  __syncwarp();

  // ---- softmax: warpgroup wg owns key block wg, thread = tile row = TMEM lane ----
  // Key validity is a bit mask per 32 keys (padding, keys past S, causal): a chunk whose 32 keys are all valid —
  // the common case — costs one FMNMX per score in the max pass and FFMA + MUFU.EX2 + FADD in the exp pass.
  const int nk = wg == 0 ? kLongKB : nk1;
  const int key0 = wg * kLongKB;
  const uint32_t t_blk = t_row + (uint32_t)(wg * 256);
  tc::mbar_wait(bars + 1, 0);
  tc::tcgen05_fence_after();
  auto key_mask = [&](int c) -> unsigned {  // c: first key of the chunk inside the block
    unsigned km = kvalid[(key0 + c) >> 5];
    if (p.causal) {
      const int kabs = key0 + c;
      km &= pos < kabs ? 0u : (pos - kabs >= 31 ? 0xffffffffu : ((2u << (pos - kabs)) - 1u));
    }
    return km;
  };
  float m = -INFINITY;  // maximum of the RAW scores (the scale is positive)
  {
    uint32_t va[32], vb[32];
    auto row_max = [&](const uint32_t (&v)[32], int c) {
      const unsigned km = key_mask(c);
      if (km == 0xffffffffu) {
        float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[j]));
          m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
        }
        m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((km >> j) & 1u) m = fmaxf(m, __uint_as_float(v[j]));
      }
    };
    tc::tmem_ld_32x32(t_blk, va);
    for (int c = 0; c < nk; c += 64) {
      tc::tmem_ld_wait();
      if (c + 32 < nk) tc::tmem_ld_32x32(t_blk + c + 32, vb);
      row_max(va, c);
      if (c + 32 < nk) {
        tc::tmem_ld_wait();
        if (c + 64 < nk) tc::tmem_ld_32x32(t_blk + c + 64, va);
        row_max(vb, c + 32);
      }
    }
  }
  sMax[wg * 128 + row] = m;
  __syncthreads();
  m = fmaxf(sMax[row], sMax[128 + row]);
  const float ms = (m == -INFINITY || !(m == m)) ? 0.f : m * p.scale_log2;
  float sum = 0.f;
  {
    uint32_t va[32], vb[32];
    auto probs = [&](const uint32_t (&v)[32], int c) {
      const unsigned km = key_mask(c);
      uint32_t pk[16];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (km == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -ms));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -ms));
          const float e2 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 2]), p.scale_log2, -ms));
          const float e3 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 3]), p.scale_log2, -ms));
          s0 += e0; s1 += e1; s2 += e2; s3 += e3;
          pk[j] = pack2(e0, e1);
          pk[j + 1] = pack2(e2, e3);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -ms));
          float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -ms));
          if (!((km >> (2 * j)) & 1u)) e0 = 0.f;
          if (!((km >> (2 * j + 1)) & 1u)) e1 = 0.f;
          s0 += e0; s1 += e1;
          pk[j] = pack2(e0, e1);
        }
      }
      sum += (s0 + s1) + (s2 + s3);
      tc::tmem_st_32x16(t_blk + (c >> 1), pk);  // P over the consumed low columns of S_wg
    };
    tc::tmem_ld_32x32(t_blk, va);
    for (int c = 0; c < nk; c += 64) {
      tc::tmem_ld_wait();
      if (c + 32 < nk) tc::tmem_ld_32x32(t_blk + c + 32, vb);
      probs(va, c);
      if (c + 32 < nk) {
        tc::tmem_ld_wait();
        if (c + 64 < nk) tc::tmem_ld_32x32(t_blk + c + 64, va);
        probs(vb, c + 32);
      }
    }
  }
  sSum[wg * 128 + row] = sum;
  tc::tmem_st_wait();
  tc::tcgen05_fence_before();
  __syncthreads();

  if (tid == 0) {
    // ---- O = P_0 V_0 + P_1 V_1, one accumulator (same row maximum in both blocks) ----
    tc::tcgen05_fence_after();
    tc::mbar_wait(bars + 2, 0);
    tc::tcgen05_fence_after();
    const uint32_t idesc2 = tc::make_idesc_bf16_f32(128, kHD, 1);
    for (int kb = 0; kb < 2; ++kb) {
      const int steps = (kb == 0 ? kLongKB : nk1) / 16;
      for (int j = 0; j < steps; ++j) {
        const uint64_t db = tc::make_mnmajor_sw128_desc(tc::smem_u32(sKV) + kb * kv_blk + (uint32_t)j * 2048u, kv_half);
        tc::umma_bf16_ts(tmem_base + kOCol, tmem_base + (uint32_t)(kb * 256 + j * 8), db, idesc2, (kb | j) != 0 ? 1u : 0u);
      }
    }
    tc::umma_commit<1>(bars + 3);
  }
  __syncwarp();

  // ---- epilogue: 64 output columns per warpgroup ----
  tc::mbar_wait(bars + 3, 0);
  tc::tcgen05_fence_after();
  const float l = sSum[row] + sSum[128 + row];
  const float inv = l > 0.f ? 1.f / l : 0.f;
  __nv_bfloat16* orow = p.out + (size_t)(tok0 + pos) * p.ldo + (size_t)(kvh * group + h0) * kHD + wg * 64;
  {
    uint32_t v0[32], v1[32];
    tc::tmem_ld_32x32(t_row + kOCol + wg * 64, v0);
    tc::tmem_ld_32x32(t_row + kOCol + wg * 64 + 32, v1);
    tc::tmem_ld_wait();
    if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack2(__uint_as_float(v0[j * 8]) * inv, __uint_as_float(v0[j * 8 + 1]) * inv),
                            pack2(__uint_as_float(v0[j * 8 + 2]) * inv, __uint_as_float(v0[j * 8 + 3]) * inv),
                            pack2(__uint_as_float(v0[j * 8 + 4]) * inv, __uint_as_float(v0[j * 8 + 5]) * inv),
                            pack2(__uint_as_float(v0[j * 8 + 6]) * inv, __uint_as_float(v0[j * 8 + 7]) * inv));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[4 + j] = make_uint4(pack2(__uint_as_float(v1[j * 8]) * inv, __uint_as_float(v1[j * 8 + 1]) * inv),
                                pack2(__uint_as_float(v1[j * 8 + 2]) * inv, __uint_as_float(v1[j * 8 + 3]) * inv),
                                pack2(__uint_as_float(v1[j * 8 + 4]) * inv, __uint_as_float(v1[j * 8 + 5]) * inv),
                                pack2(__uint_as_float(v1[j * 8 + 6]) * inv, __uint_as_float(v1[j * 8 + 7]) * inv));
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<1>(tmem_base, 512);
  }
}

// ================================================================================================
// Persistent, warp-specialised version: one CTA per SM walks a contiguous range of tiles.
//   warp 0      TMA producer: K/V once per (sequence, kv head) into a 1-2 stage ring (+ the key-valid
//               bit mask), Q per tile into a 2-stage ring
//   warp 1      MMA issuer: S = Q K^T of tile i is issued before P V of tile i-1, so the tensor core
//               works on one tile while the softmax warps work on the other
//   warps 4-7   softmax + epilogue of even tiles (TMEM slot 0: columns [0, 256))
//   warps 8-11  softmax + epilogue of odd tiles  (TMEM slot 1: columns [256, 512))
// ================================================================================================
constexpr int kPThreads = 384;

struct TileInfo {
  int unit, b, kvh, h0, p0, nheads;
  bool first_of_unit, last_of_unit;
};

__device__ __forceinline__ TileInfo decode_tile(const AttnParams& p, int t, int tiles_per_unit, int t_begin, int t_end) {
  TileInfo ti;
  ti.unit = t / tiles_per_unit;
  const int blk = t - ti.unit * tiles_per_unit;
  ti.b = ti.unit / p.nkv;
  ti.kvh = ti.unit - ti.b * p.nkv;
  const int group = p.nh / p.nkv;
  if (p.heads_per_blk == 1) {
    ti.h0 = blk / p.blks_per_head;
    ti.p0 = (blk - ti.h0 * p.blks_per_head) * 128;
    ti.nheads = 1;
  } else {
    ti.h0 = blk * p.heads_per_blk;
    ti.p0 = 0;
    ti.nheads = min(p.heads_per_blk, group - ti.h0);
  }
  ti.first_of_unit = (t == t_begin) || blk == 0;
  ti.last_of_unit = (t == t_end - 1) || blk == tiles_per_unit - 1;
  return ti;
}


template <int KVS>
__device__ __forceinline__ void attention_tc_persistent_body(const CUtensorMap& tmQ, const CUtensorMap& tmKV,
                                                             const AttnParams& p, int total_tiles, int tiles_per_unit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NK = p.NK;
  const uint32_t kv_half = (uint32_t)NK * 128u;
  const uint32_t kv_stage = 4 * kv_half;       // K halves, V halves
  uint8_t* sQ = smem;                           // [2 stages][2 halves][128][128 B]
  uint8_t* sKV = sQ + 2 * 32768;                // [KVS][K0 K1 V0 V1]
  uint32_t* kvalid = reinterpret_cast<uint32_t*>(sKV + KVS * kv_stage);  // [KVS][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(kvalid + KVS * 8);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [2]
  uint64_t* kv_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_ready = bars + 10;      // [2]
  uint64_t* o_full = bars + 12;       // [2]
  uint64_t* slot_free = bars + 14;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = min(total_tiles, (int)blockIdx.x * per), t_end = min(total_tiles, t_begin + per);
  const int group = p.nh / p.nkv;
  const int S = p.S;

  pdl_trigger();
  if (tid == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmKV);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(q_full + i, 1);
      tc::mbar_init(q_empty + i, 1);
      tc::mbar_init(kv_full + i, 1);
      tc::mbar_init(kv_empty + i, 1);
      tc::mbar_init(s_full + i, 1);
      tc::mbar_init(p_ready + i, 4);
      tc::mbar_init(o_full + i, 1);
      tc::mbar_init(slot_free + i, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc<1>(tmem_slot, 512);
    tc::tmem_relinquish<1>();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0) {
    // ===================== producer =====================
    int n_unit = -1;  // units seen by this CTA so far - 1
    for (int t = t_begin; t < t_end; ++t) {
      const int i = t - t_begin;
      const TileInfo ti = decode_tile(p, t, tiles_per_unit, t_begin, t_end);
      const int64_t tok0 = (int64_t)ti.b * S;
      if (ti.first_of_unit) {
        ++n_unit;
        const int ks = n_unit % KVS;
        const uint32_t kph = (uint32_t)(n_unit / KVS) & 1u;
        tc::mbar_wait(kv_empty + ks, kph ^ 1);
        // key-valid bits: real token of this sequence (padding and rows past S are masked out)
        for (int c = 0; c < NK; c += 32) {
          const int k = c + lane;
          const bool ok = k < S && p.mask[tok0 + k] != 0;
          const unsigned w = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) kvalid[ks * 8 + (c >> 5)] = w;
        }
        __syncwarp();
        if (lane == 0) {
          uint8_t* st = sKV + ks * kv_stage;
          tc::mbar_arrive_expect_tx(kv_full + ks, kv_stage);
          for (int half = 0; half < 2; ++half) {
            tc::tma_load_2d(st + half * kv_half, &tmKV, kv_full + ks, (p.nh + ti.kvh) * kHD + half * 64, (int)tok0);
            tc::tma_load_2d(st + (2 + half) * kv_half, &tmKV, kv_full + ks, (p.nh + p.nkv + ti.kvh) * kHD + half * 64,
                            (int)tok0);
          }
        }
      }
      if (lane == 0) {
        const int qs = i & 1;
        tc::mbar_wait(q_empty + qs, ((uint32_t)(i >> 1) & 1u) ^ 1);
        tc::mbar_arrive_expect_tx(q_full + qs, (uint32_t)ti.nheads * 2u * (uint32_t)p.RB * 128u);
        for (int h = 0; h < ti.nheads; ++h)
          for (int half = 0; half < 2; ++half)
            tc::tma_load_2d(sQ + qs * 32768 + half * 16384 + h * p.RB * 128, &tmQ, q_full + qs,
                            (ti.kvh * group + ti.h0 + h) * kHD + half * 64, (int)(tok0 + ti.p0));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Dynamic order: whichever of {P V of the oldest tile whose probabilities are ready, Q K^T of the
    // next tile whose operands and TMEM slot are ready} can go is issued first, so a finished softmax
    // never waits behind a Q K^T that is itself waiting for an epilogue.
    if (lane == 0) {
      const uint32_t idesc1 = tc::make_idesc_bf16_f32(128, NK);
      const uint32_t idesc2 = tc::make_idesc_bf16_f32(128, kHD, 1);
      const int n_tiles = t_end - t_begin;
      int next_qk = 0, next_pv = 0;
      int n_unit = -1;
      int ks_of[2] = {0, 0};
      bool last_of[2] = {false, false};
      while (next_pv < n_tiles) {
        // ---- P V of tile next_pv ----
        if (next_pv < next_qk) {
          const int j = next_pv, slot = j & 1;
          if (tc::mbar_test_wait(p_ready + slot, (uint32_t)(j >> 1) & 1u)) {
            tc::tcgen05_fence_after();
            const int ks = ks_of[slot];
            const uint32_t sv = tc::smem_u32(sKV + ks * kv_stage + 2 * kv_half);
            const uint32_t tb = tmem_base + (uint32_t)(slot * 256);
            for (int k = 0; k < NK / 16; ++k) {
              const uint64_t db = tc::make_mnmajor_sw128_desc(sv + (uint32_t)k * 2048u, kv_half);
              tc::umma_bf16_ts(tb + kOCol, tb + (uint32_t)(k * 8), db, idesc2, k != 0 ? 1u : 0u);
            }
            tc::umma_commit<1>(o_full + slot);
            if (last_of[slot]) tc::umma_commit<1>(kv_empty + ks);
            ++next_pv;
          }
        }
        // ---- Q K^T of tile next_qk ----
        if (next_qk < n_tiles && next_qk - next_pv < 2) {
          const int i = next_qk, qs = i & 1, slot = i & 1;
          const uint32_t ph = (uint32_t)(i >> 1) & 1u;
          const TileInfo ti = decode_tile(p, t_begin + i, tiles_per_unit, t_begin, t_end);
          const int nu = n_unit + (ti.first_of_unit ? 1 : 0);
          const int ks = nu % KVS;
          // a single K/V stage is refilled only after the previous unit's last P V: that one goes first
          const bool blocked = (KVS == 1) && ti.first_of_unit && i > 0 && next_pv < i;
          if (!blocked && tc::mbar_test_wait(q_full + qs, ph) && tc::mbar_test_wait(slot_free + slot, ph ^ 1) &&
              (!ti.first_of_unit || tc::mbar_test_wait(kv_full + ks, (uint32_t)(nu / KVS) & 1u))) {
            tc::tcgen05_fence_after();
            const uint32_t sq = tc::smem_u32(sQ + qs * 32768);
            const uint32_t sk = tc::smem_u32(sKV + ks * kv_stage);
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) {
              const uint32_t off = (uint32_t)(k >> 2), within = (uint32_t)(k & 3) * 32u;
              const uint64_t da = tc::make_kmajor_sw128_desc(sq + off * 16384u + within);
              const uint64_t db = tc::make_kmajor_sw128_desc(sk + off * kv_half + within);
              tc::umma_bf16<1>(tmem_base + (uint32_t)(slot * 256), da, db, idesc1, k != 0 ? 1u : 0u);
            }
            tc::umma_commit<1>(s_full + slot);
            tc::umma_commit<1>(q_empty + qs);
            n_unit = nu;
            ks_of[slot] = ks;
            last_of[slot] = ti.last_of_unit;
            ++next_qk;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue =====================
    const int grp = (warp - 4) >> 2;  // tiles with (i & 1) == grp
    const int row = (warp & 3) * 32 + lane;
    int n_unit = -1;
    for (int t = t_begin; t < t_end; ++t) {
      const int i = t - t_begin;
      const TileInfo ti = decode_tile(p, t, tiles_per_unit, t_begin, t_end);
      if (ti.first_of_unit) ++n_unit;
      if ((i & 1) != grp) continue;
      const int ks = n_unit % KVS;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const int sub = row / p.RB;
      const int pos = ti.p0 + row - sub * p.RB;
      const bool row_ok = sub < ti.nheads && pos < S;
      const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(grp * 256);
      tc::mbar_wait(s_full + grp, ph);
      tc::tcgen05_fence_after();
      auto key_mask = [&](int c) -> unsigned {
        unsigned km = kvalid[ks * 8 + (c >> 5)];
        if (p.causal) km &= pos < c ? 0u : (pos - c >= 31 ? 0xffffffffu : ((2u << (pos - c)) - 1u));
        return km;
      };
      // pass A: row maximum of the raw scores over the valid keys (TMEM loads one chunk ahead)
      float m = -INFINITY;
      {
        uint32_t va[32], vb[32];
        auto row_max = [&](const uint32_t (&v)[32], int c) {
          const unsigned km = key_mask(c);
          if (km == 0xffffffffu) {
            // four independent chains instead of one 32-deep dependency
            float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[j]));
              m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
            }
            m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((km >> j) & 1u) m = fmaxf(m, __uint_as_float(v[j]));
          }
        };
        tc::tmem_ld_32x32(t_row, va);
        for (int c = 0; c < NK; c += 64) {
          tc::tmem_ld_wait();
          if (c + 32 < NK) tc::tmem_ld_32x32(t_row + c + 32, vb);
          row_max(va, c);
          if (c + 32 < NK) {
            tc::tmem_ld_wait();
            if (c + 64 < NK) tc::tmem_ld_32x32(t_row + c + 64, va);
            row_max(vb, c + 32);
          }
        }
      }
      const float ms = (m == -INFINITY || !(m == m)) ? 0.f : m * p.scale_log2;
      // pass B: probabilities -> packed bf16 back into TMEM, row sum
      float sum = 0.f;
      {
        uint32_t va[32], vb[32];
        auto probs = [&](const uint32_t (&v)[32], int c) {
          const unsigned km = key_mask(c);
          uint32_t pk[16];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          if (km == 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -ms));
              const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -ms));
              const float e2 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 2]), p.scale_log2, -ms));
              const float e3 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 3]), p.scale_log2, -ms));
              s0 += e0; s1 += e1; s2 += e2; s3 += e3;
              pk[j] = pack2(e0, e1);
              pk[j + 1] = pack2(e2, e3);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -ms));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -ms));
              if (!((km >> (2 * j)) & 1u)) e0 = 0.f;
              if (!((km >> (2 * j + 1)) & 1u)) e1 = 0.f;
              s0 += e0; s1 += e1;
              pk[j] = pack2(e0, e1);
            }
          }
          sum += (s0 + s1) + (s2 + s3);
          // P overwrites the (already consumed) low columns of S: keys [c, c+32) -> columns [c/2, c/2+16)
          tc::tmem_st_32x16(t_row + (c >> 1), pk);
        };
        tc::tmem_ld_32x32(t_row, va);
        for (int c = 0; c < NK; c += 64) {
          tc::tmem_ld_wait();
          if (c + 32 < NK) tc::tmem_ld_32x32(t_row + c + 32, vb);
          probs(va, c);
          if (c + 32 < NK) {
            tc::tmem_ld_wait();
            if (c + 64 < NK) tc::tmem_ld_32x32(t_row + c + 64, va);
            probs(vb, c + 32);
          }
        }
      }
      tc::tmem_st_wait();
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_relaxed(p_ready + grp);  // tcgen05.st waited for + fenced above
      // epilogue: two 64-column halves, both TMEM loads of a half in flight together
      tc::mbar_wait(o_full + grp, ph);
      tc::tcgen05_fence_after();
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      __nv_bfloat16* orow =
          p.out + (size_t)((int64_t)ti.b * S + pos) * p.ldo + (size_t)(ti.kvh * group + ti.h0 + sub) * kHD;
#pragma unroll 1
      for (int c = 0; c < kHD; c += 64) {
        uint32_t v[2][32];
        tc::tmem_ld_32x32(t_row + kOCol + c, v[0]);
        tc::tmem_ld_32x32(t_row + kOCol + c + 32, v[1]);
        tc::tmem_ld_wait();
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[h * 4 + j] =
                  make_uint4(pack2(__uint_as_float(v[h][j * 8]) * inv, __uint_as_float(v[h][j * 8 + 1]) * inv),
                             pack2(__uint_as_float(v[h][j * 8 + 2]) * inv, __uint_as_float(v[h][j * 8 + 3]) * inv),
                             pack2(__uint_as_float(v[h][j * 8 + 4]) * inv, __uint_as_float(v[h][j * 8 + 5]) * inv),
                             pack2(__uint_as_float(v[h][j * 8 + 6]) * inv, __uint_as_float(v[h][j * 8 + 7]) * inv));
        }
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_relaxed(slot_free + grp);  // TMEM reads only: do not wait for the output stores
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int KVS>
__global__ __launch_bounds__(kPThreads, 1) void attention_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                               const __grid_constant__ CUtensorMap tmKV,
                                                                               const AttnParams p, int total_tiles,
                                                                               int tiles_per_unit) {
  attention_tc_persistent_body<KVS>(tmQ, tmKV, p, total_tiles, tiles_per_unit);
}

// Same kernel capped at 104 registers (a few hundred bytes of spills in the softmax warps): 384 x 104 leaves
// room on the SM for the co-resident scan CTA of QueryPipeline (ivf_scan_ring.cu), as the GEMMs do.
template <int KVS>
__global__ __maxnreg__(104) void attention_tc_persistent_small_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmKV,
                                                                      const AttnParams p, int total_tiles,
                                                                      int tiles_per_unit) {
  attention_tc_persistent_body<KVS>(tmQ, tmKV, p, total_tiles, tiles_per_unit);
}

}  // namespace

bool attention_tc_supported(int S) { return S >= 1 && S <= 512; }

// qkv: bf16 [B*S, ld] = [q heads | k heads | v heads] x 128; out: bf16 [B*S, ldo] = q heads x 128
void attention_tc(const void* qkv, int ld, const int* mask, void* out, int ldo, int B, int S, int nh, int nkv,
                  int causal, float scale_log2, int sms, bool persistent, cudaStream_t st) {
  ABSB_CHECK(attention_tc_supported(S), ABSB_ERR_INVALID, "attention_tc: S=%d outside [1,512]", S);
  AttnParams p{};
  p.S = S;
  p.nh = nh;
  p.nkv = nkv;
  p.causal = causal;
  p.NK = (int)ceil_div(S, 16) * 16;
  const int group = nh / nkv;
  int blocks;
  if (S >= 128) {
    p.RB = 128;
    p.heads_per_blk = 1;
    p.blks_per_head = (int)ceil_div(S, 128);
    blocks = group * p.blks_per_head;
  } else {
    p.RB = (int)ceil_div(S, 8) * 8;
    p.heads_per_blk = std::min(group, 128 / p.RB);
    p.blks_per_head = 1;
    blocks = (int)ceil_div(group, p.heads_per_blk);
  }
  p.scale_log2 = scale_log2;
  p.mask = mask;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  const int64_t T = (int64_t)B * S;
  if (S > 256) {
    // two key blocks of <= 256 per q tile (attention_tc_long_kernel)
    const CUtensorMap tmQl = make_tmap_bf16(qkv, T, ld, ld, 128);
    const CUtensorMap tmKVl = make_tmap_bf16(qkv, T, ld, ld, kLongKB);
    const size_t smem_l = 1024 + 2 * 128 * 128 + 4 * (size_t)kLongKB * 128 + (512 + 512) * 4 + 4 * 8 + 16;
    static bool configured_l = false;
    if (!configured_l) {
      ABSB_CUDA(cudaFuncSetAttribute(attention_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      prefer_max_shared(attention_tc_long_kernel);
      configured_l = true;
    }
    dim3 grid((unsigned)blocks, (unsigned)nkv, (unsigned)B);
    launch_pdl(attention_tc_long_kernel, grid, dim3(kLongThreads), smem_l, st, tmQl, tmKVl, p);
    return;
  }
  const CUtensorMap tmQ = make_tmap_bf16(qkv, T, ld, ld, p.RB);
  const CUtensorMap tmKV = make_tmap_bf16(qkv, T, ld, ld, p.NK);
  if (persistent) {
    const int tiles_per_unit = blocks;
    const int total = tiles_per_unit * nkv * B;
    const int kvs = p.NK <= 128 ? 2 : 1;
    const size_t smem_p = 1024 + 2 * 32768 + (size_t)kvs * 4 * p.NK * 128 + 2 * 8 * 4 + 16 * 8 + 16;
    static bool configured_p = false;
    if (!configured_p) {
      ABSB_CUDA(cudaFuncSetAttribute(attention_tc_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      ABSB_CUDA(cudaFuncSetAttribute(attention_tc_persistent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      ABSB_CUDA(cudaFuncSetAttribute(attention_tc_persistent_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      ABSB_CUDA(cudaFuncSetAttribute(attention_tc_persistent_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      prefer_max_shared(attention_tc_persistent_kernel<1>);
      prefer_max_shared(attention_tc_persistent_kernel<2>);
      prefer_max_shared(attention_tc_persistent_small_kernel<1>);
      prefer_max_shared(attention_tc_persistent_small_kernel<2>);
      configured_p = true;
    }
    const int grid_p = std::max(1, std::min(total, sms));
    // co-resident mode (the GEMMs run under a reduced shared-memory budget): register-capped twin, if its
    // shared memory also fits beside the scan CTA
    const bool small = gemm_coresident_mode() && smem_p <= 161 * 1024;
    const dim3 g(grid_p), b(kPThreads);
    if (kvs == 2) {
      if (small) launch_pdl(attention_tc_persistent_small_kernel<2>, g, b, smem_p, st, tmQ, tmKV, p, total, tiles_per_unit);
      else launch_pdl(attention_tc_persistent_kernel<2>, g, b, smem_p, st, tmQ, tmKV, p, total, tiles_per_unit);
    } else {
      if (small) launch_pdl(attention_tc_persistent_small_kernel<1>, g, b, smem_p, st, tmQ, tmKV, p, total, tiles_per_unit);
      else launch_pdl(attention_tc_persistent_kernel<1>, g, b, smem_p, st, tmQ, tmKV, p, total, tiles_per_unit);
    }
    return;
  }
  const size_t smem = 1024 + 2 * 128 * 128 + 4 * (size_t)p.NK * 128 + 256 * 4 + 4 * 8 + 16;
  static bool configured = false;
  if (!configured) {
    ABSB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)blocks, (unsigned)nkv, (unsigned)B);
  launch_pdl(attention_tc_kernel, grid, dim3(kThreads), smem, st, tmQ, tmKV, p);
}

}  // namespace absb
