// attention_tc.cu — bidirectional GQA attention of the stella/Qwen2 encoder on the 5th-generation
// tensor cores, for sequences of up to 256 tokens (the bulk-encode batches of
// `sidecar-search build -b 32`, /root/reference/Makefile:65, and app.py's queries, README.md:28;
// SURVEY §2c E5).  Longer sequences (up to max_seq_len 512) take the mma.sync kernel in encoder.cu.
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

}  // namespace

bool attention_tc_supported(int S) { return S >= 1 && S <= 256; }

// qkv: bf16 [B*S, ld] = [q heads | k heads | v heads] x 128; out: bf16 [B*S, ldo] = q heads x 128
void attention_tc(const void* qkv, int ld, const int* mask, void* out, int ldo, int B, int S, int nh, int nkv,
                  int causal, float scale_log2, cudaStream_t st) {
  ABSB_CHECK(attention_tc_supported(S), ABSB_ERR_INVALID, "attention_tc: S=%d outside [1,256]", S);
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
  const CUtensorMap tmQ = make_tmap_bf16(qkv, T, ld, ld, p.RB);
  const CUtensorMap tmKV = make_tmap_bf16(qkv, T, ld, ld, p.NK);
  const size_t smem = 1024 + 2 * 128 * 128 + 4 * (size_t)p.NK * 128 + 256 * 4 + 4 * 8 + 16;
  static bool configured = false;
  if (!configured) {
    ABSB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)blocks, (unsigned)nkv, (unsigned)B);
  attention_tc_kernel<<<grid, kThreads, smem, st>>>(tmQ, tmKV, p);
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb
