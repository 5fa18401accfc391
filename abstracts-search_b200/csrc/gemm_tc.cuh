// gemm_tc.cuh — host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace absb {

enum GemmEpilogue {
  EPI_BF16_BIAS = 0,    // out bf16 [M,N]   = acc (+ bias[N])
  EPI_F32_BIAS = 1,     // out f32  [M,N]   = acc (+ bias[N])
  EPI_F32_ADD = 2,      // out f32  [M,N]  += acc                      (residual stream)
  EPI_SWIGLU_BF16 = 3,  // out bf16 [M,N/2] = silu(gate) * up, weight rows interleaved per 256-row
                        //                    tile: [128 gate rows | 128 up rows]
  EPI_ARGMAX = 5,  // no C at all: per (row, 128-column span) the best score and its column, for
                   // quantizer.assign (k = 1) without ever writing the [M, nlist] score matrix
  EPI_BF16_BIAS_ROPE = 4,  // EPI_BF16_BIAS + rotate-half RoPE on the 128-wide heads that start below
                           // rope.cols (the q and k heads of the fused QKV projection)
};

// RoPE side input of EPI_BF16_BIAS_ROPE: cs[pos, i] = (cos, sin) of pos * inv_freq[i], i < 64;
// the position of GEMM row r is r % S.
struct GemmRope {
  const float2* cs = nullptr;  // (cos, sin) table, PAIR-major: cs[i * ld + pos], i < head_dim / 2
  int S = 1;                   // sequence length: position of GEMM row r is r % S
  int cols = 0;                // leading output columns that are rotated (q and k heads)
  int ld = 0;                  // positions per table row (>= S)
};

constexpr int kMaxGemmSegs = 8;

// K-segment table: the reduction runs over nseg segments of length K; segment s reads A columns
// [a_off[s], a_off[s]+K) and B columns [b_off[s], b_off[s]+K).  Used by the split-bf16 GEMM.
struct GemmSegs {
  int nseg = 0;
  int a_off[kMaxGemmSegs] = {0};
  int b_off[kMaxGemmSegs] = {0};
  int64_t a_cols = 0, b_cols = 0;  // full row length of A and B
};

// C[M,N] = A[M,K] * B[N,K]^T; A, B bf16 with row pitches lda, ldb (elements).
void gemm_bf16_tc(int epi, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                  int64_t ldc, const float* bias, const GemmSegs* segs, int sms, cudaStream_t st,
                  const GemmRope* rope = nullptr);

// Row-wise arg-max of A * B^T from split-bf16 operands (see gemm_split3_f32): out_idx[r] = smallest
// column holding the row maximum, out_score[r] (optional) that maximum.  ws_max / ws_idx are scratch
// of argmax_partials_per_row(N) entries per row.
int argmax_partials_per_row(int N);
void gemm_split3_argmax(int M, int N, int K, const void* A3, const void* B3, float* ws_max, int* ws_idx,
                        long long* out_idx, float* out_score, int sms, cudaStream_t st);

// TMA descriptor of a bf16 row-major matrix [rows, cols] (pitch ld elements): boxes of 64 columns x
// box_rows rows, SWIZZLE_128B.
CUtensorMap make_tmap_bf16(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

// attention_tc.cu: tcgen05 attention for S <= 256 (qkv bf16 [B*S, ld], out bf16 [B*S, ldo])
bool attention_tc_supported(int S);
void attention_tc(const void* qkv, int ld, const int* mask, void* out, int ldo, int B, int S, int nh, int nkv,
                  int causal, float scale_log2, int sms, bool persistent, cudaStream_t st);

// test hook: 0 auto; 1 = 1 CTA x BN 256; 2 = CTA pair x BN 256; 3 = CTA pair x BN 192
void gemm_set_variant(int v);
void gemm_set_ksplit(int n);  // 0 auto, 1 off, n forced (residual-add epilogue only)
// Shared memory a GEMM CTA may take (absb_gemm_set_smem_budget); a budget below the full 227 KB means the
// encoder shares its SMs with a co-resident scan CTA.
void gemm_set_smem_budget(int bytes);
bool gemm_coresident_mode();

// fp32 [rows,K] -> bf16 [rows, 3K] = [hi | mid | lo]
void split3_bf16(int64_t rows, int K, const float* x, void* out, cudaStream_t st);
// S[M,N] (fp32, pitch lds) = A * B^T from split operands A3 [M,3K], B3 [N,3K]
void gemm_split3_f32(int M, int N, int K, const void* A3, const void* B3, float* S, int64_t lds, int sms,
                     cudaStream_t st);

}  // namespace absb
