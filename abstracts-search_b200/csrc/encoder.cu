// encoder.cu — stella_en_1.5B_v5 forward pass: the arithmetic behind SentenceTransformer.encode()
// as called by `sidecar-search build -b 32` (/root/reference/Makefile:65, README.md:60) and by the
// app.py query loop (/root/reference/README.md:28).  SURVEY.md §8(a) a1–a3, §2c E1–E8.
//
//   embed -> 28 x { RMSNorm, QKV(+bias), RoPE(theta=1e6), bidirectional GQA attention with
//   key-padding mask, O-proj + residual, RMSNorm, SwiGLU FFN + residual } -> final RMSNorm ->
//   masked mean pool -> Dense(1536 -> 1024, bias) -> optional L2 normalise.
//
// Numerics: weights and GEMM operands bf16, fp32 accumulation (TMEM), fp32 residual stream, fp32
// softmax / norm statistics.  Every Linear runs on tcgen05 (gemm_tc.cu) with its epilogue fused:
// bias (QKV, Dense), SiLU(gate)*up (FFN in), residual add (O-proj, FFN out).
#include <cuda_bf16.h>

#include <cmath>
#include <cstring>
#include <map>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace absb {

namespace {

using bf16 = __nv_bfloat16;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ small kernels -----------
__global__ void embed_kernel(int64_t T, int H, int vocab, const long long* __restrict__ ids,
                             const bf16* __restrict__ table, float* __restrict__ h) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= T) return;
  long long id = ids[w];
  if (id < 0 || id >= vocab) id = 0;
  const uint2* src = reinterpret_cast<const uint2*>(table + (size_t)id * H);  // 4 bf16
  float4* dst = reinterpret_cast<float4*>(h + (size_t)w * H);
  for (int j = lane; j < H / 4; j += 32) {
    const uint2 v = __ldg(src + j);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    dst[j] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One warp per token: y = x * rsqrt(mean(x^2) + eps) * w.  OUT = bf16 (GEMM operand) or float.
// Rows of up to 32 * 4 * kRmsRegs floats (1536 = stella's hidden size) are held in registers: all
// loads of a row are in flight at once and the row is read only once.
constexpr int kRmsRegs = 12;

template <typename OUT>
__device__ __forceinline__ void rms_store(OUT* __restrict__ y, size_t t, int H, int j, const float4& v, float rinv,
                                          const float4& g) {
  const float o0 = v.x * rinv * g.x, o1 = v.y * rinv * g.y, o2 = v.z * rinv * g.z, o3 = v.w * rinv * g.w;
  if constexpr (sizeof(OUT) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), b = __floats2bfloat162_rn(o2, o3);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(y + t * H)[j] = pk;
  } else {
    reinterpret_cast<float4*>(y + t * H)[j] = make_float4(o0, o1, o2, o3);
  }
}

template <typename OUT>
__global__ __launch_bounds__(256, 2) void rmsnorm_kernel(int64_t T, int H, float eps, const float* __restrict__ x,
                                                         const float* __restrict__ w, OUT* __restrict__ y,
                                                         float* __restrict__ rinv_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // grid-stride over rows (the launch caps the grid at two CTAs per SM): a warp has the loads of its NEXT row in
  // flight while it reduces, scales and stores the current one, so the SM never drains between rows
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (t >= T) return;
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int n4 = H / 4;
  if (n4 <= 32 * kRmsRegs) {
    float4 v[kRmsRegs], nx[kRmsRegs];
    auto load_row = [&](float4 (&r)[kRmsRegs], int64_t row) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * H);
#pragma unroll
      for (int i = 0; i < kRmsRegs; ++i) {
        const int j = lane + 32 * i;
        r[i] = j < n4 ? xr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_row(v, t);
    for (;;) {
      const int64_t tn = t + nw;
      const bool more = tn < T;
      if (more) load_row(nx, tn);
      // same summation order as the streaming path below: per lane ascending j, then the warp tree
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < kRmsRegs; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      ss = warp_sum(ss);
      const float rinv = rsqrtf(ss / (float)H + eps);
      if (rinv_out && lane == 0) rinv_out[t] = rinv;
      if (y) {
#pragma unroll
        for (int i = 0; i < kRmsRegs; ++i) {
          const int j = lane + 32 * i;
          if (j < n4) rms_store(y, (size_t)t, H, j, v[i], rinv, __ldg(wr + j));
        }
      }
      if (!more) break;
#pragma unroll
      for (int i = 0; i < kRmsRegs; ++i) v[i] = nx[i];
      t = tn;
    }
    return;
  }
  for (; t < T; t += nw) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)t * H);
    float ss = 0.f;
    for (int j = lane; j < n4; j += 32) {
      const float4 v = xr[j];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float rinv = rsqrtf(ss / (float)H + eps);
    if (rinv_out && lane == 0) rinv_out[t] = rinv;
    if (y)
      for (int j = lane; j < n4; j += 32) rms_store(y, (size_t)t, H, j, xr[j], rinv, __ldg(wr + j));
  }
}

// Masked mean pool of the final-normed hidden states: pooled[b, j] = w[j] * sum_s m[b,s] *
// h[b,s,j] * rinv[b,s] / max(sum_s m[b,s], 1e-9)   (sentence-transformers Pooling, mean mode).
__global__ void pool_kernel(int S, int H, const float* __restrict__ h, const float* __restrict__ rinv,
                            const int* __restrict__ mask, const float* __restrict__ w, bf16* __restrict__ pooled) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float acc = 0.f, cnt = 0.f;
  for (int s = 0; s < S; ++s) {
    const int64_t t = (int64_t)b * S + s;
    if (mask[t]) {
      acc += h[(size_t)t * H + j] * rinv[t];
      cnt += 1.f;
    }
  }
  pooled[(size_t)b * H + j] = __float2bfloat16_rn(acc * w[j] / fmaxf(cnt, 1e-9f));
}

// One warp per row: x /= max(||x||_2, 1e-12)   (torch.nn.functional.normalize)
__global__ void l2norm_kernel(int64_t B, int E, float* __restrict__ x) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= B) return;
  float* xr = x + (size_t)r * E;
  float ss = 0.f;
  for (int j = lane; j < E; j += 32) ss += xr[j] * xr[j];
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int j = lane; j < E; j += 32) xr[j] *= inv;
}

// ------------------------------------------------------------------ attention ---------------
// Flash-style attention on mma.sync m16n8k16 (bf16 in, fp32 accumulate), head_dim 128; 64-key tiles
// staged in shared memory and shared by the q heads of one kv head.
constexpr int kHD = 128;
constexpr int kKT = 64;         // keys per tile
constexpr int kRowPad = kHD + 8;  // 272-byte rows: conflict-free fragment loads and ldmatrix

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kAttnWarps = 12;

// One CTA per (batch, kv head, block of 12 work units); a unit = 16 query rows of one of the q heads
// that share this kv head (GQA), one warp per unit, so K/V tiles are staged once for all of them.
__global__ __launch_bounds__(kAttnWarps * 32) void attention_kernel(const bf16* __restrict__ qkv, int ld,
                                                                  const int* __restrict__ mask, bf16* __restrict__ out,
                                                                  int ldo, int S, int nh, int nkv, int causal,
                                                                  float scale_log2) {
  __shared__ __align__(16) bf16 sK[kKT][kRowPad];
  __shared__ __align__(16) bf16 sV[kKT][kRowPad];
  __shared__ float sBias[kKT];
  pdl_trigger();
  pdl_wait();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kvh = blockIdx.y, b = blockIdx.z;
  const int group = nh / nkv;
  const int nblk = (S + 15) / 16;
  const int unit = blockIdx.x * kAttnWarps + warp;
  const bool active = unit < group * nblk;
  const int head = kvh * group + (active ? unit / nblk : 0);
  const int rb = active ? unit % nblk : 0;
  const int64_t tok0 = (int64_t)b * S;
  const bf16* Qb = qkv + (size_t)head * kHD;
  const bf16* Kb = qkv + (size_t)(nh + kvh) * kHD;
  const bf16* Vb = qkv + (size_t)(nh + nkv + kvh) * kHD;

  const int r0 = rb * 16 + g, r1 = r0 + 8;
  uint32_t qa[kHD / 16][4];
#pragma unroll
  for (int ks = 0; ks < kHD / 16; ++ks) {
    const int c = ks * 16 + t4 * 2;
    const bool ok0 = active && r0 < S, ok1 = active && r1 < S;
    qa[ks][0] = ok0 ? *reinterpret_cast<const uint32_t*>(Qb + (size_t)(tok0 + r0) * ld + c) : 0u;
    qa[ks][1] = ok1 ? *reinterpret_cast<const uint32_t*>(Qb + (size_t)(tok0 + r1) * ld + c) : 0u;
    qa[ks][2] = ok0 ? *reinterpret_cast<const uint32_t*>(Qb + (size_t)(tok0 + r0) * ld + c + 8) : 0u;
    qa[ks][3] = ok1 ? *reinterpret_cast<const uint32_t*>(Qb + (size_t)(tok0 + r1) * ld + c + 8) : 0u;
  }

  float o[kHD / 8][4];
#pragma unroll
  for (int i = 0; i < kHD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // keys needed by this CTA: all of them, or (causal) up to the last row of its last unit
  int kmax_cta = S;
  if (causal) {
    const int last_unit = min((int)(blockIdx.x + 1) * kAttnWarps, group * nblk) - 1;
    const int first_unit = blockIdx.x * kAttnWarps;
    // units of several heads may share the CTA: the largest row block among them bounds the keys
    int rb_max = 0;
    for (int u = first_unit; u <= last_unit; ++u) rb_max = max(rb_max, u % nblk);
    kmax_cta = min(S, rb_max * 16 + 16);
  }
  const int kmax_warp = causal ? min(S, rb * 16 + 16) : S;
  for (int k0 = 0; k0 < kmax_cta; k0 += kKT) {
    __syncthreads();  // previous tile fully consumed
    // stage K, V tiles: 64 rows x 16 chunks of 16 bytes
    for (int i = threadIdx.x; i < kKT * (kHD / 8); i += blockDim.x) {
      const int r = i / (kHD / 8), c = (i % (kHD / 8)) * 8;
      const int key = k0 + r;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (key < S) {
        kv = *reinterpret_cast<const uint4*>(Kb + (size_t)(tok0 + key) * ld + c);
        vv = *reinterpret_cast<const uint4*>(Vb + (size_t)(tok0 + key) * ld + c);
      }
      *reinterpret_cast<uint4*>(&sK[r][c]) = kv;
      *reinterpret_cast<uint4*>(&sV[r][c]) = vv;
    }
    if (threadIdx.x < kKT) {
      const int key = k0 + threadIdx.x;
      sBias[threadIdx.x] = (key < S && mask[tok0 + key] != 0) ? 0.f : -INFINITY;
    }
    __syncthreads();
    if (!active || k0 >= kmax_warp) continue;
    // 8-key column blocks of this tile that hold any key at all (short sequences: skip the padding)
    const int nb_valid = min(kKT, S - k0 + 7) / 8;

    // S = Q K^T for this warp's 16 rows x 64 keys
    float s[kKT / 8][4];
#pragma unroll
    for (int nb = 0; nb < kKT / 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
      if (nb < nb_valid) {
#pragma unroll
        for (int ks = 0; ks < kHD / 16; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sK[nb * 8 + g][ks * 16 + t4 * 2]);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sK[nb * 8 + g][ks * 16 + t4 * 2 + 8]);
          mma_bf16_16816(s[nb], qa[ks], b0, b1);
        }
      }
    }
    // scale, mask, running max
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < kKT / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kl = nb * 8 + t4 * 2 + e;
        float bias = sBias[kl];
        float v0 = s[nb][e] * scale_log2 + bias;
        float v1 = s[nb][2 + e] * scale_log2 + bias;
        if (causal) {
          if (k0 + kl > r0) v0 = -INFINITY;
          if (k0 + kl > r1) v1 = -INFINITY;
        }
        s[nb][e] = v0;
        s[nb][2 + e] = v1;
        mx0 = fmaxf(mx0, v0);
        mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 2));
    const float nm0 = fmaxf(m0, mx0), nm1 = fmaxf(m1, mx1);
    const float sm0 = nm0 == -INFINITY ? 0.f : nm0, sm1 = nm1 == -INFINITY ? 0.f : nm1;
    const float a0 = exp2f(m0 - sm0), a1 = exp2f(m1 - sm1);  // m = -inf -> 0
    m0 = nm0;
    m1 = nm1;
    l0 *= a0;
    l1 *= a1;
#pragma unroll
    for (int i = 0; i < kHD / 8; ++i) {
      o[i][0] *= a0; o[i][1] *= a0;
      o[i][2] *= a1; o[i][3] *= a1;
    }
#pragma unroll
    for (int nb = 0; nb < kKT / 8; ++nb) {
      s[nb][0] = exp2f(s[nb][0] - sm0);
      s[nb][1] = exp2f(s[nb][1] - sm0);
      s[nb][2] = exp2f(s[nb][2] - sm1);
      s[nb][3] = exp2f(s[nb][3] - sm1);
      l0 += s[nb][0] + s[nb][1];
      l1 += s[nb][2] + s[nb][3];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < kKT / 16; ++kk) {
      if (kk * 2 >= nb_valid) continue;  // 16-key block entirely past the sequence: P = 0 there
      uint32_t pa[4];
      pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const int mi = lane >> 3, mr = lane & 7;
#pragma unroll
      for (int dn = 0; dn < kHD / 16; ++dn) {
        const bf16* addr = &sV[kk * 16 + (mi & 1) * 8 + mr][dn * 16 + (mi >> 1) * 8];
        uint32_t v0, v1, v2, v3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                     : "r"((uint32_t)__cvta_generic_to_shared(addr)));
        mma_bf16_16816(o[2 * dn], pa, v0, v1);
        mma_bf16_16816(o[2 * dn + 1], pa, v2, v3);
      }
    }
  }
  if (!active) return;
  l0 += __shfl_xor_sync(kFull, l0, 1);
  l0 += __shfl_xor_sync(kFull, l0, 2);
  l1 += __shfl_xor_sync(kFull, l1, 1);
  l1 += __shfl_xor_sync(kFull, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  bf16* Ob = out + (size_t)head * kHD;
#pragma unroll
  for (int nb = 0; nb < kHD / 8; ++nb) {
    const int c = nb * 8 + t4 * 2;
    if (r0 < S) *reinterpret_cast<uint32_t*>(Ob + (size_t)(tok0 + r0) * ldo + c) = pack2(o[nb][0] * i0, o[nb][1] * i0);
    if (r1 < S) *reinterpret_cast<uint32_t*>(Ob + (size_t)(tok0 + r1) * ldo + c) = pack2(o[nb][2] * i1, o[nb][3] * i1);
  }
}

// ------------------------------------------------------------------ weights -----------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// normal(0, std) from a counter-based hash (Box-Muller); OUT bf16 or float
template <typename OUT>
__global__ void init_normal_kernel(int64_t n, uint64_t key, float std, float mean, OUT* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t w = mix64(key + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    const float u1 = ((float)(uint32_t)(w >> 40) + 1.0f) * (1.0f / 16777217.0f);  // (0,1]
    const float u2 = (float)(uint32_t)((w >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    const float v = mean + std * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    if constexpr (sizeof(OUT) == 2) out[i] = __float2bfloat16_rn(v);
    else out[i] = v;
  }
}

__global__ void f32_to_bf16_kernel(int64_t n, const float* __restrict__ in, bf16* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(int64_t n, const bf16* __restrict__ in, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

int blocks_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), 148 * 16)); }

}  // namespace

// ------------------------------------------------------------------ Encoder -----------------
struct Layer {
  DBuf<float> ln1, ln2, bqkv;
  DBuf<bf16> wqkv, wo, wgu, wd;
};

// A named view into device storage, used by load/get_weight.  Row `r` of the logical [rows, cols]
// parameter lives at base + map_row(r) * cols.
struct ParamRef {
  void* base = nullptr;
  bool is_bf16 = false;
  int64_t rows = 0, cols = 0;
  int64_t row0 = 0;       // plain offset (q/k/v inside wqkv, biases inside bqkv)
  int interleave = -1;    // -1 none; 0 = gate rows, 1 = up rows of the [128 gate | 128 up] tiling
};

struct Encoder {
  absb_enc_config cfg;
  int device;
  DeviceProps props;
  cudaStream_t own_stream = nullptr;
  DBuf<bf16> embed, dense_w;
  DBuf<float> final_norm, dense_b;
  std::vector<Layer> layers;
  DBuf<float2> rope_cs;
  bool weights_ready = false;
  int attention_impl = 1;  // 1 = tcgen05: persistent kernel for S <= 256, two-key-block kernel for 257-512;
                           // 2 = one-tile-per-CTA tcgen05 kernel (same two-block kernel above 256); 0 = always mma.sync

  // activations (sized for cap_tokens)
  int64_t cap_tokens = 0, cap_batch = 0;
  DBuf<float> h, rinv, emb;
  DBuf<bf16> xn, qkv, ao, act, pooled;
  DBuf<long long> ids_ws;
  DBuf<int> mask_ws;
  int last_B = 0, last_S = 0;
  double last_flops = 0;
  int64_t last_launches = 0;

  // optional per-phase timing (absb_enc_set_profile)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  std::vector<int> ev_kind;  // 0 gemm, 1 attention, 2 other
  double prof_ms[3] = {0, 0, 0};
  double prof_gemm_flops = 0;
  int64_t prof_forwards = 0;

  Encoder(const absb_enc_config& c, int dev) : cfg(c), device(dev) {
    ABSB_CHECK(c.head_dim == kHD, ABSB_ERR_UNSUPPORTED, "head_dim must be %d (got %d)", kHD, c.head_dim);
    ABSB_CHECK(c.num_heads > 0 && c.num_kv_heads > 0 && c.num_heads % c.num_kv_heads == 0, ABSB_ERR_INVALID, "bad head counts");
    ABSB_CHECK(c.hidden_size > 0 && c.hidden_size % 32 == 0, ABSB_ERR_UNSUPPORTED, "hidden_size %% 32 != 0");
    ABSB_CHECK(c.intermediate_size > 0 && c.intermediate_size % 128 == 0, ABSB_ERR_UNSUPPORTED, "intermediate_size %% 128 != 0");
    ABSB_CHECK(c.embed_dim > 0 && c.embed_dim % 32 == 0, ABSB_ERR_UNSUPPORTED, "embed_dim %% 32 != 0");
    ABSB_CHECK(c.vocab_size > 0 && c.num_layers > 0 && c.max_seq_len > 0, ABSB_ERR_INVALID, "bad config");
    props = device_props(dev);
    DeviceGuard g(dev);
    ABSB_CUDA(cudaStreamCreate(&own_stream));
    // every kernel of a forward may run while a co-resident scan CTA holds 64 KB of the SM (QueryPipeline)
    prefer_max_shared(embed_kernel);
    prefer_max_shared(rmsnorm_kernel<bf16>);
    prefer_max_shared(rmsnorm_kernel<float>);
    prefer_max_shared(pool_kernel);
    prefer_max_shared(l2norm_kernel);
    prefer_max_shared(attention_kernel);
    const int H = c.hidden_size, I = c.intermediate_size, QKV = qkv_dim();
    embed.alloc_exact((size_t)c.vocab_size * H);
    final_norm.alloc_exact(H);
    dense_w.alloc_exact((size_t)c.embed_dim * H);
    dense_b.alloc_exact(c.embed_dim);
    layers.resize(c.num_layers);
    for (auto& L : layers) {
      L.ln1.alloc_exact(H);
      L.ln2.alloc_exact(H);
      L.bqkv.alloc_exact(QKV);
      L.wqkv.alloc_exact((size_t)QKV * H);
      L.wo.alloc_exact((size_t)H * c.num_heads * kHD);
      L.wgu.alloc_exact((size_t)2 * I * H);
      L.wd.alloc_exact((size_t)H * I);
    }
    // RoPE table, computed the way transformers does it in fp32: inv_freq = 1/theta^(2i/hd),
    // angle = pos * inv_freq
    const int half = kHD / 2;
    std::vector<float2> cs((size_t)c.max_seq_len * half);
    for (int p = 0; p < c.max_seq_len; ++p)
      for (int i = 0; i < half; ++i) {
        const float inv_freq = 1.0f / powf(c.rope_theta, (float)(2 * i) / (float)kHD);
        const float ang = (float)p * inv_freq;
        cs[(size_t)i * c.max_seq_len + p] = make_float2(cosf(ang), sinf(ang));  // pair-major (gemm_tc.cuh)
      }
    rope_cs.alloc_exact(cs.size());
    ABSB_CUDA(cudaMemcpy(rope_cs.p, cs.data(), cs.size() * sizeof(float2), cudaMemcpyHostToDevice));
  }

  ~Encoder() {
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

  int qkv_dim() const { return (cfg.num_heads + 2 * cfg.num_kv_heads) * kHD; }

  // ---- parameter naming (Hugging Face Qwen2Model + sentence-transformers Dense) ----------
  ParamRef find(const std::string& name_in) {
    std::string name = name_in;
    if (name.rfind("model.", 0) == 0) name = name.substr(6);
    const int H = cfg.hidden_size, I = cfg.intermediate_size, nh = cfg.num_heads, nkv = cfg.num_kv_heads;
    ParamRef r;
    if (name == "embed_tokens.weight") { r.base = embed.p; r.is_bf16 = true; r.rows = cfg.vocab_size; r.cols = H; return r; }
    if (name == "norm.weight") { r.base = final_norm.p; r.rows = 1; r.cols = H; return r; }
    if (name == "dense.weight" || name == "linear.weight") { r.base = dense_w.p; r.is_bf16 = true; r.rows = cfg.embed_dim; r.cols = H; return r; }
    if (name == "dense.bias" || name == "linear.bias") { r.base = dense_b.p; r.rows = 1; r.cols = cfg.embed_dim; return r; }
    int li = -1;
    char rest[128];
    if (sscanf(name.c_str(), "layers.%d.%127s", &li, rest) == 2 && li >= 0 && li < cfg.num_layers) {
      Layer& L = layers[li];
      const std::string s = rest;
      if (s == "input_layernorm.weight") { r.base = L.ln1.p; r.rows = 1; r.cols = H; return r; }
      if (s == "post_attention_layernorm.weight") { r.base = L.ln2.p; r.rows = 1; r.cols = H; return r; }
      if (s == "self_attn.q_proj.weight") { r.base = L.wqkv.p; r.is_bf16 = true; r.rows = nh * kHD; r.cols = H; return r; }
      if (s == "self_attn.k_proj.weight") { r.base = L.wqkv.p; r.is_bf16 = true; r.rows = nkv * kHD; r.cols = H; r.row0 = nh * kHD; return r; }
      if (s == "self_attn.v_proj.weight") { r.base = L.wqkv.p; r.is_bf16 = true; r.rows = nkv * kHD; r.cols = H; r.row0 = (nh + nkv) * kHD; return r; }
      if (s == "self_attn.q_proj.bias") { r.base = L.bqkv.p; r.rows = 1; r.cols = nh * kHD; return r; }
      if (s == "self_attn.k_proj.bias") { r.base = L.bqkv.p; r.rows = 1; r.cols = nkv * kHD; r.row0 = nh * kHD; return r; }
      if (s == "self_attn.v_proj.bias") { r.base = L.bqkv.p; r.rows = 1; r.cols = nkv * kHD; r.row0 = (nh + nkv) * kHD; return r; }
      if (s == "self_attn.o_proj.weight") { r.base = L.wo.p; r.is_bf16 = true; r.rows = H; r.cols = nh * kHD; return r; }
      if (s == "mlp.gate_proj.weight") { r.base = L.wgu.p; r.is_bf16 = true; r.rows = I; r.cols = H; r.interleave = 0; return r; }
      if (s == "mlp.up_proj.weight") { r.base = L.wgu.p; r.is_bf16 = true; r.rows = I; r.cols = H; r.interleave = 1; return r; }
      if (s == "mlp.down_proj.weight") { r.base = L.wd.p; r.is_bf16 = true; r.rows = H; r.cols = I; return r; }
    }
    fail(ABSB_ERR_INVALID, "unknown parameter name '%s'", name_in.c_str());
  }

  // device address of logical row r (bias vectors: rows == 1 and row0 offsets ELEMENTS)
  static char* row_ptr(const ParamRef& p, int64_t r) {
    const size_t es = p.is_bf16 ? 2 : 4;
    if (p.rows == 1) return static_cast<char*>(p.base) + (size_t)p.row0 * es;
    int64_t phys = p.row0 + r;
    if (p.interleave >= 0) phys = (r / 128) * 256 + p.interleave * 128 + (r % 128);
    return static_cast<char*>(p.base) + (size_t)phys * p.cols * es;
  }

  void load_weight(const char* name, const void* data, int dtype, const int64_t* shape, int ndim) {
    ABSB_CHECK(dtype == 0 || dtype == 1, ABSB_ERR_INVALID, "dtype must be 0 (f32) or 1 (bf16)");
    const ParamRef p = find(name);
    int64_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    ABSB_CHECK(numel == p.rows * p.cols, ABSB_ERR_INVALID, "parameter '%s' expects %lld elements, got %lld", name,
               (long long)(p.rows * p.cols), (long long)numel);
    cudaStream_t st = own_stream;
    // stage the source on the device in its own dtype, then convert / scatter row groups
    const size_t src_es = dtype == 1 ? 2 : 4;
    DBuf<unsigned char> stage;
    stage.alloc_exact((size_t)numel * src_es);
    ABSB_CUDA(cudaMemcpyAsync(stage.p, data, (size_t)numel * src_es, cudaMemcpyHostToDevice, st));
    // contiguous runs of logical rows that are also physically contiguous
    const int64_t run = p.interleave >= 0 ? 128 : std::max<int64_t>(p.rows, 1);
    for (int64_t r0 = 0; r0 < std::max<int64_t>(p.rows, 1); r0 += run) {
      const int64_t nr = std::min(run, std::max<int64_t>(p.rows, 1) - r0);
      const int64_t n = nr * p.cols;
      char* dst = row_ptr(p, r0);
      const unsigned char* src = stage.p + (size_t)r0 * p.cols * src_es;
      if (p.is_bf16 && dtype == 0) {
        f32_to_bf16_kernel<<<blocks_for(n), 256, 0, st>>>(n, reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(dst));
      } else if (!p.is_bf16 && dtype == 1) {
        bf16_to_f32_kernel<<<blocks_for(n), 256, 0, st>>>(n, reinterpret_cast<const bf16*>(src), reinterpret_cast<float*>(dst));
      } else {
        ABSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * src_es, cudaMemcpyDeviceToDevice, st));
      }
      ABSB_CUDA(cudaGetLastError());
    }
    ABSB_CUDA(cudaStreamSynchronize(st));
    weights_ready = true;
  }

  void get_weight(const char* name, float* out, int64_t numel) {
    const ParamRef p = find(name);
    ABSB_CHECK(numel == p.rows * p.cols, ABSB_ERR_INVALID, "parameter '%s' has %lld elements, buffer has %lld", name,
               (long long)(p.rows * p.cols), (long long)numel);
    cudaStream_t st = own_stream;
    DBuf<float> stage;
    stage.alloc_exact((size_t)numel);
    const int64_t run = p.interleave >= 0 ? 128 : std::max<int64_t>(p.rows, 1);
    for (int64_t r0 = 0; r0 < std::max<int64_t>(p.rows, 1); r0 += run) {
      const int64_t nr = std::min(run, std::max<int64_t>(p.rows, 1) - r0);
      const int64_t n = nr * p.cols;
      const char* src = row_ptr(p, r0);
      float* dst = stage.p + (size_t)r0 * p.cols;
      if (p.is_bf16) bf16_to_f32_kernel<<<blocks_for(n), 256, 0, st>>>(n, reinterpret_cast<const bf16*>(src), dst);
      else ABSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
      ABSB_CUDA(cudaGetLastError());
    }
    ABSB_CUDA(cudaMemcpyAsync(out, stage.p, (size_t)numel * 4, cudaMemcpyDeviceToHost, st));
    ABSB_CUDA(cudaStreamSynchronize(st));
  }

  void init_random(uint64_t seed, float std) {
    cudaStream_t st = own_stream;
    uint64_t tid = 0;
    auto nb = [&](DBuf<bf16>& b) {
      init_normal_kernel<bf16><<<blocks_for((int64_t)b.cap), 256, 0, st>>>((int64_t)b.cap, mix64(seed ^ (++tid * 0xD1B54A32D192ED03ull)), std, 0.f, b.p);
      ABSB_CUDA(cudaGetLastError());
    };
    auto nf = [&](DBuf<float>& b, float mean, float s) {
      init_normal_kernel<float><<<blocks_for((int64_t)b.cap), 256, 0, st>>>((int64_t)b.cap, mix64(seed ^ (++tid * 0xD1B54A32D192ED03ull)), s, mean, b.p);
      ABSB_CUDA(cudaGetLastError());
    };
    nb(embed);
    nf(final_norm, 1.f, 0.05f);
    nb(dense_w);
    nf(dense_b, 0.f, std);
    for (auto& L : layers) {
      nf(L.ln1, 1.f, 0.05f);
      nf(L.ln2, 1.f, 0.05f);
      nf(L.bqkv, 0.f, std);
      nb(L.wqkv);
      nb(L.wo);
      nb(L.wgu);
      nb(L.wd);
    }
    ABSB_CUDA(cudaStreamSynchronize(st));
    weights_ready = true;
  }

  // ---- forward ------------------------------------------------------------------------
  void ensure_capacity(int B, int S) {
    const int64_t T = (int64_t)B * S;
    if (T > cap_tokens) {
      cap_tokens = T;
      h.alloc_exact((size_t)T * cfg.hidden_size);
      rinv.alloc_exact((size_t)T);
      xn.alloc_exact((size_t)T * cfg.hidden_size);
      qkv.alloc_exact((size_t)T * qkv_dim());
      ao.alloc_exact((size_t)T * cfg.num_heads * kHD);
      act.alloc_exact((size_t)T * cfg.intermediate_size);
    }
    if (B > cap_batch) {
      cap_batch = B;
      pooled.alloc_exact((size_t)B * cfg.hidden_size);
      emb.alloc_exact((size_t)B * cfg.embed_dim);
    }
  }

  struct Span {
    Encoder* e;
    cudaStream_t st;
    int kind;
    bool on;
    cudaEvent_t stop = nullptr;
    Span(Encoder* e_, cudaStream_t st_, int kind_) : e(e_), st(st_), kind(kind_), on(e_->profile) {
      if (!on) return;
      if (e->ev_used == e->ev_pool.size()) {
        cudaEvent_t a, b;
        ABSB_CUDA(cudaEventCreate(&a));
        ABSB_CUDA(cudaEventCreate(&b));
        e->ev_pool.emplace_back(a, b);
      }
      auto& pr = e->ev_pool[e->ev_used++];
      e->ev_kind.push_back(kind);
      stop = pr.second;
      cudaEventRecord(pr.first, st);
    }
    ~Span() {
      if (on) cudaEventRecord(stop, st);
    }
  };

  void gemm(int epi, int M, int N, int K, const void* A, const void* W, void* out, int64_t ldc, const float* bias,
            cudaStream_t st, const GemmRope* rope = nullptr) {
    Span sp(this, st, 0);
    gemm_bf16_tc(epi, M, N, K, A, K, W, K, out, ldc, bias, nullptr, props.sm_count, st, rope);
    last_flops += 2.0 * M * N * K;
    if (profile) prof_gemm_flops += 2.0 * M * N * K;
    ++last_launches;
  }

  void forward_dev(int B, int S, const long long* ids, const int* mask, int normalize, float* out, cudaStream_t st) {
    ABSB_CHECK(weights_ready, ABSB_ERR_STATE, "encoder weights have not been loaded");
    ABSB_CHECK(B >= 1 && S >= 1, ABSB_ERR_INVALID, "B=%d S=%d", B, S);
    ABSB_CHECK(S <= cfg.max_seq_len, ABSB_ERR_INVALID, "sequence length %d exceeds max_seq_len %d", S, cfg.max_seq_len);
    ensure_capacity(B, S);
    const int64_t T = (int64_t)B * S;
    ABSB_CHECK(T < ((int64_t)1 << 31) / 64, ABSB_ERR_INVALID, "too many tokens in one forward");
    const int H = cfg.hidden_size, I = cfg.intermediate_size, nh = cfg.num_heads, nkv = cfg.num_kv_heads;
    const int QKV = qkv_dim();
    const int wblocks = (int)ceil_div(T * 32, 256);
    const int rblocks = std::min(wblocks, 2 * props.sm_count);  // rmsnorm: grid-stride, two CTAs per SM
    last_B = B;
    last_S = S;
    last_flops = 0;
    last_launches = 0;
    {
      Span sp(this, st, 2);
      launch_pdl(embed_kernel, dim3(wblocks), dim3(256), 0, st, T, H, cfg.vocab_size, ids, embed.p, h.p);
    }
    ++last_launches;
    const float scale_log2 = (1.0f / sqrtf((float)kHD)) * 1.4426950408889634f;
    for (int l = 0; l < cfg.num_layers; ++l) {
      Layer& L = layers[l];
      {
        Span sp(this, st, 2);
        launch_pdl(rmsnorm_kernel<bf16>, dim3(rblocks), dim3(256), 0, st, T, H, cfg.rms_eps, h.p, L.ln1.p, xn.p, nullptr);
      }
      {
        // QKV projection with bias and RoPE (q and k heads) fused into the epilogue
        GemmRope rope;
        rope.cs = rope_cs.p;
        rope.S = S;
        rope.cols = (nh + nkv) * kHD;
        rope.ld = cfg.max_seq_len;
        gemm(EPI_BF16_BIAS_ROPE, (int)T, QKV, H, xn.p, L.wqkv.p, qkv.p, QKV, L.bqkv.p, st, &rope);
      }
      {
        Span sp(this, st, 1);
        if (attention_impl != 0 && attention_tc_supported(S)) {
          // tcgen05: S = QK^T and O = PV on the 5th-gen tensor cores, P kept in TMEM
          attention_tc(qkv.p, QKV, mask, ao.p, nh * kHD, B, S, nh, nkv, cfg.causal, scale_log2, props.sm_count,
                       attention_impl == 1, st);
        } else {
          const int units = (nh / nkv) * (int)ceil_div(S, 16);
          dim3 grid((unsigned)ceil_div(units, kAttnWarps), (unsigned)nkv, (unsigned)B);
          launch_pdl(attention_kernel, grid, dim3(kAttnWarps * 32), 0, st, qkv.p, QKV, mask, ao.p, nh * kHD, S, nh, nkv, cfg.causal,
                     scale_log2);
        }
      }
      last_flops += 4.0 * (double)B * S * S * kHD * nh;
      gemm(EPI_F32_ADD, (int)T, H, nh * kHD, ao.p, L.wo.p, h.p, H, nullptr, st);
      {
        Span sp(this, st, 2);
        launch_pdl(rmsnorm_kernel<bf16>, dim3(rblocks), dim3(256), 0, st, T, H, cfg.rms_eps, h.p, L.ln2.p, xn.p, nullptr);
      }
      gemm(EPI_SWIGLU_BF16, (int)T, 2 * I, H, xn.p, L.wgu.p, act.p, I, nullptr, st);
      gemm(EPI_F32_ADD, (int)T, H, I, act.p, L.wd.p, h.p, H, nullptr, st);
      last_launches += 3;
    }
    {
      Span sp(this, st, 2);
      launch_pdl(rmsnorm_kernel<bf16>, dim3(rblocks), dim3(256), 0, st, T, H, cfg.rms_eps, h.p, final_norm.p, nullptr, rinv.p);
      dim3 pg((unsigned)ceil_div(H, 128), (unsigned)B);
      launch_pdl(pool_kernel, pg, dim3(128), 0, st, S, H, h.p, rinv.p, mask, final_norm.p, pooled.p);
    }
    gemm(EPI_F32_BIAS, B, cfg.embed_dim, H, pooled.p, dense_w.p, out, cfg.embed_dim, dense_b.p, st);
    if (normalize) {
      Span sp(this, st, 2);
      launch_pdl(l2norm_kernel, dim3((unsigned)ceil_div((int64_t)B * 32, 256)), dim3(256), 0, st, (int64_t)B, cfg.embed_dim, out);
      ++last_launches;
    }
    last_launches += 2;
    if (profile) ++prof_forwards;
  }

  // fold the events of the most recent forward into prof_ms (call after the stream is idle)
  void fold_profile() {
    for (size_t i = 0; i < ev_used; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev_pool[i].first, ev_pool[i].second) == cudaSuccess) prof_ms[ev_kind[i]] += ms;
    }
    ev_used = 0;
    ev_kind.clear();
  }
};

}  // namespace absb

using namespace absb;

struct absb_enc_s {
  Encoder enc;
  absb_enc_s(const absb_enc_config& c, int dev) : enc(c, dev) {}
};

#define NEED(p) ABSB_CHECK((p) != nullptr, ABSB_ERR_INVALID, "null argument: " #p)

extern "C" {

int absb_enc_create(const absb_enc_config* cfg, int device, absb_enc_t* out) {
  ABSB_API_BEGIN
  NEED(cfg); NEED(out);
  const DeviceProps p = device_props(device);
  ABSB_CHECK(p.cc_major == 10, ABSB_ERR_UNSUPPORTED,
             "device %d is sm_%d%d; libabsb200 is built for sm_100a (B200) only and has no fallback", device,
             p.cc_major, p.cc_minor);
  *out = new absb_enc_s(*cfg, device);
  ABSB_API_END
}

int absb_enc_destroy(absb_enc_t e) {
  ABSB_API_BEGIN
  delete e;
  ABSB_API_END
}

int absb_enc_load_weight(absb_enc_t e, const char* name, const void* data, int dtype, const int64_t* shape, int ndim) {
  ABSB_API_BEGIN
  NEED(e); NEED(name); NEED(data); NEED(shape);
  DeviceGuard g(e->enc.device);
  e->enc.load_weight(name, data, dtype, shape, ndim);
  ABSB_API_END
}

int absb_enc_init_random(absb_enc_t e, uint64_t seed, float std) {
  ABSB_API_BEGIN
  NEED(e);
  DeviceGuard g(e->enc.device);
  e->enc.init_random(seed, std);
  ABSB_API_END
}

int absb_enc_get_weight(absb_enc_t e, const char* name, float* out, int64_t numel) {
  ABSB_API_BEGIN
  NEED(e); NEED(name); NEED(out);
  DeviceGuard g(e->enc.device);
  e->enc.get_weight(name, out, numel);
  ABSB_API_END
}

int absb_enc_forward_dev(absb_enc_t e, int B, int S, const int64_t* input_ids_dev, const int32_t* attention_mask_dev,
                         int normalize, float* out_dev, void* stream) {
  ABSB_API_BEGIN
  NEED(e); NEED(input_ids_dev); NEED(attention_mask_dev); NEED(out_dev);
  DeviceGuard g(e->enc.device);
  e->enc.forward_dev(B, S, reinterpret_cast<const long long*>(input_ids_dev), attention_mask_dev, normalize, out_dev,
                     (cudaStream_t)stream);
  ABSB_API_END
}

int absb_enc_forward(absb_enc_t e, int B, int S, const int64_t* input_ids, const int32_t* attention_mask, int normalize,
                     float* out) {
  ABSB_API_BEGIN
  NEED(e); NEED(input_ids); NEED(attention_mask); NEED(out);
  Encoder& enc = e->enc;
  ABSB_CHECK(B >= 1 && S >= 1, ABSB_ERR_INVALID, "B=%d S=%d", B, S);
  DeviceGuard g(enc.device);
  cudaStream_t st = enc.own_stream;
  const size_t T = (size_t)B * S;
  enc.ids_ws.reserve(T);
  enc.mask_ws.reserve(T);
  enc.ensure_capacity(B, S);
  ABSB_CUDA(cudaMemcpyAsync(enc.ids_ws.p, input_ids, T * sizeof(long long), cudaMemcpyHostToDevice, st));
  ABSB_CUDA(cudaMemcpyAsync(enc.mask_ws.p, attention_mask, T * sizeof(int), cudaMemcpyHostToDevice, st));
  enc.forward_dev(B, S, enc.ids_ws.p, enc.mask_ws.p, normalize, enc.emb.p, st);
  ABSB_CUDA(cudaMemcpyAsync(out, enc.emb.p, (size_t)B * enc.cfg.embed_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
  ABSB_CUDA(cudaStreamSynchronize(st));
  ABSB_API_END
}

int absb_enc_last_hidden(absb_enc_t e, float* out, int64_t numel) {
  ABSB_API_BEGIN
  NEED(e); NEED(out);
  Encoder& enc = e->enc;
  const int64_t T = (int64_t)enc.last_B * enc.last_S;
  ABSB_CHECK(T > 0, ABSB_ERR_STATE, "no forward has run yet");
  ABSB_CHECK(numel == T * enc.cfg.hidden_size, ABSB_ERR_INVALID, "buffer has %lld elements, last_hidden_state has %lld",
             (long long)numel, (long long)(T * enc.cfg.hidden_size));
  DeviceGuard g(enc.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  DBuf<float> y;
  y.alloc_exact((size_t)numel);
  rmsnorm_kernel<float><<<(unsigned)ceil_div(T * 32, 256), 256, 0, enc.own_stream>>>(T, enc.cfg.hidden_size, enc.cfg.rms_eps,
                                                                                   enc.h.p, enc.final_norm.p, y.p, nullptr);
  ABSB_CUDA(cudaGetLastError());
  ABSB_CUDA(cudaMemcpyAsync(out, y.p, (size_t)numel * 4, cudaMemcpyDeviceToHost, enc.own_stream));
  ABSB_CUDA(cudaStreamSynchronize(enc.own_stream));
  ABSB_API_END
}

int absb_enc_last_stats(absb_enc_t e, double* flops, int64_t* launches) {
  ABSB_API_BEGIN
  NEED(e);
  if (flops) *flops = e->enc.last_flops;
  if (launches) *launches = e->enc.last_launches;
  ABSB_API_END
}

int absb_enc_set_attention_impl(absb_enc_t e, int impl) {
  ABSB_API_BEGIN
  NEED(e);
  ABSB_CHECK(impl >= 0 && impl <= 2, ABSB_ERR_INVALID, "attention impl %d", impl);
  e->enc.attention_impl = impl;
  ABSB_API_END
}

int absb_enc_set_profile(absb_enc_t e, int on) {
  ABSB_API_BEGIN
  NEED(e);
  Encoder& enc = e->enc;
  DeviceGuard g(enc.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  if (enc.profile) enc.fold_profile();
  enc.profile = on != 0;
  if (on == 2) {  // reset
    enc.prof_ms[0] = enc.prof_ms[1] = enc.prof_ms[2] = 0;
    enc.prof_gemm_flops = 0;
    enc.prof_forwards = 0;
  }
  ABSB_API_END
}

int absb_enc_profile_spans(absb_enc_t e, void* base_event, float* out, int64_t cap, int64_t* n) {
  ABSB_API_BEGIN
  NEED(e); NEED(base_event); NEED(out); NEED(n);
  Encoder& enc = e->enc;
  DeviceGuard g(enc.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  int64_t m = 0;
  for (size_t i = 0; i < enc.ev_used && m < cap; ++i) {
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, (cudaEvent_t)base_event, enc.ev_pool[i].first) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&t1, (cudaEvent_t)base_event, enc.ev_pool[i].second) != cudaSuccess) continue;
    out[3 * m] = (float)enc.ev_kind[i];
    out[3 * m + 1] = t0;
    out[3 * m + 2] = t1;
    ++m;
  }
  *n = m;
  ABSB_API_END
}

int absb_enc_get_profile(absb_enc_t e, double* gemm_ms, double* gemm_flops, double* attention_ms, double* other_ms,
                         int64_t* forwards) {
  ABSB_API_BEGIN
  NEED(e);
  Encoder& enc = e->enc;
  DeviceGuard g(enc.device);
  ABSB_CUDA(cudaDeviceSynchronize());
  enc.fold_profile();
  if (gemm_ms) *gemm_ms = enc.prof_ms[0];
  if (gemm_flops) *gemm_flops = enc.prof_gemm_flops;
  if (attention_ms) *attention_ms = enc.prof_ms[1];
  if (other_ms) *other_ms = enc.prof_ms[2];
  if (forwards) *forwards = enc.prof_forwards;
  ABSB_API_END
}

}  // extern "C"
