// synth.cu — device side of the counter-based synthetic corpus generator.
//
// Bit-for-bit twin of oracle/synth.py (SURVEY.md §8d "Synthetic inputs"): the reference ships no
// data, so bench.py and the full-size parity tests fill the index from this generator while the
// oracle regenerates exactly the same rows on the CPU.  One thread produces the 8 columns that
// come out of one 64-bit hash word and stores them as two float4.
#include "common.cuh"

namespace absb {

namespace {

constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ull;
constexpr uint64_t C1 = 0xBF58476D1CE4E5B9ull;
constexpr uint64_t C2 = 0x94D049BB133111EBull;
constexpr uint64_t C3 = 0xD1B54A32D192ED03ull;
constexpr uint64_t SALT_MU = 0x6D75ull;
constexpr uint64_t SALT_EPS = 0x657073ull;
constexpr uint64_t SALT_CL = 0x636Cull;
constexpr uint64_t SALT_QSRC = 0x71737263ull;
constexpr uint64_t SALT_QDELTA = 0x7164ull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * C1;
  z = (z ^ (z >> 27)) * C2;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t row_key(uint64_t seed, uint64_t row) {
  return mix64(seed + GOLD * (row + 1));
}
__host__ __device__ __forceinline__ uint64_t word(uint64_t key, uint64_t g) {
  return mix64(key ^ (C3 * (g + 1)));
}
__host__ __device__ __forceinline__ int cluster_of(uint64_t seed, uint64_t row, int nlist) {
  return (int)((row_key(seed ^ SALT_CL, row) >> 33) % (uint64_t)nlist);
}

__global__ void synth_fill_kernel(int kind, uint64_t seed, int64_t row0,
                                  const long long* __restrict__ row_ids, int64_t n, int d,
                                  int nlist, int64_t corpus_rows, float* __restrict__ out) {
  const int groups = d / 8;
  const int64_t total = n * groups;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / groups;
    const int g = (int)(t % groups);
    const uint64_t r = row_ids ? (uint64_t)row_ids[i] : (uint64_t)(row0 + i);
    int v[8];
    if (kind == 1) {  // centroid r
      const uint64_t w = word(row_key(seed ^ SALT_MU, r), g);
#pragma unroll
      for (int b = 0; b < 8; ++b) v[b] = (int)((w >> (8 * b)) & 0xFF) % 193 - 96;
    } else {
      uint64_t src = r;
      if (kind == 2) src = (row_key(seed ^ SALT_QSRC, r) >> 1) % (uint64_t)corpus_rows;
      const int c = cluster_of(seed, src, nlist);
      const uint64_t wm = word(row_key(seed ^ SALT_MU, (uint64_t)c), g);
      const uint64_t we = word(row_key(seed ^ SALT_EPS, src), g);
#pragma unroll
      for (int b = 0; b < 8; ++b)
        v[b] = (int)((wm >> (8 * b)) & 0xFF) % 193 - 96 + (int)((we >> (8 * b)) & 0xFF) % 63 - 31;
      if (kind == 2) {
        const uint64_t wd = word(row_key(seed ^ SALT_QDELTA, r), g);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          int x = v[b] + (int)((wd >> (8 * b)) & 0xFF) % 31 - 15;
          v[b] = x < -127 ? -127 : (x > 127 ? 127 : x);
        }
      }
    }
    float4 lo = make_float4(v[0] * 0.0078125f, v[1] * 0.0078125f, v[2] * 0.0078125f, v[3] * 0.0078125f);
    float4 hi = make_float4(v[4] * 0.0078125f, v[5] * 0.0078125f, v[6] * 0.0078125f, v[7] * 0.0078125f);
    float4* o = reinterpret_cast<float4*>(out + i * d + g * 8);
    o[0] = lo;
    o[1] = hi;
  }
}

// Unit-norm variant (kind | 4): the same integer row v, written as v / sqrt(sum v^2) — a real-valued fp32
// corpus (what `index fill` stores: L2-normalised embeddings, /root/reference/Makefile:24-25) that is
// NOT exactly representable in fp16.  sum v^2 <= 1024 * 127^2 < 2^24 is an exact integer in fp32, and
// sqrt / divide are the correctly rounded IEEE operations, so numpy reproduces every bit
// (oracle/synth.py: v.astype(f32) / np.sqrt(f32(ss))).  One warp per row.
__global__ void synth_fill_unit_kernel(int kind, uint64_t seed, int64_t row0, const long long* __restrict__ row_ids,
                                       int64_t n, int d, int nlist, int64_t corpus_rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int groups = d / 8;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < n;
       i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const uint64_t r = row_ids ? (uint64_t)row_ids[i] : (uint64_t)(row0 + i);
    uint64_t src = r;
    if (kind == 2) src = (row_key(seed ^ SALT_QSRC, r) >> 1) % (uint64_t)corpus_rows;
    const int c = kind == 1 ? 0 : cluster_of(seed, src, nlist);
    const uint64_t key_mu = row_key(seed ^ SALT_MU, kind == 1 ? r : (uint64_t)c);
    const uint64_t key_eps = row_key(seed ^ SALT_EPS, src);
    const uint64_t key_d = row_key(seed ^ SALT_QDELTA, r);
    int ss = 0;
    for (int pass = 0; pass < 2; ++pass) {
      float inv_den = 0.f;
      if (pass == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        inv_den = __fsqrt_rn((float)ss);
      }
      for (int g = lane; g < groups; g += 32) {
        int v[8];
        const uint64_t wm = word(key_mu, g);
        if (kind == 1) {
#pragma unroll
          for (int b = 0; b < 8; ++b) v[b] = (int)((wm >> (8 * b)) & 0xFF) % 193 - 96;
        } else {
          const uint64_t we = word(key_eps, g);
#pragma unroll
          for (int b = 0; b < 8; ++b)
            v[b] = (int)((wm >> (8 * b)) & 0xFF) % 193 - 96 + (int)((we >> (8 * b)) & 0xFF) % 63 - 31;
          if (kind == 2) {
            const uint64_t wd = word(key_d, g);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              int x = v[b] + (int)((wd >> (8 * b)) & 0xFF) % 31 - 15;
              v[b] = x < -127 ? -127 : (x > 127 ? 127 : x);
            }
          }
        }
        if (pass == 0) {
#pragma unroll
          for (int b = 0; b < 8; ++b) ss += v[b] * v[b];
        } else {
          float4* o = reinterpret_cast<float4*>(out + i * d + g * 8);
          // ss == 0 cannot happen for d >= 8 in practice; keep the row zero instead of NaN if it does
          const bool ok = ss > 0;
          o[0] = make_float4(ok ? __fdiv_rn((float)v[0], inv_den) : 0.f, ok ? __fdiv_rn((float)v[1], inv_den) : 0.f,
                             ok ? __fdiv_rn((float)v[2], inv_den) : 0.f, ok ? __fdiv_rn((float)v[3], inv_den) : 0.f);
          o[1] = make_float4(ok ? __fdiv_rn((float)v[4], inv_den) : 0.f, ok ? __fdiv_rn((float)v[5], inv_den) : 0.f,
                             ok ? __fdiv_rn((float)v[6], inv_den) : 0.f, ok ? __fdiv_rn((float)v[7], inv_den) : 0.f);
        }
      }
    }
  }
}

__global__ void synth_cluster_kernel(uint64_t seed, int64_t row0, int64_t n, int nlist,
                                     long long* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = cluster_of(seed, (uint64_t)(row0 + i), nlist);
}

}  // namespace

void synth_fill(int kind, uint64_t seed, int64_t row0, const long long* row_ids, int64_t n, int d,
                int nlist, int64_t corpus_rows, float* out, cudaStream_t st) {
  ABSB_CHECK(kind >= 0 && kind <= 6 && (kind & 3) <= 2, ABSB_ERR_INVALID, "synth kind %d", kind);
  ABSB_CHECK(d > 0 && d % 8 == 0, ABSB_ERR_INVALID, "synth needs d %% 8 == 0 (d=%d)", d);
  ABSB_CHECK(nlist > 0 && n >= 0, ABSB_ERR_INVALID, "synth nlist/n");
  ABSB_CHECK((kind & 3) != 2 || corpus_rows > 0, ABSB_ERR_INVALID, "queries need corpus_rows");
  if (n == 0) return;
  if (kind & 4) {
    ABSB_CHECK(d <= 1024, ABSB_ERR_UNSUPPORTED, "unit-norm synth rows need d <= 1024 (exact integer norm in fp32)");
    const int blocks = (int)std::min<int64_t>(ceil_div(n * 32, 256), 148 * 16);
    synth_fill_unit_kernel<<<blocks, 256, 0, st>>>(kind & 3, seed, row0, row_ids, n, d, nlist, corpus_rows, out);
    ABSB_CUDA(cudaGetLastError());
    return;
  }
  const int64_t total = n * (d / 8);
  const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  synth_fill_kernel<<<blocks, 256, 0, st>>>(kind, seed, row0, row_ids, n, d, nlist, corpus_rows, out);
  ABSB_CUDA(cudaGetLastError());
}

void synth_cluster(uint64_t seed, int64_t row0, int64_t n, int nlist, long long* out, cudaStream_t st) {
  if (n == 0) return;
  const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 16);
  synth_cluster_kernel<<<blocks, 256, 0, st>>>(seed, row0, n, nlist, out);
  ABSB_CUDA(cudaGetLastError());
}

}  // namespace absb
