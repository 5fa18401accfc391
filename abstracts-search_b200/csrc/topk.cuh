// topk.cuh — warp-resident k-best container used by the list scan, the coarse select and the merges.
//
// Replaces faiss's per-query binary heap (heap_replace_top / heap_reorder, reached from
// IndexIVF::search — /root/reference/Makefile:31-32, README.md:16) with a structure that fits the
// 32-wide warp: the current best 32*SLOTS candidates stay SORTED across the warp's registers
// (rank r lives in lane r / SLOTS, slot r % SLOTS), the k-th best is cached warp-uniformly as the
// admission threshold, and an insertion is one ballot + one shuffle-up + a short predicated shift.
// More than 99% of scanned vectors fail the threshold test and cost one compare.
//
// Total order (identical to oracle/ivf_oracle.c `worse`): score descending, then id ascending.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace absb {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr long long kIdSentinel = 0x7fffffffffffffffLL;

__device__ __forceinline__ bool ranks_before(float sa, long long ia, float sb, long long ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

template <int SLOTS>
struct WarpTopK {
  float s[SLOTS];
  long long id[SLOTS];
  float thr_s;       // k-th best, warp-uniform
  long long thr_id;  // "
  int k;
  int lane;

  __device__ __forceinline__ void init(int k_, int lane_) {
    k = k_;
    lane = lane_;
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      s[i] = -INFINITY;
      id[i] = kIdSentinel;
    }
    thr_s = -INFINITY;
    thr_id = kIdSentinel;
  }

  // Cheap warp-uniform pre-test: can a candidate with this score possibly enter?
  __device__ __forceinline__ bool may_enter(float cs) const { return cs >= thr_s; }

  __device__ __forceinline__ bool admits(float cs, long long cid) const {
    return ranks_before(cs, cid, thr_s, thr_id);
  }

  // Warp-uniform candidate.  All 32 lanes must call.
  __device__ __forceinline__ void insert(float cs, long long cid) {
    const bool before_last = ranks_before(cs, cid, s[SLOTS - 1], id[SLOTS - 1]);
    const unsigned m = __ballot_sync(kFullMask, before_last);
    if (m == 0) return;
    const int L = __ffs(m) - 1;  // first lane whose block the candidate enters
    float in_s = __shfl_up_sync(kFullMask, s[SLOTS - 1], 1);
    long long in_id = __shfl_up_sync(kFullMask, id[SLOTS - 1], 1);
    if (lane >= L) {
      int p = 0;
      if (lane == L) {
        in_s = cs;
        in_id = cid;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) p += ranks_before(s[i], id[i], cs, cid) ? 1 : 0;
      }
#pragma unroll
      for (int i = SLOTS - 1; i >= 1; --i) {
        if (i > p) {
          s[i] = s[i - 1];
          id[i] = id[i - 1];
        } else if (i == p) {
          s[i] = in_s;
          id[i] = in_id;
        }
      }
      if (p == 0) {
        s[0] = in_s;
        id[0] = in_id;
      }
    }
    refresh_threshold();
  }

  __device__ __forceinline__ void refresh_threshold() {
    const int r = k - 1;
    float ts = s[0];
    long long ti = id[0];
    if (SLOTS > 1) {
      const int slot = r % SLOTS;
#pragma unroll
      for (int i = 1; i < SLOTS; ++i)
        if (i == slot) {
          ts = s[i];
          ti = id[i];
        }
    }
    thr_s = __shfl_sync(kFullMask, ts, r / SLOTS);
    thr_id = __shfl_sync(kFullMask, ti, r / SLOTS);
  }

  // Offer one candidate per lane (lane-private values); `valid` masks lanes without one.
  __device__ __forceinline__ void offer_lanes(float cs, long long cid, bool valid) {
    unsigned m = __ballot_sync(kFullMask, valid && admits(cs, cid));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float bs = __shfl_sync(kFullMask, cs, src);
      const long long bi = __shfl_sync(kFullMask, cid, src);
      if (admits(bs, bi)) insert(bs, bi);
    }
  }

  // Write ranks [0,k) to out_s/out_id (rank-major).  Sentinels are written as they are;
  // finalisation to faiss's (-FLT_MAX, -1) padding happens in the last merge.
  __device__ __forceinline__ void store(float* out_s, long long* out_id) const {
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const int r = lane * SLOTS + i;
      if (r < k) {
        out_s[r] = s[i];
        out_id[r] = id[i];
      }
    }
  }
};

// Number of register slots per lane needed for k results.
inline int slots_for_k(int k) {
  int s = 1;
  while (s * 32 < k) s *= 2;
  return s;
}

// Dispatch helper: calls f(std::integral_constant<int,SLOTS>) for the smallest SLOTS >= k/32.
#define ABSB_DISPATCH_SLOTS(k, ...)                                                    \
  do {                                                                                 \
    const int _slots = ::absb::slots_for_k(k);                                         \
    if (_slots == 1) { constexpr int SLOTS = 1; __VA_ARGS__; }                         \
    else if (_slots == 2) { constexpr int SLOTS = 2; __VA_ARGS__; }                    \
    else if (_slots == 4) { constexpr int SLOTS = 4; __VA_ARGS__; }                    \
    else if (_slots == 8) { constexpr int SLOTS = 8; __VA_ARGS__; }                    \
    else ::absb::fail(ABSB_ERR_INVALID, "k=%d exceeds ABSB_MAX_K=%d", (int)(k), ABSB_MAX_K); \
  } while (0)

}  // namespace absb
