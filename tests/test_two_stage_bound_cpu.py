"""The inequality behind the two-stage scan's per-query proof (csrc/ivf_scan16.cu), checked in numpy:

    |fl32(q . fp16(x)) - fl32(q . x)|  <=  B(q) = |q| * (r_max + 80 u (x_max + r_max)),  u = 2^-24,

with r_max = max |x - fp16(x)|_2 and x_max = max |x|_2 over the corpus.  The kernels accumulate 32
chained FMAs per lane and 5 butterfly levels (<= 37 roundings per pass); here both passes are
emulated in that exact lane order in float32 (products rounded separately, i.e. with MORE roundings
than the FMAs the GPU uses) and compared against the bound, on distributions that stress it: unit
gaussian rows, heavy tails, values in fp16's subnormal range, values near fp16's overflow."""
import numpy as np
import pytest

U = np.float32(2.0 ** -24)


def lane_order_dot(q, x, per_lane):
    """fp32 dot in the kernels' order: lane l owns `per_lane`-element groups (g * 32 + l), chained
    left to right, then a 5-level butterfly over the 32 lanes."""
    d = q.shape[0]
    groups = d // per_lane
    prod = (q.astype(np.float32) * x.astype(np.float32)).astype(np.float32).reshape(groups, per_lane)
    acc = np.zeros(32, dtype=np.float32)
    for g0 in range(0, groups, 32):  # lane l takes group g0 + l, elements in order
        block = prod[g0:g0 + 32]
        for e in range(per_lane):
            acc = (acc + block[:, e]).astype(np.float32)
    w = 16
    while w:
        acc = (acc + acc[np.arange(32) ^ w]).astype(np.float32)
        w >>= 1
    return acc[0]


def bound(q, r_max, x_max):
    qn = np.float32(np.linalg.norm(q.astype(np.float64))) * np.float32(1.0001)
    r = np.float32(r_max) * np.float32(1.0001)
    xm = np.float32(x_max) * np.float32(1.0001)
    return float(qn) * (float(r) + 80.0 * float(U) * (float(xm) + float(r)))


@pytest.mark.parametrize("kind", ["unit_gauss", "heavy_tail", "subnormal", "near_overflow", "lattice"])
def test_fp16_pass_score_is_within_the_bound_of_the_fp32_score(kind):
    rng = np.random.default_rng(sum(ord(ch) for ch in kind))
    d, n, nq = 1024, 64, 8
    if kind == "unit_gauss":
        x = rng.standard_normal((n, d))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "heavy_tail":
        x = rng.standard_t(2, (n, d)) * 0.1
    elif kind == "subnormal":
        x = rng.standard_normal((n, d)) * 1e-6
    elif kind == "near_overflow":
        x = rng.standard_normal((n, d)) * 1.5e4
    else:
        x = rng.integers(-127, 128, (n, d)) / 128.0
    x = x.astype(np.float32)
    q = (x[rng.integers(0, n, nq)] * (1 + 0.1 * rng.standard_normal((nq, d)))).astype(np.float32)
    with np.errstate(over="ignore"):
        x16 = x.astype(np.float16).astype(np.float32)
    r_max = float(np.max(np.linalg.norm(x.astype(np.float64) - x16.astype(np.float64), axis=1)))
    x_max = float(np.max(np.linalg.norm(x.astype(np.float64), axis=1)))
    if kind == "lattice":
        assert r_max == 0.0, "multiples of 1/128 below 1 are exact in fp16"
    worst = 0.0
    for i in range(nq):
        B = bound(q[i], r_max, x_max)
        for j in range(0, n, 7):
            a = lane_order_dot(q[i], x16[j], 8)   # stage 1: 8 halves per 128-bit load
            e = lane_order_dot(q[i], x[j], 4)     # stage 2 / single pass: float4 loads
            if not np.isfinite(B):
                continue
            assert abs(float(a) - float(e)) <= B, (kind, i, j, float(a), float(e), B)
            worst = max(worst, abs(float(a) - float(e)) / B if B > 0 else 0.0)
    assert worst <= 1.0
