"""Counter-based synthetic corpora (TEST INFRASTRUCTURE — part of the oracle, never the product).

The reference ships no data and no golden vectors, so every corpus used for
parity and for bench.py is generated from a counter-based integer hash that is implemented twice,
bit for bit: here in numpy and in `abstracts-search_b200/csrc/synth.cu` on the device
(`absb_synth_fill_dev`).  SURVEY.md §8(d) "Synthetic inputs" is the specification.

All values live on an integer lattice: component = int in [-127, 127] divided by 128.  Every
product of two components is an integer / 2^14 with |integer| <= 127^2 and every partial sum of
1024 such products stays below 2^24 in magnitude, so an fp32 inner product is EXACT in any
summation order, with or without FMA.  That is what lets the tests demand bit-identical scores
as well as bit-identical ids from the CUDA path.

    mu[j]      (j < nlist)      cluster centre, components in [-96, 96]
    eps[r]                      per-row noise,   components in [-31, 31]
    c(r)                        cluster of corpus row r (uniform over lists)
    corpus  x[r] = (mu[c(r)] + eps[r]) / 128
    centroid[j]  =  mu[j] / 128
    query   q[i] = clamp(mu[c(s)] + eps[s] + delta[i], -127, 127) / 128,  s = src(i), delta in [-15, 15]
"""
from __future__ import annotations

import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
C1 = np.uint64(0xBF58476D1CE4E5B9)
C2 = np.uint64(0x94D049BB133111EB)
C3 = np.uint64(0xD1B54A32D192ED03)

SALT_MU = np.uint64(0x6D75)
SALT_EPS = np.uint64(0x657073)
SALT_CL = np.uint64(0x636C)
SALT_QSRC = np.uint64(0x71737263)
SALT_QDELTA = np.uint64(0x7164)

KIND_CORPUS, KIND_CENTROIDS, KIND_QUERIES = 0, 1, 2
KIND_UNIT = 4  # flag: the same integer row divided by its L2 norm (real-valued fp32, unit length)


def _mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * C1
        z = (z ^ (z >> np.uint64(27))) * C2
        return z ^ (z >> np.uint64(31))


def _row_key(seed: np.uint64, rows: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) + GOLD * (rows.astype(np.uint64) + np.uint64(1)))


def _bytes(seed: np.uint64, rows: np.ndarray, d: int) -> np.ndarray:
    """uint8 [len(rows), d]: byte k of word g is column 8g+k."""
    assert d % 8 == 0
    key = _row_key(seed, rows)[:, None]
    g = np.arange(d // 8, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        w = _mix64(key ^ (C3 * (g + np.uint64(1))))
    shifts = (np.arange(8, dtype=np.uint64) * np.uint64(8))[None, None, :]
    b = (w[:, :, None] >> shifts) & np.uint64(0xFF)
    return b.reshape(len(rows), d).astype(np.int32)


def cluster_of(seed: int, rows: np.ndarray, nlist: int) -> np.ndarray:
    """c(r): int32 cluster / generating-list id of corpus rows."""
    k = _row_key(np.uint64(seed) ^ SALT_CL, np.asarray(rows))
    return ((k >> np.uint64(33)) % np.uint64(nlist)).astype(np.int32)


def mu_int(seed: int, lists: np.ndarray, d: int) -> np.ndarray:
    return _bytes(np.uint64(seed) ^ SALT_MU, np.asarray(lists), d) % 193 - 96


def eps_int(seed: int, rows: np.ndarray, d: int) -> np.ndarray:
    return _bytes(np.uint64(seed) ^ SALT_EPS, np.asarray(rows), d) % 63 - 31


def corpus_int(seed: int, rows: np.ndarray, d: int, nlist: int) -> np.ndarray:
    rows = np.asarray(rows)
    return mu_int(seed, cluster_of(seed, rows, nlist), d) + eps_int(seed, rows, d)


def corpus(seed: int, row0: int, n: int, d: int, nlist: int) -> np.ndarray:
    """float32 [n, d] corpus rows row0..row0+n."""
    rows = np.arange(row0, row0 + n, dtype=np.int64)
    return (corpus_int(seed, rows, d, nlist).astype(np.float32) / np.float32(128.0))


def corpus_rows(seed: int, rows: np.ndarray, d: int, nlist: int) -> np.ndarray:
    return (corpus_int(seed, rows, d, nlist).astype(np.float32) / np.float32(128.0))


def centroids(seed: int, nlist: int, d: int, list0: int = 0, n: int | None = None) -> np.ndarray:
    n = nlist - list0 if n is None else n
    lists = np.arange(list0, list0 + n, dtype=np.int64)
    return mu_int(seed, lists, d).astype(np.float32) / np.float32(128.0)


def unit_rows(v: np.ndarray) -> np.ndarray:
    """int32 lattice rows -> float32 unit-norm rows, bit-identical to synth.cu's unit variant: the squared
    norm (< 2^24 for d <= 1024) is an exact fp32 integer, sqrt and divide are correctly rounded IEEE
    operations on both sides.  Real-valued embeddings that fp16 cannot hold exactly — what
    `index fill` stores (/root/reference/Makefile:24-25) — yet regenerable by the oracle bit for bit."""
    v = np.asarray(v, dtype=np.int32)
    assert v.shape[1] <= 1024
    ss = (v.astype(np.int64) * v).sum(axis=1)
    den = np.sqrt(ss.astype(np.float32), dtype=np.float32)
    out = np.zeros(v.shape, dtype=np.float32)
    np.divide(v.astype(np.float32), den[:, None], out=out, where=(ss > 0)[:, None])
    return out


def corpus_rows_unit(seed: int, rows: np.ndarray, d: int, nlist: int) -> np.ndarray:
    return unit_rows(corpus_int(seed, rows, d, nlist))


def corpus_unit(seed: int, row0: int, n: int, d: int, nlist: int) -> np.ndarray:
    return corpus_rows_unit(seed, np.arange(row0, row0 + n, dtype=np.int64), d, nlist)


def queries_unit(seed: int, q0: int, n: int, d: int, nlist: int, corpus_rows_total: int) -> np.ndarray:
    idx = np.arange(q0, q0 + n, dtype=np.int64)
    src = query_src(seed, idx, corpus_rows_total)
    delta = _bytes(np.uint64(seed) ^ SALT_QDELTA, idx, d) % 31 - 15
    return unit_rows(np.clip(corpus_int(seed, src, d, nlist) + delta, -127, 127))


def query_src(seed: int, idx: np.ndarray, corpus_rows_total: int) -> np.ndarray:
    k = _row_key(np.uint64(seed) ^ SALT_QSRC, np.asarray(idx))
    return ((k >> np.uint64(1)) % np.uint64(corpus_rows_total)).astype(np.int64)


def queries(seed: int, q0: int, n: int, d: int, nlist: int, corpus_rows_total: int) -> np.ndarray:
    """float32 [n, d]: perturbed corpus rows."""
    idx = np.arange(q0, q0 + n, dtype=np.int64)
    src = query_src(seed, idx, corpus_rows_total)
    delta = _bytes(np.uint64(seed) ^ SALT_QDELTA, idx, d) % 31 - 15
    q = np.clip(corpus_int(seed, src, d, nlist) + delta, -127, 127)
    return q.astype(np.float32) / np.float32(128.0)


def rows_of_lists(seed: int, corpus_rows_total: int, nlist: int, lists, block: int = 1 << 22):
    """Inverse of cluster_of for a few lists: dict list -> ascending int64 rows.  Lets the oracle
    rebuild only the probed lists of a huge index (SURVEY §7.2 'Memory at 207M')."""
    want = np.unique(np.asarray(lists, dtype=np.int64))
    out = {int(l): [] for l in want}
    for r0 in range(0, corpus_rows_total, block):
        rows = np.arange(r0, min(r0 + block, corpus_rows_total), dtype=np.int64)
        c = cluster_of(seed, rows, nlist)
        m = np.isin(c, want)
        for l, r in zip(c[m], rows[m]):
            out[int(l)].append(int(r))
    return {l: np.asarray(v, dtype=np.int64) for l, v in out.items()}
