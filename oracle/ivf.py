"""CPU oracle for the IVF / Flat inner-product search path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED — see the header of ivf_oracle.c: the arithmetic lives in faiss, which is neither
vendored in /root/reference nor installable offline, and the reference holds no tests or golden
vectors for it.  This module restates the faiss semantics listed in SURVEY.md
§8(a) a4–a8 at the reference's call sites:

    faiss.index_factory(1024, "IVF65536,Flat", METRIC_INNER_PRODUCT)   Makefile:38-39, README.md:60
    Index.train   -> Clustering::train                                  Makefile:38-39
    Index.add     -> quantizer.assign + append in insertion order       Makefile:24-25
    Index.search  -> quantizer.search(nprobe) + scan_codes + k-best     Makefile:31-32, README.md:16,28

Two implementations check each other: plain numpy (this file) and the C restatement in
ivf_oracle.c (OpenMP, used for the bigger cases and as bench.py's CPU baseline).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle() -> str:
    path = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "ivf_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return path


def clib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.orc_num_threads.restype = ctypes.c_int
        _LIB.orc_split_clusters.restype = ctypes.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ------------------------------------------------------------------ k-best --------------------
def topk_by_score_then_id(scores: np.ndarray, ids: np.ndarray, k: int):
    """Best-k of one candidate set ordered by (score desc, id asc); pads with (-FLT_MAX, -1)
    like faiss does for inner product."""
    order = np.lexsort((ids, -scores.astype(np.float64)))[:k]
    D = np.full(k, -FLT_MAX, dtype=np.float32)
    I = np.full(k, -1, dtype=np.int64)
    D[: len(order)] = scores[order]
    I[: len(order)] = ids[order]
    return D, I


# ------------------------------------------------------------------ random --------------------
def mt19937_raw(seed: int, n: int) -> np.ndarray:
    """Raw std::mt19937(seed) outputs (numpy's legacy RandomState seeds with init_genrand too)."""
    rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)
    return rs.randint(0, 2**32, size=n, dtype=np.uint32)


def rand_perm(n: int, seed: int, use_c: bool = True) -> np.ndarray:
    """faiss rand_perm: Fisher–Yates with i2 = i + mt() % (n - i)."""
    if use_c:
        perm = np.empty(n, dtype=np.int32)
        clib().orc_rand_perm(ctypes.c_int64(n), ctypes.c_int64(seed), _p(perm, ctypes.c_int32))
        return perm.astype(np.int64)
    raw = mt19937_raw(seed, max(n - 1, 0))
    perm = list(range(n))
    for i in range(n - 1):
        i2 = i + int(raw[i]) % (n - i)
        perm[i], perm[i2] = perm[i2], perm[i]
    return np.asarray(perm, dtype=np.int64)


# ------------------------------------------------------------------ IndexFlatIP ---------------
class FlatIP:
    """faiss.IndexFlatIP restated."""

    def __init__(self, d: int):
        self.d = d
        self.xb = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.xb.shape[0]

    def reset(self):
        self.xb = np.zeros((0, self.d), dtype=np.float32)

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x], axis=0)

    def search(self, q, k: int, impl: str = "numpy"):
        q = np.ascontiguousarray(q, dtype=np.float32)
        n = q.shape[0]
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        if impl == "c":
            clib().orc_flat_search(ctypes.c_int64(n), _p(q, ctypes.c_float), ctypes.c_int64(self.ntotal),
                                   _p(self.xb, ctypes.c_float), ctypes.c_int(self.d), ctypes.c_int(k),
                                   _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
            return D, I
        ids = np.arange(self.ntotal, dtype=np.int64)
        for i0 in range(0, n, 256):
            s = q[i0:i0 + 256] @ self.xb.T  # fp32 sgemm, what faiss does above 20 queries
            for j in range(s.shape[0]):
                D[i0 + j], I[i0 + j] = topk_by_score_then_id(s[j], ids, k)
        return D, I

    def scores_f64(self, q):
        return np.asarray(q, dtype=np.float64) @ self.xb.astype(np.float64).T


# ------------------------------------------------------------------ Clustering ----------------
def renorm_l2(c: np.ndarray) -> np.ndarray:
    """faiss fvec_renorm_L2 (Clustering::post_process_centroids when cp.spherical): every row with a
    non-zero norm is scaled by 1 / sqrtf(|row|^2), in fp32; zero rows stay zero."""
    c = np.ascontiguousarray(c, dtype=np.float32)
    nr = np.einsum("ij,ij->i", c, c, dtype=np.float32)
    inv = np.zeros_like(nr)
    np.divide(np.float32(1.0), np.sqrt(nr, dtype=np.float32), out=inv, where=nr > 0)
    return c * inv[:, None]


def kmeans_train(x: np.ndarray, k: int, niter: int = 10, max_points_per_centroid: int = 256,
                 seed: int = 1234, assign_fn=None, return_history: bool = False, spherical: bool = False):
    """faiss Clustering::train with an IndexFlatIP assignment index:
    subsample to k*max_ppc rows with rand_perm(seed); centroids = rows rand_perm(seed+1)[:k];
    niter x {assign = argmax IP, centroid = mean, split empty clusters}.  
    spherical (faiss ClusteringParameters.spherical; index_factory turns it on for inner-product IVF
    indexes — external, unpinned, hence a switch): post_process_centroids = fvec_renorm_L2 after the
    initial draw and after every iteration's mean + split."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    assert n >= k, "faiss: number of training points should be at least as large as number of clusters"
    if n > k * max_points_per_centroid:
        perm = rand_perm(n, seed)
        x = np.ascontiguousarray(x[perm[: k * max_points_per_centroid]])
        n = x.shape[0]
    if n == k:
        cent = x.copy()
        return (cent, []) if return_history else cent
    perm = rand_perm(n, seed + 1)
    cent = np.ascontiguousarray(x[perm[:k]])
    if spherical:
        cent = renorm_l2(cent)
    lib = clib()
    history = []
    for _ in range(niter):
        if assign_fn is not None:
            assign = np.ascontiguousarray(assign_fn(x, cent), dtype=np.int64)
        else:
            assign = assign_argmax_ip(x, cent)
        hassign = np.zeros(k, dtype=np.float32)
        new = np.zeros((k, d), dtype=np.float32)
        lib.orc_compute_centroids(ctypes.c_int64(n), _p(x, ctypes.c_float), ctypes.c_int(d),
                                  ctypes.c_int64(k), _p(assign, ctypes.c_int64),
                                  _p(new, ctypes.c_float), _p(hassign, ctypes.c_float))
        # faiss leaves an empty cluster's centroid at zero before the split copies over it
        nsplit = lib.orc_split_clusters(ctypes.c_int(d), ctypes.c_int64(k), ctypes.c_int64(n),
                                        _p(hassign, ctypes.c_float), _p(new, ctypes.c_float))
        cent = renorm_l2(new) if spherical else new
        history.append((assign, int(nsplit)))
    return (cent, history) if return_history else cent


def assign_argmax_ip(x: np.ndarray, cent: np.ndarray) -> np.ndarray:
    """quantizer.assign: argmax_c <x, c>, smallest list number on exact ties."""
    out = np.empty(x.shape[0], dtype=np.int64)
    for i0 in range(0, x.shape[0], 4096):
        s = x[i0:i0 + 4096] @ cent.T
        out[i0:i0 + 4096] = np.argmax(s, axis=1)  # first maximum = smallest id
    return out


def compute_centroids_numpy(x, assign, k):
    """numpy cross-check of orc_compute_centroids (row-order fp32 sums)."""
    d = x.shape[1]
    cent = np.zeros((k, d), dtype=np.float32)
    cnt = np.zeros(k, dtype=np.float32)
    for i in range(x.shape[0]):
        cent[assign[i]] += x[i]
        cnt[assign[i]] += 1
    nz = cnt > 0
    cent[nz] *= (np.float32(1.0) / cnt[nz])[:, None]
    return cent, cnt


# ------------------------------------------------------------------ IndexIVFFlat --------------
class IVFFlat:
    """faiss.IndexIVFFlat (inner product) restated; lists keep insertion order."""

    def __init__(self, d: int, nlist: int):
        self.d, self.nlist = d, nlist
        self.nprobe = 1
        self.niter, self.max_points_per_centroid, self.seed = 10, 256, 1234
        self.spherical = False  # ClusteringParameters.spherical (faiss index_factory: True for inner product)
        self.centroids = None
        self.ntotal = 0
        self.codes = [np.zeros((0, d), dtype=np.float32) for _ in range(nlist)]
        self.ids = [np.zeros((0,), dtype=np.int64) for _ in range(nlist)]
        self._csr = None

    @property
    def is_trained(self):
        return self.centroids is not None

    def train(self, x):
        self.centroids = kmeans_train(x, self.nlist, self.niter, self.max_points_per_centroid, self.seed,
                                      spherical=self.spherical)

    def set_centroids(self, c):
        c = np.ascontiguousarray(c, dtype=np.float32)
        assert c.shape == (self.nlist, self.d)
        self.centroids = c

    def reset(self):
        self.codes = [np.zeros((0, self.d), dtype=np.float32) for _ in range(self.nlist)]
        self.ids = [np.zeros((0,), dtype=np.int64) for _ in range(self.nlist)]
        self.ntotal, self._csr = 0, None

    def assign(self, x):
        return assign_argmax_ip(np.ascontiguousarray(x, dtype=np.float32), self.centroids)

    def add(self, x, ids=None, list_ids=None):
        assert self.is_trained, "faiss: index not trained"
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        if ids is None:
            ids = np.arange(self.ntotal, self.ntotal + n, dtype=np.int64)
        ids = np.asarray(ids, dtype=np.int64)
        if list_ids is None:
            list_ids = self.assign(x)
        list_ids = np.asarray(list_ids, dtype=np.int64)
        order = np.argsort(list_ids, kind="stable")
        sl = list_ids[order]
        bounds = np.searchsorted(sl, np.arange(self.nlist + 1))
        for l in np.unique(sl[sl >= 0]):
            rows = order[bounds[l]:bounds[l + 1]]
            self.codes[l] = np.concatenate([self.codes[l], x[rows]], axis=0)
            self.ids[l] = np.concatenate([self.ids[l], ids[rows]])
        self.ntotal += int((list_ids >= 0).sum())
        self._csr = None

    add_with_ids = add

    def list_sizes(self):
        return np.asarray([len(i) for i in self.ids], dtype=np.int64)

    def coarse(self, q, nprobe, impl="numpy"):
        q = np.ascontiguousarray(q, dtype=np.float32)
        n = q.shape[0]
        Dc = np.empty((n, nprobe), dtype=np.float32)
        Ic = np.empty((n, nprobe), dtype=np.int64)
        if impl == "c":
            clib().orc_coarse(ctypes.c_int64(n), _p(q, ctypes.c_float), ctypes.c_int64(self.nlist),
                              _p(self.centroids, ctypes.c_float), ctypes.c_int(self.d),
                              ctypes.c_int(nprobe), _p(Dc, ctypes.c_float), _p(Ic, ctypes.c_int64))
            return Dc, Ic
        lists = np.arange(self.nlist, dtype=np.int64)
        for i0 in range(0, n, 256):
            s = q[i0:i0 + 256] @ self.centroids.T
            for j in range(s.shape[0]):
                Dc[i0 + j], Ic[i0 + j] = topk_by_score_then_id(s[j], lists, nprobe)
        return Dc, Ic

    def _as_csr(self):
        if self._csr is None:
            sizes = self.list_sizes()
            off = np.zeros(self.nlist + 1, dtype=np.int64)
            np.cumsum(sizes, out=off[1:])
            codes = np.concatenate(self.codes, axis=0) if self.ntotal else np.zeros((0, self.d), np.float32)
            ids = np.concatenate(self.ids) if self.ntotal else np.zeros((0,), np.int64)
            self._csr = (off, np.ascontiguousarray(codes), np.ascontiguousarray(ids))
        return self._csr

    def search_preassigned(self, q, k, coarse_ids, impl="numpy"):
        q = np.ascontiguousarray(q, dtype=np.float32)
        coarse_ids = np.ascontiguousarray(coarse_ids, dtype=np.int64)
        n, nprobe = coarse_ids.shape
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        if impl == "c":
            off, codes, ids = self._as_csr()
            ns = ctypes.c_int64(0)
            clib().orc_ivf_scan(ctypes.c_int64(n), _p(q, ctypes.c_float), ctypes.c_int(self.d),
                                ctypes.c_int(k), ctypes.c_int(nprobe), _p(coarse_ids, ctypes.c_int64),
                                _p(off, ctypes.c_int64), _p(codes, ctypes.c_float), _p(ids, ctypes.c_int64),
                                _p(D, ctypes.c_float), _p(I, ctypes.c_int64), ctypes.byref(ns))
            self.last_nscanned = ns.value
            return D, I
        for i in range(n):
            ls = [int(l) for l in coarse_ids[i] if l >= 0]
            if ls:
                c = np.concatenate([self.codes[l] for l in ls], axis=0)
                cid = np.concatenate([self.ids[l] for l in ls])
            else:
                c, cid = np.zeros((0, self.d), np.float32), np.zeros((0,), np.int64)
            s = (c @ q[i]).astype(np.float32) if len(cid) else np.zeros((0,), np.float32)
            D[i], I[i] = topk_by_score_then_id(s, cid, k)
        return D, I

    def search(self, q, k, nprobe=None, impl="numpy"):
        nprobe = self.nprobe if nprobe is None else nprobe
        _, Ic = self.coarse(q, min(nprobe, self.nlist), impl="numpy")
        return self.search_preassigned(q, k, Ic, impl=impl)

    # -------- fp64 diagnostics ---------------------------------------------
    def ambiguity(self, q, k, nprobe):
        """Per query: (coarse_margin, fine_margin) in fp64 — the gap between rank nprobe/nprobe+1
        centroids and between rank k/k+1 results among the probed lists.  A query whose margin is
        below the fp32 rounding bound may legitimately differ between two correct fp32
        implementations; tests demand exact ids only where both margins are safe."""
        q64 = np.asarray(q, dtype=np.float64)
        c64 = self.centroids.astype(np.float64)
        n = q64.shape[0]
        cm = np.full(n, np.inf)
        fm = np.full(n, np.inf)
        for i in range(n):
            s = c64 @ q64[i]
            o = np.argsort(-s, kind="stable")
            if nprobe < self.nlist:
                cm[i] = s[o[nprobe - 1]] - s[o[nprobe]]
            ls = o[:nprobe]
            c = np.concatenate([self.codes[l] for l in ls], axis=0).astype(np.float64)
            if c.shape[0] > k:
                f = np.sort(c @ q64[i])[::-1]
                gaps = f[:k] - f[1:k + 1]
                fm[i] = gaps.min()  # any adjacent swap inside the top-k+1 changes ids or order
        return cm, fm
