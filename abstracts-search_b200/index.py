"""faiss-shaped Python surface over libabsb200.so.

Mirrors the slice of the faiss Python API that `sidecar-search index train|fill|tune`
(/root/reference/Makefile:38-39, 24-25, 31-32) and the app.py query loop
(/root/reference/README.md:16,28) use:

    index = index_factory(1024, "IVF65536,Flat", METRIC_INNER_PRODUCT)
    index.train(x); index.add(x) / index.add_with_ids(x, ids)
    index.nprobe = 32;  D, I = index.search(q, 10)

Same names, argument meaning and error behaviour as faiss's SWIG wrappers: inputs are coerced to
C-contiguous float32 [n, d] (shape mismatch -> AssertionError), C++ failures raise RuntimeError,
missing results are I = -1 / D = -3.4028235e38.  Arrays may be numpy (host; copies happen inside the
C call) or CUDA torch tensors (no host hop; results come back as CUDA tensors).
"""
from __future__ import annotations

import re
import struct
import warnings
from ctypes import byref, c_float, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import check, lib, ptr

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def _is_cuda_tensor(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def _as_f32_matrix(x, d: int):
    """faiss's replacement_* wrappers: `n, d = x.shape; assert d == self.d;
    x = np.ascontiguousarray(x, dtype='float32')`."""
    if _is_cuda_tensor(x):
        import torch

        assert x.dim() == 2 and x.shape[1] == d, f"expected [n, {d}], got {tuple(x.shape)}"
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        return x
    x = np.asarray(x)
    n, dd = x.shape  # ValueError for non-2-D input, as in faiss
    assert dd == d, f"expected [n, {d}], got {x.shape}"
    return np.ascontiguousarray(x, dtype=np.float32)


def _as_i64_vector(ids, n: int, like):
    if _is_cuda_tensor(like):
        import torch

        ids = torch.as_tensor(ids, device=like.device).to(torch.int64).contiguous()
        assert ids.shape == (n,), "not same nb of vectors as ids"
        return ids
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    assert ids.shape == (n,), "not same nb of vectors as ids"
    return ids


def _device_index(x) -> int:
    return x.device.index if x.device.index is not None else 0


class ClusteringParameters:
    """faiss.ClusteringParameters: the fields Clustering::train reads on this path."""

    def __init__(self):
        self.niter = 10
        self.max_points_per_centroid = 256
        self.min_points_per_centroid = 39
        self.seed = 1234
        self.verbose = False
        # normalise the centroids after every iteration.  faiss's struct default is False; its
        # index_factory turns it on for inner-product IVF indexes (mirrored in index_factory below).
        self.spherical = False


class SearchParametersIVF:
    def __init__(self, nprobe: int = 1, max_codes: int = 0):
        self.nprobe = nprobe
        self.max_codes = max_codes


class IndexFlatIP:
    """faiss.IndexFlatIP(d): exact inner-product search."""

    metric_type = METRIC_INNER_PRODUCT
    is_trained = True

    def __init__(self, d: int, device: int = 0):
        self.d = int(d)
        self.device = int(device)
        self._h = c_void_p()
        check(lib().absb_flat_create(self.d, METRIC_INNER_PRODUCT, self.device, byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().absb_flat_destroy(h)

    @property
    def ntotal(self) -> int:
        n = c_int64()
        check(lib().absb_flat_ntotal(self._h, byref(n)))
        return n.value

    def train(self, x):
        _as_f32_matrix(x, self.d)

    def reset(self):
        check(lib().absb_flat_reset(self._h))

    def add(self, x):
        x = _as_f32_matrix(x, self.d)
        if _is_cuda_tensor(x):
            assert _device_index(x) == self.device
            check(lib().absb_flat_add_dev(self._h, x.shape[0], ptr(x), _lib.current_stream_ptr()))
        else:
            check(lib().absb_flat_add(self._h, x.shape[0], ptr(x)))

    def search(self, x, k: int):
        x = _as_f32_matrix(x, self.d)
        assert k > 0
        n = x.shape[0]
        if _is_cuda_tensor(x):
            import torch

            D = torch.empty((n, k), dtype=torch.float32, device=x.device)
            I = torch.empty((n, k), dtype=torch.int64, device=x.device)
            check(lib().absb_flat_search_dev(self._h, n, ptr(x), k, ptr(D), ptr(I), _lib.current_stream_ptr()))
            return D, I
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        check(lib().absb_flat_search(self._h, n, ptr(x), k, ptr(D), ptr(I)))
        return D, I

    def reconstruct_n(self, n0: int = 0, ni: int = -1):
        if ni == -1:
            ni = self.ntotal - n0
        out = np.empty((ni, self.d), dtype=np.float32)
        check(lib().absb_flat_reconstruct(self._h, n0, ni, ptr(out)))
        return out

    def reconstruct(self, key: int):
        return self.reconstruct_n(key, 1)[0]


class _Quantizer:
    """View of an IVF index's coarse quantiser (an IndexFlatIP over the centroids)."""

    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, owner: "IndexIVFFlat"):
        self._o = owner
        self.d = owner.d

    @property
    def ntotal(self) -> int:
        return self._o.nlist if self._o.is_trained else 0

    is_trained = True

    def add(self, centroids):
        c = _as_f32_matrix(centroids, self.d)
        assert c.shape[0] == self._o.nlist, "the quantizer must receive exactly nlist centroids"
        self._o.set_centroids(c)

    def reconstruct_n(self, n0: int = 0, ni: int = -1):
        c = self._o.get_centroids()
        return c[n0:] if ni == -1 else c[n0:n0 + ni]

    def search(self, x, k: int):
        return self._o.coarse(x, k)

    def assign(self, x):
        return self._o.assign(x)


class _InvLists:
    def __init__(self, owner: "IndexIVFFlat"):
        self._o = owner
        self.nlist = owner.nlist
        self.code_size = owner.d * 4

    def list_size(self, l: int) -> int:
        return int(self._o.list_sizes()[l])

    def get_codes(self, l: int) -> np.ndarray:
        return self._o.get_list(l)[0]

    def get_ids(self, l: int) -> np.ndarray:
        return self._o.get_list(l)[1]

    def imbalance_factor(self) -> float:
        s = self._o.list_sizes().astype(np.float64)
        tot = s.sum()
        return float(self.nlist * (s * s).sum() / (tot * tot)) if tot else 0.0


class IndexIVFFlat:
    """faiss.IndexIVFFlat with an IndexFlatIP quantiser, METRIC_INNER_PRODUCT."""

    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int, nlist: int, metric: int = METRIC_INNER_PRODUCT, device: int = 0):
        if metric != METRIC_INNER_PRODUCT:
            raise RuntimeError("only METRIC_INNER_PRODUCT is on the abstracts-search path")
        self.d, self.nlist, self.device = int(d), int(nlist), int(device)
        self.nprobe = 1
        self.cp = ClusteringParameters()
        self._h = c_void_p()
        check(lib().absb_ivf_create(self.d, self.nlist, metric, self.device, byref(self._h)))
        self.quantizer = _Quantizer(self)
        self.invlists = _InvLists(self)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().absb_ivf_destroy(h)

    # ---- state ---------------------------------------------------------------------------
    @property
    def ntotal(self) -> int:
        n = c_int64()
        check(lib().absb_ivf_ntotal(self._h, byref(n)))
        return n.value

    @property
    def is_trained(self) -> bool:
        t = c_int()
        check(lib().absb_ivf_is_trained(self._h, byref(t)))
        return bool(t.value)

    def reset(self):
        check(lib().absb_ivf_reset(self._h))

    def set_shard(self, rank: int, world: int):
        """Keep only inverted lists l with l % world == rank."""
        check(lib().absb_ivf_set_shard(self._h, rank, world))

    def set_tunables(self, scan_chunk: int = -1, coarse_impl: int = -1, scan_ctas_per_sm: int = -1):
        check(lib().absb_ivf_set_tunables(self._h, scan_chunk, coarse_impl, scan_ctas_per_sm))

    def set_two_stage(self, shortlist: int = 64):
        """Two-stage fine scan: fp16 shadow codes (+50% memory, written by add()) give a shortlist, the
        fp32 codes its exact scores; queries the error bound cannot prove fall back to the single-pass
        scan, so results are identical.  Enable before the first add(); 0 switches it off."""
        check(lib().absb_ivf_set_two_stage(self._h, int(shortlist)))
        self.two_stage = int(shortlist)

    def two_stage_fallbacks(self) -> int:
        n = c_int64()
        check(lib().absb_ivf_two_stage_fallbacks(self._h, byref(n)))
        return n.value

    def set_scan_impl(self, impl: int = -1, ring_warps: int = 0, ring_depth: int = 0, ring_stage_vecs: int = 0):
        """Fine-scan kernel: 1 = shared-memory ring fed by the bulk-copy engine (default for d = 1024; small
        CTAs that fit beside the encoder's GEMM CTAs), 0 = register-resident scan.  Same results."""
        check(lib().absb_ivf_set_scan_impl(self._h, int(impl), int(ring_warps), int(ring_depth), int(ring_stage_vecs)))

    def set_scan_order(self, list_major: bool = True):
        """Work-queue order of the fine scan: list-major (default, L2 reuse across queries) or query-major."""
        check(lib().absb_ivf_set_scan_order(self._h, 1 if list_major else 0))

    # ---- train ---------------------------------------------------------------------------
    def train(self, x):
        x = _as_f32_matrix(x, self.d)
        n = x.shape[0]
        cp = self.cp
        if n < self.nlist:
            raise RuntimeError(
                f"Number of training points ({n}) should be at least as large as number of clusters ({self.nlist})")
        if n < self.nlist * cp.min_points_per_centroid:
            warnings.warn(f"WARNING clustering {n} points to {self.nlist} centroids: please provide at least "
                          f"{self.nlist * cp.min_points_per_centroid} training points")
        check(lib().absb_ivf_set_clustering(self._h, cp.niter, cp.max_points_per_centroid,
                                            cp.min_points_per_centroid, cp.seed))
        check(lib().absb_ivf_set_clustering_spherical(self._h, 1 if getattr(cp, "spherical", False) else 0))
        if _is_cuda_tensor(x):
            check(lib().absb_ivf_train_dev(self._h, n, ptr(x), _lib.current_stream_ptr()))
        else:
            check(lib().absb_ivf_train(self._h, n, ptr(x)))

    def set_centroids(self, c):
        c = _as_f32_matrix(c, self.d)
        assert c.shape[0] == self.nlist
        if _is_cuda_tensor(c):
            check(lib().absb_ivf_set_centroids_dev(self._h, ptr(c), _lib.current_stream_ptr()))
        else:
            check(lib().absb_ivf_set_centroids(self._h, ptr(c)))

    def get_centroids(self) -> np.ndarray:
        c = np.empty((self.nlist, self.d), dtype=np.float32)
        check(lib().absb_ivf_get_centroids(self._h, ptr(c)))
        return c

    # ---- add -----------------------------------------------------------------------------
    def add(self, x):
        self._add(x, None, None)

    def add_with_ids(self, x, ids):
        self._add(x, ids, None)

    def add_core(self, x, ids, list_ids):
        """IndexIVF::add_core with precomputed list numbers (-1 skips the row)."""
        self._add(x, ids, list_ids)

    def _add(self, x, ids, list_ids):
        x = _as_f32_matrix(x, self.d)
        n = x.shape[0]
        if ids is not None:
            ids = _as_i64_vector(ids, n, x)
        if list_ids is not None:
            list_ids = _as_i64_vector(list_ids, n, x)
        L = lib()
        if _is_cuda_tensor(x):
            st = _lib.current_stream_ptr()
            if list_ids is None:
                check(L.absb_ivf_add_dev(self._h, n, ptr(x), ptr(ids), st))
            else:
                check(L.absb_ivf_add_preassigned_dev(self._h, n, ptr(x), ptr(ids), ptr(list_ids), st))
        elif list_ids is None:
            check(L.absb_ivf_add(self._h, n, ptr(x), ptr(ids)))
        else:
            check(L.absb_ivf_add_preassigned(self._h, n, ptr(x), ptr(ids), ptr(list_ids)))

    # ---- search --------------------------------------------------------------------------
    def coarse(self, x, nprobe: int):
        x = _as_f32_matrix(x, self.d)
        n = x.shape[0]
        if _is_cuda_tensor(x):
            import torch

            Dc = torch.empty((n, nprobe), dtype=torch.float32, device=x.device)
            Ic = torch.empty((n, nprobe), dtype=torch.int64, device=x.device)
            check(lib().absb_ivf_coarse_dev(self._h, n, ptr(x), nprobe, ptr(Dc), ptr(Ic), _lib.current_stream_ptr()))
            return Dc, Ic
        Dc = np.empty((n, nprobe), dtype=np.float32)
        Ic = np.empty((n, nprobe), dtype=np.int64)
        check(lib().absb_ivf_coarse(self._h, n, ptr(x), nprobe, ptr(Dc), ptr(Ic)))
        return Dc, Ic

    def assign(self, x) -> np.ndarray:
        x = _as_f32_matrix(x, self.d)
        if _is_cuda_tensor(x):
            return self.coarse(x, 1)[1][:, 0]
        out = np.empty((x.shape[0],), dtype=np.int64)
        check(lib().absb_ivf_assign(self._h, x.shape[0], ptr(x), ptr(out)))
        return out

    def search(self, x, k: int, params: SearchParametersIVF | None = None):
        x = _as_f32_matrix(x, self.d)
        assert k > 0
        nprobe = int(params.nprobe if params is not None else self.nprobe)
        nprobe = min(nprobe, self.nlist)
        n = x.shape[0]
        if _is_cuda_tensor(x):
            import torch

            D = torch.empty((n, k), dtype=torch.float32, device=x.device)
            I = torch.empty((n, k), dtype=torch.int64, device=x.device)
            check(lib().absb_ivf_search_dev(self._h, n, ptr(x), k, nprobe, ptr(D), ptr(I), _lib.current_stream_ptr()))
            return D, I
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        check(lib().absb_ivf_search(self._h, n, ptr(x), k, nprobe, ptr(D), ptr(I)))
        return D, I

    def search_preassigned(self, x, k: int, Iq, Dq=None):
        """IndexIVF::search_preassigned: Iq [n, nprobe] list numbers (-1 = none)."""
        x = _as_f32_matrix(x, self.d)
        n = x.shape[0]
        if _is_cuda_tensor(x):
            import torch

            Iq = Iq.to(torch.int64).contiguous()
            assert Iq.dim() == 2 and Iq.shape[0] == n
            D = torch.empty((n, k), dtype=torch.float32, device=x.device)
            I = torch.empty((n, k), dtype=torch.int64, device=x.device)
            check(lib().absb_ivf_search_preassigned_dev(self._h, n, ptr(x), k, Iq.shape[1], ptr(Iq), ptr(D), ptr(I),
                                                        _lib.current_stream_ptr()))
            return D, I
        Iq = np.ascontiguousarray(Iq, dtype=np.int64)
        assert Iq.ndim == 2 and Iq.shape[0] == n
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        check(lib().absb_ivf_search_preassigned(self._h, n, ptr(x), k, Iq.shape[1], ptr(Iq), ptr(D), ptr(I)))
        return D, I

    # ---- inverted lists ------------------------------------------------------------------
    def list_sizes(self) -> np.ndarray:
        s = np.empty((self.nlist,), dtype=np.int64)
        check(lib().absb_ivf_list_sizes(self._h, ptr(s)))
        return s

    def get_list(self, l: int):
        size = int(self.list_sizes()[l])
        codes = np.empty((size, self.d), dtype=np.float32)
        ids = np.empty((size,), dtype=np.int64)
        check(lib().absb_ivf_get_list(self._h, l, ptr(codes), ptr(ids)))
        return codes, ids

    def compact(self, scratch_pages: int = 0) -> None:
        """Make every list's pages physically consecutive, in place (bounded scratch).  Results are
        unchanged; after an index was filled by many small add() calls the scan needs far fewer
        work items.  No faiss counterpart is needed there (its lists are one array each)."""
        if scratch_pages > 0:
            check(lib().absb_ivf_compact_scratch(self._h, int(scratch_pages)))
        else:
            check(lib().absb_ivf_compact(self._h))

    # ---- measurement ---------------------------------------------------------------------
    def last_stats(self) -> dict:
        v, b, w, l = c_int64(), c_int64(), c_int64(), c_int64()
        check(lib().absb_ivf_last_stats(self._h, byref(v), byref(b), byref(w), byref(l)))
        return {"vectors": v.value, "bytes": b.value, "items": w.value, "launches": l.value}

    def set_profile(self, on: int):
        check(lib().absb_ivf_set_profile(self._h, int(on)))

    def get_profile(self) -> dict:
        from ctypes import c_double

        s, c, o, n = c_double(), c_double(), c_double(), c_int64()
        check(lib().absb_ivf_get_profile(self._h, byref(s), byref(c), byref(o), byref(n)))
        s16, n16 = c_double(), c_int64()
        check(lib().absb_ivf_get_profile_scan16(self._h, byref(s16), byref(n16)))
        return {"scan_ms": s.value, "coarse_gemm_ms": c.value, "other_ms": o.value, "scan_launches": n.value,
                "scan16_ms": s16.value, "scan16_launches": n16.value}

    def time_scan(self, iters: int = 10) -> float:
        ms = c_float()
        check(lib().absb_ivf_time_scan(self._h, iters, _lib.current_stream_ptr(), byref(ms)))
        return ms.value


def index_factory(d: int, description: str, metric: int = METRIC_L2, device: int = 0):
    """faiss.index_factory for the descriptions on this path: "Flat" and "IVF<nlist>,Flat"."""
    if metric != METRIC_INNER_PRODUCT:
        raise RuntimeError("only METRIC_INNER_PRODUCT is on the abstracts-search path")
    desc = description.strip()
    if desc == "Flat":
        return IndexFlatIP(d, device=device)
    m = re.fullmatch(r"IVF(\d+),Flat", desc)
    if m:
        ix = IndexIVFFlat(d, int(m.group(1)), metric, device=device)
        # faiss index_factory: `if (metric == METRIC_INNER_PRODUCT) index_ivf->cp.spherical = true` — what
        # `sidecar-search index train` gets (/root/reference/Makefile:38-39).  faiss is external and
        # unpinned here, so this is a switch: set `index.cp.spherical = False` for the plain Lloyd update.
        ix.cp.spherical = True
        return ix
    raise RuntimeError(f"could not parse index description {description!r} (supported: Flat, IVF<n>,Flat)")


def extract_index_ivf(index):
    assert isinstance(index, IndexIVFFlat)
    return index


class ParameterSpace:
    """faiss.ParameterSpace().set_index_parameter(index, "nprobe", v) — what app.py applies from
    params.json (/root/reference/Makefile:12)."""

    def set_index_parameter(self, index, name: str, value):
        if name != "nprobe":
            raise RuntimeError(f"ParameterSpace: unknown parameter {name}")
        index.nprobe = int(value)

    def set_index_parameters(self, index, description: str):
        for tok in filter(None, description.split(",")):
            k, v = tok.split("=")
            self.set_index_parameter(index, k.strip(), v)


# ------------------------------------------------------------------ (de)serialisation ---------
def write_index(index, path: str, ondisk_path: str | None = None) -> None:
    """faiss.write_index in faiss's own byte format (faiss_io.py): IndexFlatIP -> "IxFI",
    IndexIVFFlat -> "IwFl" with array inverted lists, or — with `ondisk_path` — the
    index.faiss + ondisk.ivfdata pair `sidecar-search index fill` leaves (Makefile:11)."""
    from . import faiss_io as fio

    if isinstance(index, IndexFlatIP):
        with open(path, "wb") as f:
            fio.write_flat(f, fio.FlatData(index.d, METRIC_INNER_PRODUCT, index.reconstruct_n(0, index.ntotal)))
        return
    assert isinstance(index, IndexIVFFlat)
    data = fio.IVFFlatData(index.d, index.nlist, index.nprobe, METRIC_INNER_PRODUCT, index.is_trained,
                           index.get_centroids() if index.is_trained else None)
    fio.write_ivfflat(path, data, ondisk_path, sizes=index.list_sizes(), list_fn=lambda l: index.get_list(int(l)))


def read_index(path: str, device: int = 0):
    """faiss.read_index for the index types on this path."""
    from . import faiss_io as fio

    with open(path, "rb") as f:
        cc = f.read(4)
    if cc in (b"IxFI", b"IxF2", b"IxFl"):
        flat = fio._r_flat(fio._R(np.memmap(path, dtype=np.uint8, mode="r")))
        if flat.metric != METRIC_INNER_PRODUCT:
            raise RuntimeError("only METRIC_INNER_PRODUCT is on the abstracts-search path")
        ix = IndexFlatIP(flat.d, device=device)
        ix.add(flat.xb)
        return ix
    data = fio.read_ivfflat(path)
    if data.metric != METRIC_INNER_PRODUCT:
        raise RuntimeError("only METRIC_INNER_PRODUCT is on the abstracts-search path")
    index = IndexIVFFlat(data.d, data.nlist, METRIC_INNER_PRODUCT, device=device)
    index.nprobe = data.nprobe
    if data.centroids is not None:
        index.set_centroids(data.centroids)
    # feed the lists back in bounded batches through add_core (precomputed list numbers)
    batch_c, batch_i, batch_l, rows = [], [], [], 0
    def flush():
        nonlocal batch_c, batch_i, batch_l, rows
        if rows:
            index.add_core(np.concatenate(batch_c), np.concatenate(batch_i), np.concatenate(batch_l))
        batch_c, batch_i, batch_l, rows = [], [], [], 0
    for l in range(data.nlist):
        n = len(data.ids[l])
        if n:
            batch_c.append(data.codes[l]); batch_i.append(data.ids[l]); batch_l.append(np.full(n, l, dtype=np.int64))
            rows += n
            if rows >= 262144:
                flush()
    flush()
    return index
