"""faiss on-disk formats for the indexes on this path.

`sidecar-search index train|fill` leaves `empty.faiss`, `index.faiss` and `ondisk.ivfdata` in the
index directory (/root/reference/Makefile:11-13) and app.py loads them with `faiss.read_index`
(/root/reference/README.md:16).  This module reads and writes those byte layouts directly, so that a
published abstracts-faiss index loads into the B200 index and an index built here is readable by
stock faiss.

Layouts restated from faiss's `impl/index_write.cpp` / `index_read.cpp` (faiss is not vendored in the
reference and not installable offline — the restatement is pinned only by the structural tests in
tests/test_faiss_io_cpu.py, see DESIGN.md "parity unpinned"):

  IndexFlatIP / IndexFlatL2 / IndexFlat        fourcc "IxFI" / "IxF2" / "IxFl"
      header: d i32, ntotal i64, dummy i64 (1<<20), dummy i64, is_trained u8, metric_type i32
      xb:     count u64 (number of float32 values), data
  IndexIVFFlat                                  fourcc "IwFl"
      header (as above), nlist u64, nprobe u64, quantizer (nested index), direct map
      (type u8, array: count u64 + i64 data), inverted lists
  ArrayInvertedLists                            fourcc "ilar"
      nlist u64, code_size u64, list_type fourcc "full" (sizes: count u64 + u64[nlist]) or
      "sprs" (count u64 + (list, size) u64 pairs), then per non-empty list: codes
      [n * code_size] bytes, ids [n] i64
  OnDiskInvertedLists                           fourcc "ilod"
      nlist u64, code_size u64, lists: count u64 (= nlist STRUCTS of 24 bytes, faiss WRITEVECTOR counts
      elements, not u64 words) + {size, capacity, offset} u64 triples, slots: count u64 (16-byte structs)
      + {offset, capacity} u64 pairs, filename: count u64 + bytes, totsize u64;
      in the .ivfdata file list l holds codes [capacity * code_size] at `offset`, then ids
      [capacity] i64.

Pure numpy: nothing here touches the GPU (IndexIVFFlat.to_faiss / from_faiss in index.py do).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_FLAT_FOURCC = {METRIC_INNER_PRODUCT: b"IxFI", METRIC_L2: b"IxF2"}


@dataclass
class FlatData:
    d: int
    metric: int
    xb: np.ndarray  # [ntotal, d] float32
    is_trained: bool = True


@dataclass
class IVFFlatData:
    d: int
    nlist: int
    nprobe: int
    metric: int
    is_trained: bool
    centroids: np.ndarray | None  # [nlist, d] float32 (None when the quantiser is empty)
    codes: list = field(default_factory=list)  # per list [n, d] float32
    ids: list = field(default_factory=list)  # per list [n] int64
    ondisk: dict | None = None  # {"filename", "lists" [nlist,3] u64, "totsize"} when the lists are "ilod"

    @property
    def ntotal(self) -> int:
        return int(sum(len(i) for i in self.ids))


# ------------------------------------------------------------------ writer ---------------------
def _w_header(f, d: int, ntotal: int, is_trained: bool, metric: int):
    f.write(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1 if is_trained else 0, metric))


def _w_vector(f, a: np.ndarray):
    a = np.ascontiguousarray(a)
    f.write(struct.pack("<Q", a.size))
    f.write(a.tobytes())


def write_flat(f, flat: FlatData):
    f.write(_FLAT_FOURCC[flat.metric])
    xb = np.ascontiguousarray(flat.xb, dtype=np.float32).reshape(-1, flat.d)
    _w_header(f, flat.d, xb.shape[0], flat.is_trained, flat.metric)
    _w_vector(f, xb.reshape(-1))


class _Lists:
    """Uniform view of the lists to write: either materialised (codes, ids) or a callback that
    fetches one list at a time (a 106 GB shard never sits in host memory)."""

    def __init__(self, ix: "IVFFlatData", sizes=None, list_fn=None):
        self.nlist, self.d = ix.nlist, ix.d
        if list_fn is not None:
            self.sizes = np.asarray(sizes, dtype=np.uint64)
            self.fn = list_fn
        else:
            codes = ix.codes or [np.zeros((0, ix.d), np.float32)] * ix.nlist
            ids = ix.ids or [np.zeros(0, np.int64)] * ix.nlist
            self.sizes = np.asarray([len(i) for i in ids], dtype=np.uint64)
            self.fn = lambda l: (codes[l], ids[l])

    def __iter__(self):
        for l in range(self.nlist):
            if self.sizes[l]:
                c, i = self.fn(l)
                yield l, np.ascontiguousarray(c, dtype=np.float32), np.ascontiguousarray(i, dtype=np.int64)
            else:
                yield l, np.zeros((0, self.d), np.float32), np.zeros(0, np.int64)


def _w_array_invlists(f, d: int, lists: "_Lists"):
    nlist = lists.nlist
    f.write(b"ilar")
    f.write(struct.pack("<QQ", nlist, d * 4))
    sizes = lists.sizes
    n_non0 = int((sizes > 0).sum())
    if n_non0 > nlist // 2:
        f.write(b"full")
        _w_vector(f, sizes)
    else:
        f.write(b"sprs")
        nz = np.nonzero(sizes)[0].astype(np.uint64)
        _w_vector(f, np.stack([nz, sizes[nz.astype(np.int64)]], axis=1).reshape(-1))
    for _, c, i in lists:
        if len(i):
            f.write(c.tobytes())
            f.write(i.tobytes())


def write_ivfflat(path: str, ix: IVFFlatData, ondisk_path: str | None = None, sizes=None, list_fn=None):
    """index.faiss (+ ondisk.ivfdata when `ondisk_path` is given: lists go to that file, the index
    file keeps only their table — the layout `sidecar-search index fill` produces, Makefile:11)."""
    lists = _Lists(ix, sizes, list_fn)
    with open(path, "wb") as f:
        f.write(b"IwFl")
        _w_header(f, ix.d, int(lists.sizes.sum()), ix.is_trained, ix.metric)
        f.write(struct.pack("<QQ", ix.nlist, ix.nprobe))
        cent = ix.centroids if ix.centroids is not None else np.zeros((0, ix.d), np.float32)
        write_flat(f, FlatData(ix.d, ix.metric, cent))
        f.write(struct.pack("<B", 0))  # DirectMap::NoMap
        _w_vector(f, np.zeros(0, dtype=np.int64))
        if ondisk_path is None:
            _w_array_invlists(f, ix.d, lists)
            return
        code_size = ix.d * 4
        table = np.zeros((ix.nlist, 3), dtype=np.uint64)  # size, capacity, offset
        off = 0
        with open(ondisk_path, "wb") as g:
            for l, c, i in lists:
                n = len(i)
                table[l] = (n, n, off if n else np.uint64(0xFFFFFFFFFFFFFFFF))
                if n:
                    g.write(c.tobytes())
                    g.write(i.tobytes())
                    off += n * (code_size + 8)
        f.write(b"ilod")
        f.write(struct.pack("<QQ", ix.nlist, code_size))
        # WRITEVECTOR(od->lists): the count is the number of 24-byte OnDiskOneList structs
        f.write(struct.pack("<Q", ix.nlist))
        f.write(table.tobytes())
        f.write(struct.pack("<Q", 0))  # no free slots (count of 16-byte Slot structs)
        name = os.path.basename(ondisk_path).encode()
        f.write(struct.pack("<Q", len(name)))
        f.write(name)
        f.write(struct.pack("<Q", off))


# ------------------------------------------------------------------ reader ---------------------
class _R:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, fmt: str):
        v = struct.unpack_from(fmt, self.b, self.o)
        self.o += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def fourcc(self) -> bytes:
        v = bytes(self.b[self.o:self.o + 4])
        self.o += 4
        return v

    def vector(self, dtype) -> np.ndarray:
        n = self.take("<Q")
        dt = np.dtype(dtype)
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o)
        self.o += n * dt.itemsize
        return a


def _r_header(r: _R):
    d, ntotal, _, _, trained, metric = r.take("<iqqqBi")
    if metric > 1:
        r.take("<f")  # metric_arg
    return d, ntotal, bool(trained), metric


def _r_flat(r: _R) -> FlatData:
    cc = r.fourcc()
    if cc not in (b"IxFI", b"IxF2", b"IxFl"):
        raise RuntimeError(f"unsupported quantizer fourcc {cc!r} (only IndexFlat is on the abstracts-search path)")
    d, ntotal, trained, metric = _r_header(r)
    xb = r.vector(np.float32)
    if xb.size != ntotal * d:
        raise RuntimeError("IndexFlat: vector count does not match the header")
    return FlatData(d, metric, xb.reshape(ntotal, d), trained)


def read_ivfflat(path: str, ondisk_dir: str | None = None) -> IVFFlatData:
    buf = np.memmap(path, dtype=np.uint8, mode="r")
    r = _R(buf)
    cc = r.fourcc()
    if cc != b"IwFl":
        raise RuntimeError(f"{path}: fourcc {cc!r} is not an IndexIVFFlat ('IwFl')")
    d, ntotal, trained, metric = _r_header(r)
    nlist, nprobe = r.take("<QQ")
    q = _r_flat(r)
    dm_type = r.take("<B")
    r.vector(np.int64)
    if dm_type == 2:  # DirectMap::Hashtable: vector of (id, lo) pairs
        n = r.take("<Q")
        r.o += n * 16
    out = IVFFlatData(d, nlist, nprobe, metric, trained, q.xb.copy() if q.xb.shape[0] else None)
    il = r.fourcc()
    if il == b"il00":
        out.codes = [np.zeros((0, d), np.float32) for _ in range(nlist)]
        out.ids = [np.zeros(0, np.int64) for _ in range(nlist)]
        return out
    nl, code_size = r.take("<QQ")
    if nl != nlist or code_size != d * 4:
        raise RuntimeError("inverted lists do not match an IVF,Flat index of this dimension")
    if il == b"ilar":
        lt = r.fourcc()
        sizes = np.zeros(nlist, dtype=np.int64)
        if lt == b"full":
            sizes[:] = r.vector(np.uint64).astype(np.int64)
        elif lt == b"sprs":
            pairs = r.vector(np.uint64).astype(np.int64).reshape(-1, 2)
            sizes[pairs[:, 0]] = pairs[:, 1]
        else:
            raise RuntimeError(f"unknown list type {lt!r}")
        for n in sizes:
            n = int(n)
            c = np.frombuffer(buf, dtype=np.float32, count=n * d, offset=r.o).reshape(n, d)
            r.o += n * code_size
            i = np.frombuffer(buf, dtype=np.int64, count=n, offset=r.o)
            r.o += n * 8
            out.codes.append(c)
            out.ids.append(i)
    elif il == b"ilod":
        n_structs = r.take("<Q")
        if n_structs != nlist:
            raise RuntimeError(f"OnDiskInvertedLists: {n_structs} list records for nlist = {nlist}")
        table = np.frombuffer(buf, dtype=np.uint64, count=3 * nlist, offset=r.o).reshape(nlist, 3)
        r.o += 24 * nlist
        nslots = r.take("<Q")
        r.o += nslots * 16
        name = bytes(r.vector(np.uint8)).decode()
        totsize = r.take("<Q")
        data_path = os.path.join(ondisk_dir or os.path.dirname(os.path.abspath(path)), os.path.basename(name))
        # an index without vectors has a zero-byte ondisk.ivfdata, which cannot be mapped
        any_rows = bool((table[:, 0] > 0).any())
        data = np.memmap(data_path, dtype=np.uint8, mode="r") if any_rows else None
        out.ondisk = {"filename": name, "lists": table, "totsize": totsize}
        for size, cap, off in table.astype(np.int64):
            size, cap, off = int(size), int(cap), int(off)
            if size == 0:
                out.codes.append(np.zeros((0, d), np.float32))
                out.ids.append(np.zeros(0, np.int64))
                continue
            out.codes.append(np.frombuffer(data, dtype=np.float32, count=size * d, offset=off).reshape(size, d))
            out.ids.append(np.frombuffer(data, dtype=np.int64, count=size, offset=off + cap * code_size))
    else:
        raise RuntimeError(f"unsupported inverted-list fourcc {il!r}")
    if out.ntotal != ntotal:
        raise RuntimeError("inverted lists hold a different number of vectors than the header says")
    return out


# ------------------------------------------------------------------ ids.parquet ----------------
def write_ids_parquet(path: str, ids):
    """faiss row number -> document id (OpenAlex work id), row i of the table = faiss id i
    (`ids.parquet`, /root/reference/Makefile:11; README.md:16)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    pq.write_table(pa.table({"id": pa.array(list(ids))}), path)


def read_ids_parquet(path: str):
    import pyarrow.parquet as pq

    return pq.read_table(path).column(0).to_pylist()
