"""faiss file formats (abstracts-search_b200/faiss_io.py): byte-level structure as restated from
faiss's index_write.cpp — header field widths, fourccs, full/sparse list tables, on-disk lists —
and read(write(x)) == x.  No GPU involved."""
import importlib
import struct

import numpy as np


def _fio():
    return importlib.import_module("abstracts-search_b200.faiss_io")


def _sample(nlist=8, d=4, empty=()):
    rng = np.random.default_rng(0)
    codes, ids, nxt = [], [], 0
    for l in range(nlist):
        n = 0 if l in empty else 1 + (l * 7) % 5
        codes.append(rng.standard_normal((n, d)).astype(np.float32))
        ids.append(np.arange(nxt, nxt + n, dtype=np.int64) * 3 + 1)
        nxt += n
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    return cent, codes, ids


def test_ivfflat_array_lists_layout_and_roundtrip(tmp_path):
    fio = _fio()
    cent, codes, ids = _sample()
    ix = fio.IVFFlatData(4, 8, 5, fio.METRIC_INNER_PRODUCT, True, cent, codes, ids)
    p = str(tmp_path / "index.faiss")
    fio.write_ivfflat(p, ix)
    b = open(p, "rb").read()
    ntotal = sum(len(i) for i in ids)
    # IwFl | d i32 | ntotal i64 | dummy i64 x2 | is_trained u8 | metric i32 | nlist u64 | nprobe u64
    assert b[:4] == b"IwFl"
    assert struct.unpack_from("<iqqqBi", b, 4) == (4, ntotal, 1 << 20, 1 << 20, 1, 0)
    o = 4 + 33
    assert struct.unpack_from("<QQ", b, o) == (8, 5)
    o += 16
    # nested quantiser: IxFI + header + count of floats + data
    assert b[o:o + 4] == b"IxFI"
    assert struct.unpack_from("<iqqqBi", b, o + 4) == (4, 8, 1 << 20, 1 << 20, 1, 0)
    o += 4 + 33
    assert struct.unpack_from("<Q", b, o)[0] == 8 * 4
    assert np.array_equal(np.frombuffer(b, np.float32, 32, o + 8).reshape(8, 4), cent)
    o += 8 + 32 * 4
    # direct map: NoMap + empty array
    assert struct.unpack_from("<BQ", b, o) == (0, 0)
    o += 9
    assert b[o:o + 4] == b"ilar" and struct.unpack_from("<QQ", b, o + 4) == (8, 16) and b[o + 20:o + 24] == b"full"
    assert struct.unpack_from("<Q", b, o + 24)[0] == 8
    sizes = np.frombuffer(b, np.uint64, 8, o + 32)
    assert sizes.tolist() == [len(i) for i in ids]
    o += 32 + 64
    assert np.array_equal(np.frombuffer(b, np.float32, len(ids[0]) * 4, o).reshape(-1, 4), codes[0])
    assert len(b) == o + ntotal * (16 + 8)
    back = fio.read_ivfflat(p)
    assert (back.d, back.nlist, back.nprobe, back.metric, back.is_trained, back.ntotal) == (4, 8, 5, 0, True, ntotal)
    assert np.array_equal(back.centroids, cent)
    for l in range(8):
        assert np.array_equal(back.codes[l], codes[l]) and np.array_equal(back.ids[l], ids[l])


def test_sparse_list_table_and_untrained(tmp_path):
    fio = _fio()
    cent, codes, ids = _sample(empty=(0, 1, 2, 4, 5, 7))
    p = str(tmp_path / "sparse.faiss")
    fio.write_ivfflat(p, fio.IVFFlatData(4, 8, 1, 0, True, cent, codes, ids))
    b = open(p, "rb").read()
    assert b"sprs" in b and b"full" not in b
    back = fio.read_ivfflat(p)
    assert [len(i) for i in back.ids] == [len(i) for i in ids]
    assert np.array_equal(back.codes[3], codes[3]) and np.array_equal(back.ids[6], ids[6])
    # empty.faiss: untrained, no centroids, no vectors
    p2 = str(tmp_path / "empty.faiss")
    fio.write_ivfflat(p2, fio.IVFFlatData(4, 8, 1, 0, False, None))
    e = fio.read_ivfflat(p2)
    assert not e.is_trained and e.centroids is None and e.ntotal == 0 and len(e.ids) == 8


def test_ondisk_inverted_lists(tmp_path):
    fio = _fio()
    cent, codes, ids = _sample(empty=(2,))
    p, dpath = str(tmp_path / "index.faiss"), str(tmp_path / "ondisk.ivfdata")
    fio.write_ivfflat(p, fio.IVFFlatData(4, 8, 3, 0, True, cent, codes, ids), ondisk_path=dpath)
    b = open(p, "rb").read()
    assert b"ilod" in b and b"ondisk.ivfdata" in b and b"ilar" not in b
    ntotal = sum(len(i) for i in ids)
    assert len(open(dpath, "rb").read()) == ntotal * (16 + 8)
    back = fio.read_ivfflat(p)
    assert back.ondisk["filename"] == "ondisk.ivfdata" and back.ondisk["totsize"] == ntotal * 24
    t = back.ondisk["lists"]
    assert t[2].tolist()[:2] == [0, 0] and int(t[0][2]) == 0 and int(t[1][2]) == len(ids[0]) * 24
    for l in range(8):
        assert np.array_equal(back.codes[l], codes[l]) and np.array_equal(back.ids[l], ids[l])


def test_flat_and_ids_parquet(tmp_path):
    fio = _fio()
    xb = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = str(tmp_path / "flat.faiss")
    with open(p, "wb") as f:
        fio.write_flat(f, fio.FlatData(4, fio.METRIC_INNER_PRODUCT, xb))
    b = open(p, "rb").read()
    assert b[:4] == b"IxFI" and len(b) == 4 + 33 + 8 + 48
    back = fio._r_flat(fio._R(np.frombuffer(b, np.uint8)))
    assert np.array_equal(back.xb, xb) and back.metric == 0
    ids = ["https://openalex.org/W1", "https://openalex.org/W22", "https://openalex.org/W333"]
    q = str(tmp_path / "ids.parquet")
    fio.write_ids_parquet(q, ids)
    assert fio.read_ids_parquet(q) == ids
