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


def _hand_built_faiss_ondisk(tmp_path, free_slots=((4096, 64),)):
    """index.faiss + ondisk.ivfdata assembled field by field the way faiss's own writers emit them
    (write_index -> write_ivf_header -> write_InvertedLists -> OnDiskInvertedListsIOHook::write; every
    WRITEVECTOR is `size_t count of ELEMENTS` + raw elements) — nothing here calls faiss_io's writer.
    Lists have capacity > size and sit at non-monotonic offsets, as after faiss's in-place growth."""
    d, nlist, nprobe = 4, 3, 2
    rng = np.random.default_rng(5)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    sizes, caps = [2, 0, 3], [4, 0, 3]
    codes = [rng.standard_normal((n, d)).astype(np.float32) for n in sizes]
    ids = [np.arange(10 * l, 10 * l + n, dtype=np.int64) for l, n in enumerate(sizes)]
    code_size = d * 4
    # data file: list 2 first, then a hole, then list 0 (codes [capacity * code_size], ids [capacity] i64)
    offs = [256, 0xFFFFFFFFFFFFFFFF, 0]
    tot = 256 + caps[0] * (code_size + 8)
    data = bytearray(tot)
    for l in (0, 2):
        o = offs[l]
        data[o:o + sizes[l] * code_size] = codes[l].tobytes()
        io = o + caps[l] * code_size
        data[io:io + sizes[l] * 8] = ids[l].tobytes()
    (tmp_path / "ondisk.ivfdata").write_bytes(bytes(data))
    hdr = lambda nt: struct.pack("<iqqqBi", d, nt, 1 << 20, 1 << 20, 1, 0)  # noqa: E731  d, ntotal, dummies, trained, metric
    b = b"IwFl" + hdr(sum(sizes)) + struct.pack("<QQ", nlist, nprobe)
    b += b"IxFI" + hdr(nlist) + struct.pack("<Q", nlist * d) + cent.tobytes()  # quantizer: xb as floats
    b += struct.pack("<B", 0) + struct.pack("<Q", 0)  # DirectMap NoMap + empty array
    b += b"ilod" + struct.pack("<QQ", nlist, code_size)
    b += struct.pack("<Q", nlist)  # WRITEVECTOR(od->lists): nlist structs of {size, capacity, offset}
    for l in range(nlist):
        b += struct.pack("<QQQ", sizes[l], caps[l], offs[l])
    b += struct.pack("<Q", len(free_slots))  # WRITEVECTOR(slots): structs of {offset, capacity}
    for o, c in free_slots:
        b += struct.pack("<QQ", o, c)
    name = b"/some/build/dir/ondisk.ivfdata"  # faiss stores the path it was written with
    b += struct.pack("<Q", len(name)) + name + struct.pack("<Q", tot)
    (tmp_path / "index.faiss").write_bytes(b)
    return str(tmp_path / "index.faiss"), cent, codes, ids, b


def test_reads_hand_built_faiss_ondisk_file(tmp_path):
    """the 'ilod' list table is nlist 24-byte structs (count = nlist, not 3 * nlist)."""
    fio = _fio()
    p, cent, codes, ids, _ = _hand_built_faiss_ondisk(tmp_path)
    back = fio.read_ivfflat(p)
    assert (back.d, back.nlist, back.nprobe, back.ntotal) == (4, 3, 2, 5)
    assert np.array_equal(back.centroids, cent)
    for l in range(3):
        assert np.array_equal(back.codes[l], codes[l]) and np.array_equal(back.ids[l], ids[l])
    assert back.ondisk["filename"].endswith("ondisk.ivfdata") and back.ondisk["lists"].shape == (3, 3)


def test_writer_emits_the_hand_built_ondisk_layout(tmp_path):
    """Our writer's index.faiss for the same lists parses with an independent field-by-field walk that
    uses faiss's element counts (and is byte-identical to the hand-built file where layouts coincide)."""
    fio = _fio()
    d, nlist = 4, 3
    rng = np.random.default_rng(6)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    sizes = [2, 0, 3]
    codes = [rng.standard_normal((n, d)).astype(np.float32) for n in sizes]
    ids = [np.arange(10 * l, 10 * l + n, dtype=np.int64) for l, n in enumerate(sizes)]
    p, dp = str(tmp_path / "index.faiss"), str(tmp_path / "ondisk.ivfdata")
    fio.write_ivfflat(p, fio.IVFFlatData(d, nlist, 7, 0, True, cent, codes, ids), ondisk_path=dp)
    b = open(p, "rb").read()
    o = b.index(b"ilod") + 4
    assert struct.unpack_from("<QQ", b, o) == (nlist, d * 4)
    o += 16
    assert struct.unpack_from("<Q", b, o)[0] == nlist, "lists vector must count structs, not u64 words"
    o += 8
    table = np.frombuffer(b, np.uint64, 3 * nlist, o).reshape(nlist, 3)
    o += 24 * nlist
    assert table[:, 0].tolist() == sizes and table[1, 2] == 0xFFFFFFFFFFFFFFFF
    assert struct.unpack_from("<Q", b, o)[0] == 0  # no free slots
    o += 8
    n_name = struct.unpack_from("<Q", b, o)[0]
    assert b[o + 8:o + 8 + n_name] == b"ondisk.ivfdata"
    o += 8 + n_name
    assert struct.unpack_from("<Q", b, o)[0] == sum(sizes) * (d * 4 + 8) and o + 8 == len(b)
    raw = open(dp, "rb").read()
    for l in (0, 2):
        off = int(table[l, 2])
        assert raw[off:off + sizes[l] * 16] == codes[l].tobytes()
        assert raw[off + int(table[l, 1]) * 16: off + int(table[l, 1]) * 16 + sizes[l] * 8] == ids[l].tobytes()


def test_roundtrip_fuzz_array_sparse_and_ondisk(tmp_path):
    """hypothesis: random (d, nlist, list sizes incl. many empty lists, metric, nprobe) through every container —
    `ilar` full / sparse (the writer picks by fill ratio, as faiss does) and `ilod` + ondisk.ivfdata — must read
    back identically, and the three files of one index must agree with each other."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    fio = _fio()
    counter = [0]

    @settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(d=st.sampled_from([1, 3, 4, 16, 33]), nlist=st.integers(1, 40), fill=st.floats(0.0, 1.0),
           max_len=st.integers(1, 9), metric=st.sampled_from([0, 1]), nprobe=st.integers(1, 64), seed=st.integers(0, 2**31))
    def run(d, nlist, fill, max_len, metric, nprobe, seed):
        rng = np.random.default_rng(seed)
        codes, ids, nxt = [], [], 0
        for _ in range(nlist):
            n = int(rng.integers(1, max_len + 1)) if rng.random() < fill else 0
            codes.append(rng.standard_normal((n, d)).astype(np.float32))
            ids.append(rng.permutation(np.arange(nxt, nxt + n, dtype=np.int64)) + int(rng.integers(0, 1 << 40)))
            nxt += n
        cent = rng.standard_normal((nlist, d)).astype(np.float32)
        ix = fio.IVFFlatData(d, nlist, nprobe, metric, True, cent, codes, ids)
        counter[0] += 1
        sub = tmp_path / f"case{counter[0]}"
        sub.mkdir()
        p_arr, p_od, dpath = str(sub / "array.faiss"), str(sub / "index.faiss"), str(sub / "ondisk.ivfdata")
        fio.write_ivfflat(p_arr, ix)
        fio.write_ivfflat(p_od, ix, ondisk_path=dpath)
        for back in (fio.read_ivfflat(p_arr), fio.read_ivfflat(p_od)):
            assert (back.d, back.nlist, back.nprobe, back.metric, back.is_trained) == (d, nlist, nprobe, metric, True)
            assert back.ntotal == nxt and np.array_equal(back.centroids, cent)
            for l in range(nlist):
                assert back.codes[l].shape == (len(ids[l]), d)
                assert np.array_equal(back.codes[l], codes[l]) and np.array_equal(back.ids[l], ids[l])
        import os

        assert os.path.getsize(dpath) == nxt * (4 * d + 8)

    run()


def test_ondisk_index_without_vectors(tmp_path):
    """A trained index written with on-disk lists before any add(): ondisk.ivfdata is a zero-byte file (found by the
    fuzz above: it cannot be memory-mapped) and every list reads back empty."""
    fio = _fio()
    cent = np.arange(12, dtype=np.float32).reshape(3, 4)
    empty = [np.zeros((0, 4), np.float32)] * 3, [np.zeros(0, np.int64)] * 3
    p, dpath = str(tmp_path / "index.faiss"), str(tmp_path / "ondisk.ivfdata")
    fio.write_ivfflat(p, fio.IVFFlatData(4, 3, 1, 0, True, cent, *empty), ondisk_path=dpath)
    import os

    assert os.path.getsize(dpath) == 0
    back = fio.read_ivfflat(p)
    assert back.ntotal == 0 and [len(i) for i in back.ids] == [0, 0, 0] and np.array_equal(back.centroids, cent)
