"""GPU parity tests of the index half of the path: CUDA (through the C-ABI / Python surface) vs
the CPU oracle, on seeded inputs, golden fixtures and size-independent properties.
Integer/index results must be bit-exact; fp32 scores are bit-exact on the exact-lattice corpus and
within 1e-5 absolute on gaussian data (tolerance stated per test)."""
import ctypes

import numpy as np
import pytest

from conftest import golden
from oracle import ivf as oivf
from oracle import synth as osynth

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)


def _torch():
    import torch

    return torch


# ------------------------------------------------------------------ generator -----------------
def test_device_generator_matches_oracle_bit_for_bit(gpu_pkg):
    P = gpu_pkg
    for d, nlist, n, row0 in [(1024, 128, 300, 0), (64, 16, 1000, 12345), (256, 65536, 200, 10**9)]:
        x = P.synth.corpus(1234, row0, n, d, nlist).cpu().numpy()
        assert np.array_equal(x, osynth.corpus(1234, row0, n, d, nlist))
        c = P.synth.cluster_of(1234, row0, n, nlist).cpu().numpy()
        assert np.array_equal(c, osynth.cluster_of(1234, np.arange(row0, row0 + n), nlist))
    c = P.synth.centroids(1234, 128, 1024).cpu().numpy()
    assert np.array_equal(c, osynth.centroids(1234, 128, 1024))
    q = P.synth.queries(1234, 5, 77, 1024, 128, 8192).cpu().numpy()
    assert np.array_equal(q, osynth.queries(1234, 5, 77, 1024, 128, 8192))


def test_unit_norm_generator_matches_oracle_bit_for_bit(gpu_pkg):
    """kind | 4: integer rows divided by their (exact) L2 norm with IEEE sqrt / divide — real-valued fp32 unit
    vectors that the numpy twin reproduces bit for bit, and that fp16 cannot hold exactly."""
    P = gpu_pkg
    t = _torch()
    for d, nlist, n, row0 in [(1024, 128, 300, 0), (64, 16, 1000, 12345), (1024, 65536, 500, 10**9)]:
        x = P.synth.corpus(1234, row0, n, d, nlist, unit=True).cpu().numpy()
        assert np.array_equal(x, osynth.corpus_unit(1234, row0, n, d, nlist))
        assert np.abs(np.linalg.norm(x.astype(np.float64), axis=1) - 1).max() < 1e-6
    q = P.synth.queries(1234, 5, 77, 1024, 128, 8192, unit=True).cpu().numpy()
    assert np.array_equal(q, osynth.queries_unit(1234, 5, 77, 1024, 128, 8192))
    rows = t.tensor([5, 99, 3, 10**10 + 7], dtype=t.int64, device="cuda")
    xr = P.synth.corpus_rows(1234, rows, 1024, 128, unit=True).cpu().numpy()
    assert np.array_equal(xr, osynth.corpus_rows_unit(1234, rows.cpu().numpy(), 1024, 128))
    x = osynth.corpus_unit(1234, 0, 64, 1024, 128)
    assert (x.astype(np.float16).astype(np.float32) != x).mean() > 0.5  # lossy in fp16, unlike the lattice rows


# ------------------------------------------------------------------ golden: lattice ------------
@pytest.fixture(scope="module")
def lattice(gpu_pkg):
    g = golden("ivf_lattice_d1024.npz")
    d, nlist, n, nq = int(g["d"]), int(g["nlist"]), int(g["n"]), int(g["nq"])
    seed = int(g["seed"])
    x = osynth.corpus(seed, 0, n, d, nlist)
    q = osynth.queries(seed, 0, nq, d, nlist, n)
    c = osynth.centroids(seed, nlist, d)
    return g, x, q, c


def test_flat_search_matches_golden(gpu_pkg, lattice):
    g, x, q, _ = lattice
    ix = gpu_pkg.IndexFlatIP(x.shape[1])
    ix.add(x[:5000])
    ix.add(x[5000:])
    assert ix.ntotal == x.shape[0]
    D, I = ix.search(q, int(g["k"]))
    assert np.array_equal(I, g["If"]) and np.array_equal(D, g["Df"])
    t = _torch()
    Dd, Id = ix.search(t.from_numpy(q).cuda(), int(g["k"]))
    assert np.array_equal(Id.cpu().numpy(), g["If"]) and np.array_equal(Dd.cpu().numpy(), g["Df"])
    assert np.array_equal(ix.reconstruct_n(10, 3), x[10:13])


@pytest.mark.parametrize("coarse_impl", [0, 1])
def test_ivf_matches_golden_lattice(gpu_pkg, lattice, coarse_impl):
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = gpu_pkg.index_factory(d, f"IVF{nlist},Flat", gpu_pkg.METRIC_INNER_PRODUCT)
    ix.set_tunables(coarse_impl=coarse_impl)
    assert not ix.is_trained
    ix.set_centroids(c)
    assert ix.is_trained and ix.ntotal == 0
    ix.add(x)
    assert ix.ntotal == x.shape[0]
    assert np.array_equal(ix.list_sizes(), g["sizes"])
    assert np.array_equal(ix.assign(x), g["assign"])
    Dc, Ic = ix.coarse(q, nprobe)
    assert np.array_equal(Ic, g["Ic"]) and np.array_equal(Dc, g["Dc"])
    ix.nprobe = nprobe
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]), "top-k ids differ from the oracle"
    assert np.array_equal(D, g["D"]), "lattice scores must be bit-exact"
    # device-tensor entry, SearchParametersIVF, search_preassigned
    t = _torch()
    Dd, Id = ix.search(t.from_numpy(q).cuda(), k, params=gpu_pkg.SearchParametersIVF(nprobe=nprobe))
    assert np.array_equal(Id.cpu().numpy(), g["I"]) and np.array_equal(Dd.cpu().numpy(), g["D"])
    Dp, Ip = ix.search_preassigned(q, k, g["Ic"])
    assert np.array_equal(Ip, g["I"]) and np.array_equal(Dp, g["D"])
    # work-queue order (list-major with L2 reuse vs query-major) must not change anything
    ix.set_scan_order(False)
    Dq, Iq = ix.search(q, k)
    ix.set_scan_order(True)
    assert np.array_equal(Iq, g["I"]) and np.array_equal(Dq, g["D"])
    # probing every list == exact flat search
    ix.nprobe = nlist
    Da, Ia = ix.search(q, k)
    assert np.array_equal(Ia, g["If"]) and np.array_equal(Da, g["Df"])
    st = ix.last_stats()
    assert st["vectors"] == q.shape[0] * x.shape[0] and st["bytes"] == st["vectors"] * (4 * d + 8)


def test_list_contents_keep_insertion_order_across_adds(gpu_pkg, lattice):
    g, x, q, c = lattice
    d, nlist = int(g["d"]), int(g["nlist"])
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.set_centroids(c)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    for a, b in [(0, 1000), (1000, 1001), (1001, 5000), (5000, 8192)]:
        ix.add(x[a:b])
        o.add(x[a:b])
    assert np.array_equal(ix.list_sizes(), o.list_sizes())
    for l in (0, 17, nlist - 1):
        codes, ids = ix.get_list(l)
        assert np.array_equal(ids, o.ids[l]) and np.array_equal(codes, o.codes[l])
    ix.nprobe = int(g["nprobe"])
    D, I = ix.search(q, int(g["k"]))
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    ix.reset()
    assert ix.ntotal == 0 and ix.is_trained and ix.list_sizes().sum() == 0


@pytest.mark.parametrize("scratch_pages", [0, 1, 37])
def test_compact_after_many_small_adds_changes_nothing_but_the_work_items(gpu_pkg, lattice, scratch_pages):
    """Streaming fill (row-group-sized add() calls) scatters every list over the page pool; compact()
    permutes the pages in place.  List contents, ids and search results must stay bit-identical; the
    scan must need fewer work items afterwards; later add() calls must still work."""
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.set_centroids(c)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    n0 = 7000
    for a in range(0, n0, 250):
        ix.add(x[a:a + 250])
    o.add(x[:n0])
    ix.nprobe = nprobe
    D0, I0 = ix.search(q, k)
    items_before = ix.last_stats()["items"]
    Do, Io = o.search(q, k, nprobe=nprobe)
    assert np.array_equal(I0, Io) and np.array_equal(D0, Do)
    ix.compact(scratch_pages)
    assert np.array_equal(ix.list_sizes(), o.list_sizes())
    for l in range(0, nlist, max(1, nlist // 16)):
        codes, ids = ix.get_list(l)
        assert np.array_equal(ids, o.ids[l]) and np.array_equal(codes, o.codes[l])
    D1, I1 = ix.search(q, k)
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)
    assert ix.last_stats()["items"] < items_before
    ix.compact(scratch_pages)  # already compact: a no-op
    ix.add(x[n0:])
    o.add(x[n0:])
    D2, I2 = ix.search(q, k)
    assert np.array_equal(I2, g["I"]) and np.array_equal(D2, g["D"])
    for l in (0, nlist - 1):
        codes, ids = ix.get_list(l)
        assert np.array_equal(ids, o.ids[l]) and np.array_equal(codes, o.codes[l])


# ------------------------------------------------------------------ golden: gaussian -----------
def test_ivf_gauss_search_and_train(gpu_pkg):
    g = golden("ivf_gauss_d64.npz")
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    x, q = g["x"], g["q"]
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.set_centroids(g["centroids"])
    ix.add(x)
    agree = (ix.assign(x) == g["assign"]).mean()
    assert agree > 0.999
    ix.nprobe = nprobe
    D, I = ix.search(q, k)
    # fp32 summation order differs from the oracle's: demand exact ids wherever the fp64 margins
    # exceed the fp32 rounding bound (1e-5 on unit vectors), scores within 1e-5
    safe = (g["coarse_margin"] > 1e-5) & (g["fine_margin"] > 1e-5)
    assert safe.sum() >= 30
    assert np.array_equal(I[safe], g["I"][safe])
    assert np.abs(D[safe] - g["D"][safe]).max() < 1e-5
    # Index.train: same subsample/init/Lloyd iterations as the oracle
    ix2 = gpu_pkg.IndexIVFFlat(d, nlist)
    ix2.train(x)
    c = ix2.get_centroids()
    # a point whose two best centroids tie within fp32 rounding may switch cluster between two
    # correct fp32 implementations; demand near-identity, not bit-identity
    diff = np.abs(c - g["centroids"]).max(axis=1)
    assert np.median(diff) < 1e-5 and (diff < 1e-4).mean() >= 0.8 and diff.max() < 5e-2
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    assert (o.assign(x) == g["assign"]).mean() > 0.99


def test_train_matches_oracle_on_lattice(gpu_pkg):
    d, nlist, n = 64, 16, 4000
    x = osynth.corpus(99, 0, n, d, nlist)
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.train(x)
    ref = oivf.kmeans_train(x, nlist)
    diff = np.abs(ix.get_centroids() - ref).max(axis=1)
    assert np.median(diff) < 1e-5 and diff.max() < 5e-2
    # subsampling branch: n > nlist * max_points_per_centroid
    ix3 = gpu_pkg.IndexIVFFlat(d, 4)
    ix3.cp.max_points_per_centroid = 256
    x3 = osynth.corpus(5, 0, 4 * 300, d, 4)
    ix3.train(x3)
    diff = np.abs(ix3.get_centroids() - oivf.kmeans_train(x3, 4)).max(axis=1)
    assert np.median(diff) < 1e-5 and diff.max() < 5e-2
    with pytest.raises(RuntimeError):
        gpu_pkg.IndexIVFFlat(d, 64).train(x[:10])


def test_spherical_train_matches_oracle_and_factory_default(gpu_pkg):
    """ClusteringParameters.spherical: centroids re-normalised after the initial draw and after every
    iteration (faiss post_process_centroids).  index_factory switches it on for inner product (faiss's
    own factory does; external and unpinned, hence a switch), the bare constructor leaves it off."""
    P = gpu_pkg
    d, nlist, n = 64, 16, 4000
    x = osynth.corpus_unit(99, 0, n, d, nlist)
    fac = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    assert fac.cp.spherical is True and P.IndexIVFFlat(d, nlist).cp.spherical is False
    fac.train(x)
    c = fac.get_centroids()
    assert np.abs(np.linalg.norm(c, axis=1) - 1).max() < 1e-5
    ref = oivf.kmeans_train(x, nlist, spherical=True)
    diff = np.abs(c - ref).max(axis=1)
    assert np.median(diff) < 1e-5 and diff.max() < 5e-2
    plain = oivf.kmeans_train(x, nlist, spherical=False)
    assert np.abs(np.linalg.norm(plain, axis=1) - 1).max() > 1e-3  # the two settings really differ
    fac2 = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    fac2.cp.spherical = False
    fac2.train(x)
    diff = np.abs(fac2.get_centroids() - plain).max(axis=1)
    assert np.median(diff) < 1e-5 and diff.max() < 5e-2
    # device-tensor entry takes the same path
    t = _torch()
    fac3 = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    fac3.train(t.from_numpy(x).cuda())
    assert np.array_equal(fac3.get_centroids(), c)


# ------------------------------------------------------------------ edge cases -----------------
def test_empty_ragged_and_padded_results(gpu_pkg):
    P = gpu_pkg
    d, nlist = 64, 8
    c = osynth.centroids(1, nlist, d)
    ix = P.IndexIVFFlat(d, nlist)
    q = osynth.queries(1, 0, 3, d, nlist, 6)
    with pytest.raises(RuntimeError):
        ix.add(q)  # not trained
    ix.set_centroids(c)
    ix.nprobe = 4
    D, I = ix.search(q, 5)  # empty index
    assert (I == -1).all() and (D == -FLT_MAX).all()
    x = osynth.corpus(1, 0, 6, d, nlist)
    ix.add(x)
    ix.nprobe = 100  # > nlist: clamped like faiss
    D, I = ix.search(q, 10)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    o.add(x)
    Do, Io = o.search(q, 10, nprobe=nlist)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert (I[:, 6:] == -1).all() and (D[:, 6:] == -FLT_MAX).all()
    # zero queries, -1 coarse entries
    D0, I0 = ix.search(np.zeros((0, d), np.float32), 3)
    assert D0.shape == (0, 3) and I0.shape == (0, 3)
    Dp, Ip = ix.search_preassigned(q, 4, np.full((3, 2), -1, dtype=np.int64))
    assert (Ip == -1).all()
    # argument contract
    with pytest.raises(AssertionError):
        ix.search(np.zeros((2, d + 1), np.float32), 3)
    with pytest.raises(RuntimeError):
        ix.search(q, 100000)
    with pytest.raises(AssertionError):
        ix.add_with_ids(x, np.arange(5))
    D64, I64 = ix.search(q.astype(np.float64), 3)  # coerced like faiss's wrapper
    assert np.array_equal(I64, Io[:, :3])


def test_ties_resolve_by_id_and_large_k(gpu_pkg):
    P = gpu_pkg
    d, nlist = 64, 4
    c = osynth.centroids(2, nlist, d)
    row = osynth.corpus(2, 0, 1, d, nlist)
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_centroids(c)
    ix.add_with_ids(np.repeat(row, 7, axis=0), np.array([50, 3, 99, 7, 21, 1, 64]))
    ix.nprobe = nlist
    D, I = ix.search(row, 5)
    assert I[0].tolist() == [1, 3, 7, 21, 50]
    # k spanning every register-slot variant of the warp top-k
    n = 3000
    x = osynth.corpus(3, 0, n, d, nlist)
    q = osynth.queries(3, 0, 9, d, nlist, n)
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_centroids(c := osynth.centroids(3, nlist, d))
    ix.add(x)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    o.add(x)
    for k in (1, 32, 33, 100, 256):
        ix.nprobe = 2
        D, I = ix.search(q, k)
        Do, Io = o.search(q, k, nprobe=2, impl="c")
        assert np.array_equal(I, Io) and np.array_equal(D, Do), k


def test_write_read_roundtrip(gpu_pkg, tmp_path, lattice):
    g, x, q, c = lattice
    ix = gpu_pkg.IndexIVFFlat(int(g["d"]), int(g["nlist"]))
    ix.set_centroids(c)
    ix.add(x)
    ix.nprobe = int(g["nprobe"])
    path = str(tmp_path / "ix.absb")
    gpu_pkg.write_index(ix, path)
    ix2 = gpu_pkg.read_index(path)
    assert ix2.ntotal == ix.ntotal and ix2.nprobe == ix.nprobe
    D, I = ix2.search(q, int(g["k"]))
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    assert open(path, "rb").read(4) == b"IwFl"  # faiss's own container (faiss_io.py)
    # index.faiss + ondisk.ivfdata pair, as `sidecar-search index fill` leaves it (Makefile:11)
    p3, d3 = str(tmp_path / "index.faiss"), str(tmp_path / "ondisk.ivfdata")
    gpu_pkg.write_index(ix, p3, ondisk_path=d3)
    ix3 = gpu_pkg.read_index(p3)
    assert np.array_equal(ix3.list_sizes(), ix.list_sizes())
    D, I = ix3.search(q, int(g["k"]))
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    # flat index
    fl = gpu_pkg.IndexFlatIP(x.shape[1])
    fl.add(x[:777])
    pf = str(tmp_path / "flat.faiss")
    gpu_pkg.write_index(fl, pf)
    fl2 = gpu_pkg.read_index(pf)
    assert fl2.ntotal == 777 and np.array_equal(fl2.reconstruct_n(0, 777), x[:777])


# ------------------------------------------------------------------ shards ---------------------
def test_sharded_by_list_equals_single_index(gpu_pkg, lattice):
    P = gpu_pkg
    t = _torch()
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    world = 3
    parts = []
    for r in range(world):
        ix = P.IndexIVFFlat(d, nlist)
        ix.set_shard(r, world)
        ix.set_centroids(c)
        ix.add(x[:3000])
        ix.add(x[3000:])
        ix.nprobe = nprobe
        parts.append(ix)
    assert sum(p.ntotal for p in parts) == x.shape[0]
    sizes = sum(p.list_sizes() for p in parts)
    assert np.array_equal(sizes, g["sizes"])
    for r, p in enumerate(parts):
        assert (p.list_sizes()[np.arange(nlist) % world != r] == 0).all()
    qd = t.from_numpy(q).cuda()
    res = [p.search(qd, k) for p in parts]
    # dense [world, n, k] layout
    D_all = t.stack([r[0] for r in res]).contiguous()
    I_all = t.stack([r[1] for r in res]).contiguous()
    Dm = t.empty((q.shape[0], k), dtype=t.float32, device="cuda")
    Im = t.empty((q.shape[0], k), dtype=t.int64, device="cuda")
    L = P.lib()
    rc = L.absb_merge_shards_dev(0, world, q.shape[0], k, ctypes.c_void_p(D_all.data_ptr()),
                                 ctypes.c_void_p(I_all.data_ptr()), 0, ctypes.c_void_p(Dm.data_ptr()),
                                 ctypes.c_void_p(Im.data_ptr()), None)
    assert rc == 0
    t.cuda.synchronize()
    assert np.array_equal(Im.cpu().numpy(), g["I"]) and np.array_equal(Dm.cpu().numpy(), g["D"])
    Dh, Ih = P.merge_partials_host(D_all.cpu().numpy(), I_all.cpu().numpy(), k)
    assert np.array_equal(Ih, g["I"]) and np.array_equal(Dh, g["D"])


@pytest.mark.parametrize("ctas", [1, 2])
def test_scan_occupancy_cap_changes_no_result(gpu_pkg, lattice, ctas):
    """The resident-CTAs-per-SM cap of the fine scan is a performance knob (used by the encode/search
    overlap experiment, profiles/r01r_pipeline_experiment.md); ids and scores stay bit-exact."""
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.set_centroids(c)
    ix.add(x[:5000])
    ix.add(x[5000:])
    ix.set_tunables(scan_ctas_per_sm=ctas)
    ix.nprobe = nprobe
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])


# ------------------------------------------------------------------ two-stage scan -------------
@pytest.mark.parametrize("geometry", [(1, 2, 1), (4, 3, 1), (4, 4, 2), (8, 6, 1), (3, 2, 2)])
def test_ring_scan_equals_register_scan_bit_for_bit(gpu_pkg, lattice, geometry):
    """ivf_scan_ring.cu (list vectors staged in shared memory by cp.async.bulk, per-warp rings) against
    ivf_scan.cu / ivf_scan16.cu (vectors in registers): same lane -> element mapping, FMA chain and
    butterfly, so ids AND score bits must agree on real-valued rows too — for every ring geometry
    (warps, stages, vectors per stage), single-pass and two-stage, with ragged lists (many shorter than
    one stage, many not a multiple of it), empty lists, k = 1 / 10 / 100, and against the golden."""
    P = gpu_pkg
    warps, depth, sv = geometry
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_two_stage(64)
    ix.set_centroids(c)
    for a in range(0, x.shape[0], 3001):  # several adds: lists are page runs, not one block
        ix.add(x[a:a + 3001])
    ix.nprobe = nprobe
    for compacted in (False, True):
        if compacted:
            ix.compact()
        for two_stage in (0, 64):
            ix.set_two_stage(two_stage)
            ix.set_scan_impl(1, warps, depth, sv)
            D1, I1 = ix.search(q, k)
            assert np.array_equal(I1, g["I"]) and np.array_equal(D1, g["D"]), (compacted, two_stage)
    # real-valued rows, ragged + empty lists
    rng = np.random.default_rng(11)
    nl, n = 64, 20000
    xr = rng.standard_normal((n, d)).astype(np.float32)
    xr /= np.linalg.norm(xr, axis=1, keepdims=True)
    lists = rng.integers(0, nl - 8, n)  # the last 8 lists stay empty
    lists[:37] = nl - 9                # and one list holds at least 37 rows in the first add
    cr = rng.standard_normal((nl, d)).astype(np.float32)
    qr = xr[rng.choice(n, 50, replace=False)] + 0.02 * rng.standard_normal((50, d)).astype(np.float32)
    res = {}
    for impl in (0, 1):
        for two_stage in (0, 32):
            iy = P.IndexIVFFlat(d, nl)
            if two_stage:
                iy.set_two_stage(two_stage)
            iy.set_scan_impl(impl, warps, depth, sv)
            iy.set_centroids(cr)
            for a in range(0, n, 4999):
                iy.add_core(xr[a:a + 4999], np.arange(a, min(n, a + 4999)), lists[a:a + 4999])
            iy.nprobe = 48
            for kk in (1, 10, 100):
                res[(impl, two_stage, kk)] = iy.search(qr.astype(np.float32), kk)
    for two_stage in (0, 32):
        for kk in (1, 10, 100):
            if two_stage and kk > two_stage:
                continue
            a, b = res[(0, 0, kk)], res[(1, two_stage, kk)]
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (two_stage, kk)
            a2 = res[(0, two_stage, kk)]
            assert np.array_equal(a[1], a2[1]) and np.array_equal(a[0], a2[0]), (two_stage, kk)


def test_coresident_ring_scan_shape_matches_golden(gpu_pkg, lattice):
    """scan impl 2: the register-capped (96) ring kernel, one 8-warp CTA per SM, 2 stages of 4 KB per warp —
    the shape that shares an SM with the encoder's GEMM CTAs.  Bit-exact to the golden, both scan modes."""
    P = gpu_pkg
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_two_stage(64)
    ix.set_centroids(c)
    for a in range(0, x.shape[0], 2500):
        ix.add(x[a:a + 2500])
    ix.nprobe = nprobe
    ix.set_scan_impl(2)
    for two_stage in (64, 0, 32):
        ix.set_two_stage(two_stage)
        D, I = ix.search(q, k)
        assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"]), two_stage
    ix.compact()
    ix.set_tunables(scan_chunk=512)
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])


@pytest.mark.parametrize("shortlist", [32, 64, 128])
def test_two_stage_scan_matches_golden_lattice(gpu_pkg, lattice, shortlist):
    """fp16 shortlist + exact fp32 re-score + bound check/fallback (ivf_scan16.cu): ids and scores are
    bit-identical to the single-pass scan's golden result, before and after compaction, and for a k
    larger than the shortlist (which silently takes the single-pass scan)."""
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    ix = gpu_pkg.IndexIVFFlat(d, nlist)
    ix.set_two_stage(shortlist)
    ix.set_centroids(c)
    ix.add(x[:3000])
    ix.add(x[3000:])
    ix.nprobe = nprobe
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    assert 0 <= ix.two_stage_fallbacks() <= q.shape[0]
    ix.compact()
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    ref = gpu_pkg.IndexIVFFlat(d, nlist)
    ref.set_centroids(c)
    ref.add(x)
    ref.nprobe = nprobe
    for kk in (1, shortlist, shortlist + 8):
        Dr, Ir = ref.search(q, kk)
        D2, I2 = ix.search(q, kk)
        assert np.array_equal(I2, Ir) and np.array_equal(D2, Dr), kk
    ix.set_two_stage(0)
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    with pytest.raises(gpu_pkg.AbsbError):
        ref.set_two_stage(64)  # the shadow codes are written by add(): too late now
    with pytest.raises(gpu_pkg.AbsbError):
        ix.set_two_stage(48)


def test_two_stage_scan_on_gaussian_rows_equals_single_pass_bit_for_bit(gpu_pkg):
    """Unit-norm gaussian rows are NOT exactly representable in fp16: the shortlist is approximate, the
    re-score uses the single-pass kernel on the fp32 codes, the bound decides per query.  Every query
    must come out identical (ids and score bits) to an index without the shadow codes."""
    rng = np.random.default_rng(5)
    d, nlist, n, nq, k, nprobe = 1024, 32, 30000, 200, 10, 6
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    c = x[rng.choice(n, nlist, replace=False)].copy()
    q = x[rng.choice(n, nq, replace=False)] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    res = {}
    for mode in (0, 32, 128):
        ix = gpu_pkg.IndexIVFFlat(d, nlist)
        if mode:
            ix.set_two_stage(mode)
        ix.set_centroids(c)
        for a in range(0, n, 7000):
            ix.add(x[a:a + 7000])
        ix.nprobe = nprobe
        res[mode] = ix.search(q, k)
        if mode:
            # measured bound: rows are iid gaussian directions, so the K-th approximate score sits within
            # a few 1e-3 of the k-th exact one while the fp16 error bound is ~3e-4 * |q|: most queries are
            # proven with K = 32 and essentially all with K = 128.  A regression that sent every query to
            # the fallback would keep the results right and this assertion is what catches it.
            fb = ix.two_stage_fallbacks()
            assert fb <= (nq // 4 if mode == 32 else nq // 20), f"shortlist {mode}: {fb} of {nq} queries fell back"
    for mode in (32, 128):
        assert np.array_equal(res[mode][1], res[0][1]) and np.array_equal(res[mode][0], res[0][0]), mode


def test_two_stage_scan_falls_back_when_ties_defeat_the_bound(gpu_pkg):
    """40 distinct rows, each stored 100 times: rank 10 and rank 32 have the SAME score, so the bound
    can never separate the shortlist from the rest; those queries must take the single-pass fallback
    and still return the smallest ids among the ties."""
    rng = np.random.default_rng(9)
    d, nlist, k = 1024, 8, 10
    base = rng.standard_normal((40, d)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    x = np.repeat(base, 100, axis=0)[rng.permutation(4000)]
    c = base[:nlist].copy()
    q = base[:16].copy()
    ref = gpu_pkg.IndexIVFFlat(d, nlist)
    two = gpu_pkg.IndexIVFFlat(d, nlist)
    two.set_two_stage(32)
    for ix in (ref, two):
        ix.set_centroids(c)
        ix.add(x)
        ix.nprobe = nlist
    Dr, Ir = ref.search(q, k)
    D2, I2 = two.search(q, k)
    assert two.two_stage_fallbacks() == 16
    assert np.array_equal(I2, Ir) and np.array_equal(D2, Dr)
    assert (np.diff(Ir, axis=1) > 0).all(), "ties must come out in ascending id order"


def test_peer_exchange_allgather_emulated_ranks(gpu_pkg):
    """csrc/peer.cuh protocol with 4 ranks emulated on one GPU (buffers wired by raw pointer): every
    rank pushes into every buffer, then every rank waits and reads its own ring entry.  Five epochs
    exercise the two-deep ring; an early wait on a missing push is not tested (it would spin)."""
    t = _torch()
    world, n = 4, 1024
    pxs = gpu_pkg.PeerExchange.emulate(0, world, n * 4)
    for epoch in range(5):
        src = [t.full((n,), float(100 * epoch + r), device="cuda") + t.arange(n, device="cuda") for r in range(world)]
        for r in range(world):
            pxs[r].push(src[r])
        for r in range(world):
            got = pxs[r].wait(n * 4).contiguous().view(t.float32).view(world, n)
            assert t.equal(got, t.stack(src)), (epoch, r)
    assert all(p.status() == 0 for p in pxs)
    with pytest.raises(gpu_pkg.AbsbError):
        pxs[0].push(t.zeros(n * 2, device="cuda"))  # larger than the slot


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_through_peer_exchange_equals_single_index(gpu_pkg, lattice, world):
    """Fused exchange: absb_ivf_search_push_dev stores every shard's merged top-k into all ranks'
    buffers, absb_peer_merge_shards_dev waits inside the kernel and merges.  Ranks are emulated on
    one GPU (all pushes are issued before the first wait); ids and scores must be bit-exact to the
    single-index golden result on every rank, for repeated batches (ring reuse) and a ragged one."""
    P = gpu_pkg
    t = _torch()
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    nq = q.shape[0]
    parts = []
    for r in range(world):
        ix = P.IndexIVFFlat(d, nlist)
        ix.set_shard(r, world)
        ix.set_centroids(c)
        ix.add(x)
        parts.append(ix)
    pxs = P.PeerExchange.emulate(0, world, nq * k * 12 + 16)
    L = P.lib()
    qd = t.from_numpy(q).cuda()
    for n_use in (nq, nq, 7, nq):
        for r in range(world):
            assert L.absb_ivf_search_push_dev(parts[r]._h, pxs[r]._h, n_use, ctypes.c_void_p(qd.data_ptr()), k, nprobe, None) == 0
        for r in range(world):
            Dm = t.empty((n_use, k), dtype=t.float32, device="cuda")
            Im = t.empty((n_use, k), dtype=t.int64, device="cuda")
            assert L.absb_peer_merge_shards_dev(pxs[r]._h, n_use, k, ctypes.c_void_p(Dm.data_ptr()),
                                                ctypes.c_void_p(Im.data_ptr()), None) == 0
            t.cuda.synchronize()
            assert np.array_equal(Im.cpu().numpy(), g["I"][:n_use]) and np.array_equal(Dm.cpu().numpy(), g["D"][:n_use]), (n_use, r)
    assert all(p.status() == 0 for p in pxs)
    # a record that does not fit the slot is refused, not truncated
    small = P.PeerExchange.emulate(0, 1, 64)[0]
    assert L.absb_ivf_search_push_dev(parts[0]._h, small._h, nq, ctypes.c_void_p(qd.data_ptr()), k, nprobe, None) != 0


@pytest.mark.parametrize("shortlist", [32, 64])
def test_two_stage_results_through_peer_exchange_equal_single_index(gpu_pkg, lattice, shortlist):
    """Two-stage shards push from their LAST merge kernel (proven rows copied from the local result,
    fallback rows merged from the single-pass partials): absb_ivf_search_push_dev on a two-stage index,
    in-kernel wait + merge on the other side.  Two ranks emulated on one GPU; bit-exact to the golden
    for repeated batches, a ragged one, and — with duplicated rows that defeat the bound — through the
    fallback rows as well."""
    P = gpu_pkg
    t = _torch()
    g, x, q, c = lattice
    d, nlist, nprobe, k = int(g["d"]), int(g["nlist"]), int(g["nprobe"]), int(g["k"])
    nq, world = q.shape[0], 2
    parts = []
    for r in range(world):
        ix = P.IndexIVFFlat(d, nlist)
        ix.set_two_stage(shortlist)
        ix.set_shard(r, world)
        ix.set_centroids(c)
        ix.add(x)
        ix.nprobe = nprobe
        parts.append(ix)
    pxs = P.PeerExchange.emulate(0, world, nq * k * 12 + 32)
    L = P.lib()
    qd = t.from_numpy(q).cuda()
    for n_use in (nq, 5, nq):
        for r in range(world):
            assert L.absb_ivf_search_push_dev(parts[r]._h, pxs[r]._h, n_use, ctypes.c_void_p(qd.data_ptr()), k, nprobe, None) == 0
        for r in range(world):
            Dm = t.empty((n_use, k), dtype=t.float32, device="cuda")
            Im = t.empty((n_use, k), dtype=t.int64, device="cuda")
            assert L.absb_peer_merge_shards_dev(pxs[r]._h, n_use, k, ctypes.c_void_p(Dm.data_ptr()),
                                                ctypes.c_void_p(Im.data_ptr()), None) == 0
            t.cuda.synchronize()
            assert np.array_equal(Im.cpu().numpy(), g["I"][:n_use]) and np.array_equal(Dm.cpu().numpy(), g["D"][:n_use])
    assert all(p.status() == 0 for p in pxs)
    # fallback rows through the fused push: every row stored 80 times (> shortlist) -> the K-th approximate
    # score ties with the k-th exact one and the bound cannot separate them
    rng = np.random.default_rng(3)
    base = rng.standard_normal((30, d)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    xd = np.repeat(base, 80, axis=0)[rng.permutation(2400)]
    cd, qq = base[:8].copy(), base[:12].copy()
    ref = P.IndexIVFFlat(d, 8)
    ref.set_centroids(cd)
    ref.add(xd)
    ref.nprobe = 8
    Dr, Ir = ref.search(qq, k)
    parts = []
    for r in range(world):
        ix = P.IndexIVFFlat(d, 8)
        ix.set_two_stage(shortlist)
        ix.set_shard(r, world)
        ix.set_centroids(cd)
        ix.add(xd)
        parts.append(ix)
    pxs = P.PeerExchange.emulate(0, world, 12 * k * 12 + 32)
    qd = t.from_numpy(qq).cuda()
    for r in range(world):
        assert L.absb_ivf_search_push_dev(parts[r]._h, pxs[r]._h, 12, ctypes.c_void_p(qd.data_ptr()), k, 8, None) == 0
    for r in range(world):
        Dm = t.empty((12, k), dtype=t.float32, device="cuda")
        Im = t.empty((12, k), dtype=t.int64, device="cuda")
        assert L.absb_peer_merge_shards_dev(pxs[r]._h, 12, k, ctypes.c_void_p(Dm.data_ptr()), ctypes.c_void_p(Im.data_ptr()), None) == 0
        t.cuda.synchronize()
        assert np.array_equal(Im.cpu().numpy(), Ir) and np.array_equal(Dm.cpu().numpy(), Dr)
    assert sum(p.two_stage_fallbacks() for p in parts) > 0, "the duplicated rows were meant to force fallback rows"


def test_merge_shards_packed_record_with_odd_result_count(gpu_pkg):
    """n * k odd -> the packed {I i64 | D f32} record of each rank must keep its int64 block
    8-byte aligned (segments padded to 16 bytes); an unpadded stride is refused instead of faulting."""
    P = gpu_pkg
    t = _torch()
    world, n, k = 3, 3, 5
    rng = np.random.default_rng(0)
    D_all = np.sort(rng.standard_normal((world, n, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    I_all = rng.permutation(world * n * k).reshape(world, n, k).astype(np.int64)
    i_bytes = (n * k * 8 + 15) & ~15
    rec = i_bytes + ((n * k * 4 + 15) & ~15)
    buf = t.zeros(world * rec, dtype=t.uint8, device="cuda")
    for w in range(world):
        buf[w * rec: w * rec + n * k * 8].view(t.int64).copy_(t.from_numpy(I_all[w].reshape(-1)))
        buf[w * rec + i_bytes: w * rec + i_bytes + n * k * 4].view(t.float32).copy_(t.from_numpy(D_all[w].reshape(-1)))
    Dm = t.empty((n, k), dtype=t.float32, device="cuda")
    Im = t.empty((n, k), dtype=t.int64, device="cuda")
    L = P.lib()
    base = buf.data_ptr()
    assert L.absb_merge_shards_dev(0, world, n, k, ctypes.c_void_p(base + i_bytes), ctypes.c_void_p(base), rec,
                                   ctypes.c_void_p(Dm.data_ptr()), ctypes.c_void_p(Im.data_ptr()), None) == 0
    t.cuda.synchronize()
    Dw, Iw = P.merge_partials_host(D_all, I_all, k)
    assert np.array_equal(Im.cpu().numpy(), Iw) and np.array_equal(Dm.cpu().numpy(), Dw)
    # the r1 layout (stride n*k*12 = 180, odd ranks 4-byte aligned) is now an error, not a misaligned load
    assert L.absb_merge_shards_dev(0, world, n, k, ctypes.c_void_p(base + n * k * 8), ctypes.c_void_p(base), n * k * 12,
                                   ctypes.c_void_p(Dm.data_ptr()), ctypes.c_void_p(Im.data_ptr()), None) != 0


# ------------------------------------------------------------------ larger, property-based -----
def test_two_million_rows_properties(gpu_pkg):
    """2M x 1024 built from the device generator with precomputed list ids (the faiss add_core
    path): (1) probing all lists of a query's shortlist equals exact flat search restricted to those
    lists; (2) the oracle, rebuilding only the probed lists from the counter-based generator, agrees
    bit-for-bit; (3) searching twice is idempotent; (4) scores are sorted, ids unique."""
    P = gpu_pkg
    t = _torch()
    d, nlist, n, nq, nprobe, k = 1024, 8192, 2_000_000, 64, 32, 10
    seed = 1234
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_centroids(P.synth.centroids(seed, nlist, d))
    step = 250_000
    for r0 in range(0, n, step):
        xb = P.synth.corpus(seed, r0, step, d, nlist)
        lb = P.synth.cluster_of(seed, r0, step, nlist)
        ix.add_core(xb, t.arange(r0, r0 + step, device="cuda"), lb)
    del xb, lb
    assert ix.ntotal == n
    sizes = ix.list_sizes()
    assert sizes.sum() == n
    assert np.array_equal(sizes, np.bincount(osynth.cluster_of(seed, np.arange(n), nlist), minlength=nlist))
    q = P.synth.queries(seed, 0, nq, d, nlist, n)
    ix.nprobe = nprobe
    D, I = ix.search(q, k)
    D2, I2 = ix.search(q, k)
    assert t.equal(D, D2) and t.equal(I, I2)
    Dn, In = D.cpu().numpy(), I.cpu().numpy()
    assert (np.diff(Dn, axis=1) <= 0).all()
    assert all(len(set(r.tolist())) == k for r in In)
    _, Ic = ix.coarse(q, nprobe)
    Ic = Ic.cpu().numpy()
    qn = q.cpu().numpy()
    # oracle on the first 6 queries: regenerate only their probed lists
    sel = np.arange(6)
    lists = np.unique(Ic[sel])
    rows = osynth.rows_of_lists(seed, n, nlist, lists)
    cent = osynth.centroids(seed, nlist, d)
    Dco, Ico = oivf.FlatIP(d), None
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(cent)
    for l in lists:
        r = rows[int(l)]
        o.add(osynth.corpus_rows(seed, r, d, nlist), ids=r, list_ids=np.full(len(r), l))
    _, Ico = o.coarse(qn[sel], nprobe)
    assert np.array_equal(Ico, Ic[sel])
    Do, Io = o.search_preassigned(qn[sel], k, Ic[sel], impl="c")
    assert np.array_equal(Io, In[sel]) and np.array_equal(Do, Dn[sel])
    st = ix.last_stats()
    assert st["vectors"] == sizes[Ic].sum()


@pytest.mark.parametrize("corpus", ["lattice", "unit"])
def test_baseline_config2_ivf65536_ten_million_rows(gpu_pkg, corpus):
    """BASELINE configs[2]: IVF65536,Flat over 10M x 1024 rows, nprobe 32, k 10 on one GPU (avg 153-vector
    lists: the short-list regime) — the nlist the metric is quoted on.  The oracle regenerates only the
    probed lists of a query sample from oracle/synth.py and runs its own coarse + scan (C restatement).
      lattice: coarse ids, top-k ids AND scores bit-exact, single-pass and two-stage;
      unit (real-valued L2-normalised rows, lossy in fp16): ids exact wherever the fp64 margin exceeds
      the fp32 rounding bound 2e-5, scores within 2e-5, two-stage == single-pass bit for bit with a
      stated fallback bound (<= 5% of the queries)."""
    P = gpu_pkg
    t = _torch()
    import gc

    d, nlist, n, nq, nprobe, k, seed = 1024, 65536, 10_000_000, 128, 32, 10, 1234
    unit = corpus == "unit"
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    ix.set_two_stage(64)
    ix.set_centroids(P.synth.centroids(seed, nlist, d))
    step = 500_000
    for r0 in range(0, n, step):
        xb = P.synth.corpus(seed, r0, step, d, nlist, unit=unit)
        lb = P.synth.cluster_of(seed, r0, step, nlist)
        ix.add_core(xb, t.arange(r0, r0 + step, device="cuda"), lb)
    del xb, lb
    ix.compact()
    assert ix.ntotal == n
    sizes = ix.list_sizes()
    assert np.array_equal(sizes, np.bincount(osynth.cluster_of(seed, np.arange(n), nlist), minlength=nlist))
    q = P.synth.queries(seed, 0, nq, d, nlist, n, unit=unit)  # perturbed rows of this corpus
    ix.nprobe = nprobe
    D2, I2 = ix.search(q, k)
    fb = ix.two_stage_fallbacks()
    st = ix.last_stats()
    ix.set_two_stage(0)
    D1, I1 = ix.search(q, k)
    assert t.equal(D1, D2) and t.equal(I1, I2), "two-stage and single-pass scans disagree"
    assert fb <= nq // 20, f"{fb} of {nq} queries fell back to the single-pass scan"
    Dn, In = D1.cpu().numpy(), I1.cpu().numpy()
    assert (np.diff(Dn, axis=1) <= 0).all() and (In >= 0).all()
    _, Ic = ix.coarse(q, nprobe)
    Ic = Ic.cpu().numpy()
    assert st["vectors"] == sizes[Ic].sum()
    assert st["items"] <= 3 * nq * nprobe  # compacted short lists: one or two work items per probe
    # ---- oracle on a sample ----
    sel = np.arange(16)
    qn = q.cpu().numpy()[sel]
    cent = osynth.centroids(seed, nlist, d)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(cent)
    _, Ico = o.coarse(qn, nprobe, impl="c")
    c64 = cent.astype(np.float64) @ qn.astype(np.float64).T  # [nlist, 16]
    rows = osynth.rows_of_lists(seed, n, nlist, np.unique(Ico))
    gen = osynth.corpus_rows_unit if unit else osynth.corpus_rows
    for l, r in rows.items():
        if len(r):
            o.add(gen(seed, r, d, nlist), ids=r, list_ids=np.full(len(r), l))
    Do, Io = o.search_preassigned(qn, k, Ico, impl="c")
    off, codes, ids_csr = o._as_csr()
    checked = 0
    for i in range(len(sel)):
        cs = np.sort(c64[:, i])[::-1]
        if cs[nprobe - 1] - cs[nprobe] <= 2e-5 or np.min(cs[:nprobe - 1] - cs[1:nprobe]) <= 2e-5:
            continue  # coarse near-tie: two correct fp32 coarse searches may order / cut differently
        assert np.array_equal(Ic[i], Ico[i]), i
        if not unit:
            assert np.array_equal(In[i], Io[i]) and np.array_equal(Dn[i], Do[i]), i  # exact arithmetic: bit for bit
        else:
            assert np.abs(Dn[i] - Do[i]).max() <= 2e-5, i
            if not np.array_equal(In[i], Io[i]):
                # only a near-tie may reorder ids between two correct fp32 scans: at every rank the two ids'
                # fp64 scores must agree within the fp32 rounding of a 1024-term unit-vector product
                sc = np.concatenate([codes[off[l]:off[l + 1]].astype(np.float64) @ qn[i].astype(np.float64) for l in Ico[i]])
                idv = np.concatenate([ids_csr[off[l]:off[l + 1]] for l in Ico[i]])
                s64 = dict(zip(idv.tolist(), sc.tolist()))
                assert all(abs(s64[int(a)] - s64[int(b)]) <= 2e-6 for a, b in zip(In[i], Io[i])), i
                continue
        checked += 1
    assert checked >= 12, f"only {checked} of 16 sample queries came out id-identical"
    del ix, D1, I1, D2, I2, q
    gc.collect()
    t.cuda.empty_cache()


# ------------------------------------------------------------------ distributed build (NCCL) ---
def test_distributed_build_two_gpus(gpu_pkg):
    """train_distributed / add_distributed over NCCL (tests/dist_build_check.py under torchrun).
    Needs two visible GPUs; the single-GPU driver run skips it, `gpurun --gpus 2` exercises it."""
    import os
    import subprocess
    import sys

    t = _torch()
    if t.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "tests", "dist_build_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dist_build_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


# ------------------------------------------------------------------ §8f rows on the GPU index --
def test_parquet_fill_and_tune_on_gpu_index(gpu_pkg, tmp_path, lattice):
    P = gpu_pkg
    g, x, q, c = lattice
    d, nlist, k = int(g["d"]), int(g["nlist"]), int(g["k"])
    ids = [f"W{i}" for i in range(x.shape[0])]
    P.store.write_shards(str(tmp_path / "data"), ids, x, shard_size=3000, row_group_size=1024)
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    ix.set_centroids(c)
    assert P.store.fill_index(ix, str(tmp_path / "data"), ids_parquet=str(tmp_path / "ids.parquet")) == x.shape[0]
    assert np.array_equal(ix.list_sizes(), g["sizes"])
    ix.nprobe = int(g["nprobe"])
    D, I = ix.search(q, k)
    assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    names = P.faiss_io.read_ids_parquet(str(tmp_path / "ids.parquet"))
    assert names[int(I[0, 0])] == f"W{int(I[0, 0])}"
    # tune: exact ground truth from a flat index; recall grows with nprobe and reaches 1 at nlist
    flat = P.IndexFlatIP(d)
    flat.add(x)
    pts = P.tune.sweep(ix, q, k, nprobes=[1, 4, int(g["nprobe"]), nlist], ground_truth=flat, repeats=1)
    rec = [p["recall"] for p in pts]
    assert rec == sorted(rec) and rec[-1] == 1.0
    choice = P.tune.tune(ix, q, k, min_recall=0.9, nprobes=[1, 4, 16, nlist], ground_truth=flat,
                         params_path=str(tmp_path / "params.json"))
    assert choice["recall"] >= 0.9 and ix.nprobe == choice["nprobe"]


def test_tune_recall_against_the_oracles_exact_search(gpu_pkg, tmp_path):
    """`index tune` (/root/reference/Makefile:27-32): the recall@k the sweep reports must be the recall against
    EXACT search as the oracle computes it (fp64 all-pairs scores over the same rows) — not merely against the
    product's own flat index.  Real-valued unit-norm rows; the product's IndexFlatIP (tcgen05 split-bf16 GEMM)
    must itself return the oracle's exact top-k wherever the fp64 margin is clear."""
    P = gpu_pkg
    d, nlist, n, nq, k = 1024, 64, 30000, 96, 10
    x = osynth.corpus_unit(21, 0, n, d, nlist)
    q = osynth.queries_unit(21, 0, nq, d, nlist, n)
    # half of the queries are random directions: their true neighbours are spread over many lists, so recall
    # really depends on nprobe (a perturbed row finds its whole top-10 in its own list)
    rq = np.random.default_rng(2).standard_normal((nq // 2, d)).astype(np.float32)
    q[nq // 2:] = rq / np.linalg.norm(rq, axis=1, keepdims=True)
    s64 = q.astype(np.float64) @ x.astype(np.float64).T
    order = np.argsort(-s64, axis=1, kind="stable")[:, :k + 1]
    top = np.take_along_axis(s64, order, axis=1)
    clear = (top[:, :-1] - top[:, 1:]).min(axis=1) > 2e-6
    I_true = order[:, :k]
    flat = P.IndexFlatIP(d)
    flat.add(x)
    Df, If = flat.search(q, k)
    assert clear.sum() >= nq // 2
    assert np.array_equal(If[clear], I_true[clear])
    assert np.abs(Df - np.take_along_axis(s64, If, axis=1)).max() < 2e-6
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    ix.set_centroids(osynth.centroids(21, nlist, d))
    ix.add(x)
    nprobes = [1, 2, 4, 8, nlist]
    pts = P.tune.sweep(ix, q, k, nprobes=nprobes, ground_truth=flat, repeats=1)
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(osynth.centroids(21, nlist, d))
    o.add(x)
    for p_, pt in zip(nprobes, pts):
        _, Io = o.search(q, k, nprobe=p_, impl="c")
        want = np.mean([len(np.intersect1d(a, b)) / k for a, b in zip(Io[clear], I_true[clear])])
        ix.nprobe = p_
        _, Ig = ix.search(q, k)
        got = np.mean([len(np.intersect1d(a, b)) / k for a, b in zip(Ig[clear], I_true[clear])])
        assert abs(got - want) < 1e-9, (p_, got, want)          # product recall == oracle recall vs exact fp64 truth
        assert abs(pt["recall"] - P.tune.recall_at_k(Ig, If)) < 1e-12  # and the sweep reports what it measured
    assert pts[-1]["recall"] == 1.0 and pts[0]["recall"] < pts[-1]["recall"]
    choice = P.tune.tune(ix, q, k, min_recall=0.95, nprobes=nprobes, ground_truth=flat, params_path=str(tmp_path / "params.json"))
    assert choice["recall"] >= 0.95


def test_flat_search_ragged_sizes_on_the_tensor_path(gpu_pkg):
    """IndexFlatIP over row counts that are not multiples of the GEMM's 32-column granule or of the 65,536-row
    chunk, lattice rows (every score exact): ids and scores bit-exact to the oracle, incremental adds."""
    P = gpu_pkg
    d = 1024
    ix = P.IndexFlatIP(d)
    o = oivf.FlatIP(d)
    q = osynth.queries(3, 0, 37, d, 64, 70000)
    done = 0
    for n_add in (1, 30, 4097, 65536 - 4128 + 5, 1000):
        x = osynth.corpus(3, done, n_add, d, 64)
        ix.add(x)
        o.add(x)
        done += n_add
        for k in (1, 10):
            if k > done:
                continue
            D, I = ix.search(q, k)
            Do, Io = o.search(q, k, impl="c")
            assert np.array_equal(I, Io) and np.array_equal(D, Do), (done, k)
    D, I = ix.search(q[:1], 100)
    Do, Io = o.search(q[:1], 100, impl="c")
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_baseline_config0_flat_10k(gpu_pkg):
    """BASELINE configs[0], index half: IndexFlatIP k=10 over 10k x 1024 (gaussian unit vectors, the
    reference's CPU-runnable case) against the oracle's sgemm + top-k.  ids exact where the fp64
    k/k+1 margin exceeds the fp32 rounding bound, scores within 1e-5."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10000, 1024)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((64, 1024)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ix = gpu_pkg.IndexFlatIP(1024)
    ix.add(x)
    D, I = ix.search(q, 10)
    o = oivf.FlatIP(1024)
    o.add(x)
    Do, Io = o.search(q, 10)
    s64 = np.sort(o.scores_f64(q), axis=1)[:, ::-1]
    margin = (s64[:, :10] - s64[:, 1:11]).min(axis=1)
    safe = margin > 1e-5
    assert safe.sum() >= 50
    assert np.array_equal(I[safe], Io[safe])
    assert np.abs(D - Do).max() < 1e-5


def test_large_query_batch_splits_internally(gpu_pkg):
    """2,500 queries in one call: more than one plan/scan launch (1,024 queries each) and more than one
    host staging chunk must give the same answer as the oracle, for host and device entry points."""
    P = gpu_pkg
    t = _torch()
    d, nlist, n, nq, k, nprobe = 64, 32, 6000, 2500, 5, 3
    x = osynth.corpus(21, 0, n, d, nlist)
    q = osynth.queries(21, 0, nq, d, nlist, n)
    c = osynth.centroids(21, nlist, d)
    ix = P.IndexIVFFlat(d, nlist)
    ix.set_centroids(c)
    ix.add(x)
    ix.nprobe = nprobe
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(c)
    o.add(x)
    Do, Io = o.search(q, k, nprobe=nprobe, impl="c")
    D, I = ix.search(q, k)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    Dd, Id = ix.search(t.from_numpy(q).cuda(), k)
    assert np.array_equal(Id.cpu().numpy(), Io) and np.array_equal(Dd.cpu().numpy(), Do)
    st = ix.last_stats()
    assert st["vectors"] == o.last_nscanned


def test_small_index_fuzz_against_the_oracle(gpu_pkg):
    """hypothesis: small indexes built from tiny integer components (exact arithmetic, exact score ties everywhere),
    custom ids, empty lists, several add() calls, k beyond the probed vectors, nprobe up to nlist, zero rows — through
    both fine-scan modes (fp32 ring / register scan; fp16 shortlist + re-score + fallback) and both dimensions' kernel
    sets (d = 1024: shared-memory ring; d = 64: register scan).  (D, I) must equal the C oracle bit for bit."""
    from hypothesis import given, settings, strategies as st

    P = gpu_pkg

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 2**31), d=st.sampled_from([64, 1024]), nlist=st.integers(1, 12), n=st.integers(0, 200),
           nq=st.integers(1, 6), k=st.integers(1, 40), nprobe=st.integers(1, 12), amp=st.integers(1, 2),
           two_stage=st.sampled_from([0, 32, 64]), pieces=st.integers(1, 3))
    def run(seed, d, nlist, n, nq, k, nprobe, amp, two_stage, pieces):
        rng = np.random.default_rng(seed)
        nprobe = min(nprobe, nlist)
        # sparse rows keep the number of distinct scores small: ties between and within lists
        def rows(m):
            v = np.zeros((m, d), dtype=np.float32)
            cols = rng.integers(0, d, (m, 3))
            vals = rng.integers(-amp, amp + 1, (m, 3)).astype(np.float32) / 4.0
            for j in range(3):
                v[np.arange(m), cols[:, j]] += vals[:, j]
            return v

        x, q, c = rows(n), rows(nq), rows(nlist)
        ids = rng.permutation(5 * n + 1)[:n].astype(np.int64)
        ix = P.IndexIVFFlat(d, nlist)
        if two_stage and d == 1024:
            ix.set_two_stage(two_stage)
        ix.set_centroids(c)
        o = oivf.IVFFlat(d, nlist)
        o.set_centroids(c)
        cuts = np.linspace(0, n, pieces + 1).astype(int)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b > a:
                ix.add_with_ids(x[a:b], ids[a:b])
                o.add(x[a:b], ids=ids[a:b])
        assert ix.ntotal == n and np.array_equal(ix.list_sizes(), o.list_sizes())
        ix.nprobe = nprobe
        D, I = ix.search(q, k)
        Do, Io = o.search(q, k, nprobe=nprobe, impl="c")
        assert np.array_equal(I, Io), (I, Io)
        assert np.array_equal(D, Do)

    run()
