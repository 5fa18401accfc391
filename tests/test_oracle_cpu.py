"""CPU tests of the oracle itself: numpy statement vs C statement, golden vectors, generator
properties.  None of these touch the product library."""
import numpy as np
import pytest

from conftest import golden
from oracle import ivf as oivf
from oracle import synth as osynth


def test_mt19937_matches_numpy_randomstate():
    import ctypes

    lib = oivf.clib()
    out = np.empty(1000, dtype=np.uint32)
    lib.orc_mt_raw(ctypes.c_uint32(1234), ctypes.c_int64(1000), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert np.array_equal(out, oivf.mt19937_raw(1234, 1000))
    # std::mt19937 known answer: the 10000th output of the default-seeded (5489) engine is 4123659995
    out = np.empty(10000, dtype=np.uint32)
    lib.orc_mt_raw(ctypes.c_uint32(5489), ctypes.c_int64(10000), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert int(out[-1]) == 4123659995


def test_rand_perm_c_vs_python():
    for n, seed in [(1, 5), (2, 5), (257, 1234), (5000, 1235)]:
        a = oivf.rand_perm(n, seed, use_c=True)
        b = oivf.rand_perm(n, seed, use_c=False)
        assert np.array_equal(a, b)
        assert np.array_equal(np.sort(a), np.arange(n))


def test_lattice_dot_products_are_exact_in_any_order():
    d, nlist = 1024, 64
    x = osynth.corpus(1234, 0, 256, d, nlist)
    q = osynth.queries(1234, 0, 8, d, nlist, 256)
    xi = np.rint(x * 128).astype(np.int64)
    qi = np.rint(q * 128).astype(np.int64)
    assert np.abs(xi).max() <= 127 and np.abs(qi).max() <= 127
    exact = (qi @ xi.T).astype(np.float64) / 16384.0
    s_blas = q @ x.T
    perm = np.random.default_rng(0).permutation(d)
    s_perm = q[:, perm] @ x[:, perm].T
    s_seq = np.zeros_like(s_blas)
    for j in range(d):  # strictly sequential fp32 accumulation
        s_seq += np.outer(q[:, j], x[:, j]).astype(np.float32)
    for s in (s_blas, s_perm, s_seq):
        assert s.dtype == np.float32 and np.array_equal(s.astype(np.float64), exact)


def test_synth_generator_is_counter_based():
    a = osynth.corpus(1234, 100, 50, 64, 16)
    b = osynth.corpus(1234, 0, 200, 64, 16)[100:150]
    assert np.array_equal(a, b)
    assert not np.array_equal(osynth.corpus(1235, 100, 50, 64, 16), a)
    c = osynth.cluster_of(1234, np.arange(100000), 16)
    cnt = np.bincount(c, minlength=16)
    assert cnt.min() > 5500 and cnt.max() < 7000
    # a corpus row sits next to its generating centroid
    cent = osynth.centroids(1234, 16, 64)
    x = osynth.corpus(1234, 0, 2000, 64, 16)
    assert (np.argmax(x @ cent.T, axis=1) == osynth.cluster_of(1234, np.arange(2000), 16)).mean() > 0.99


def test_rows_of_lists_inverts_cluster_of():
    rows = osynth.rows_of_lists(1234, 5000, 16, [3, 7])
    c = osynth.cluster_of(1234, np.arange(5000), 16)
    for l in (3, 7):
        assert np.array_equal(rows[l], np.nonzero(c == l)[0])


def test_oracle_matches_golden_lattice():
    g = golden("ivf_lattice_d1024.npz")
    d, nlist, n, nq, nprobe, k = (int(g[x]) for x in ("d", "nlist", "n", "nq", "nprobe", "k"))
    x = osynth.corpus(int(g["seed"]), 0, n, d, nlist)
    q = osynth.queries(int(g["seed"]), 0, nq, d, nlist, n)
    ix = oivf.IVFFlat(d, nlist)
    ix.set_centroids(osynth.centroids(int(g["seed"]), nlist, d))
    ix.add(x)
    assert np.array_equal(ix.list_sizes(), g["sizes"])
    assert np.array_equal(ix.assign(x), g["assign"])
    Dc, Ic = ix.coarse(q, nprobe)
    assert np.array_equal(Ic, g["Ic"]) and np.array_equal(Dc, g["Dc"])
    for impl in ("numpy", "c"):
        D, I = ix.search_preassigned(q, k, Ic, impl=impl)
        assert np.array_equal(I, g["I"]) and np.array_equal(D, g["D"])
    Dc2, Ic2 = ix.coarse(q, nprobe, impl="c")
    assert np.array_equal(Ic2, g["Ic"]) and np.array_equal(Dc2, g["Dc"])
    fl = oivf.FlatIP(d)
    fl.add(x)
    for impl in ("numpy", "c"):
        Df, If = fl.search(q, k, impl=impl)
        assert np.array_equal(If, g["If"]) and np.array_equal(Df, g["Df"])
    # an IVF search that probes every list equals the flat search
    ix.nprobe = nlist
    Da, Ia = ix.search(q, k)
    assert np.array_equal(Ia, g["If"]) and np.array_equal(Da, g["Df"])


def test_oracle_matches_golden_gauss_trained():
    g = golden("ivf_gauss_d64.npz")
    d, nlist, nprobe, k = (int(g[x]) for x in ("d", "nlist", "nprobe", "k"))
    ix = oivf.IVFFlat(d, nlist)
    ix.train(g["x"])
    assert np.allclose(ix.centroids, g["centroids"], rtol=0, atol=1e-6)
    ix.set_centroids(g["centroids"])
    ix.add(g["x"])
    safe = (g["coarse_margin"] > 1e-5) & (g["fine_margin"] > 1e-5)
    assert safe.sum() >= 30
    D, I = ix.search(g["q"], k, nprobe=nprobe)
    assert np.array_equal(I[safe], g["I"][safe])
    assert np.allclose(D[safe], g["D"][safe], atol=1e-6)
    Dc, Ic = ix.search(g["q"], k, nprobe=nprobe, impl="c")
    assert np.array_equal(Ic[safe], g["I"][safe])


def test_missing_results_are_padded_like_faiss():
    d, nlist = 64, 8
    ix = oivf.IVFFlat(d, nlist)
    ix.set_centroids(osynth.centroids(1, nlist, d))
    x = osynth.corpus(1, 0, 6, d, nlist)
    ix.add(x)
    q = osynth.queries(1, 0, 3, d, nlist, 6)
    for impl in ("numpy", "c"):
        D, I = ix.search(q, 10, nprobe=nlist, impl=impl)
        assert (I[:, 6:] == -1).all() and (D[:, 6:] == -oivf.FLT_MAX).all()
        assert (np.sort(I[:, :6], axis=1) == np.arange(6)).all()
        assert (np.diff(D[:, :6], axis=1) <= 0).all()
    # empty index, -1 coarse entries
    ix.reset()
    D, I = ix.search_preassigned(q, 4, np.full((3, 2), -1, dtype=np.int64))
    assert (I == -1).all()


def test_tie_rule_is_score_desc_then_id_asc():
    d, nlist = 64, 4
    ix = oivf.IVFFlat(d, nlist)
    ix.set_centroids(osynth.centroids(2, nlist, d))
    row = osynth.corpus(2, 0, 1, d, nlist)
    x = np.repeat(row, 7, axis=0)
    ids = np.array([50, 3, 99, 7, 21, 1, 64], dtype=np.int64)
    ix.add_with_ids(x, ids)
    for impl in ("numpy", "c"):
        D, I = ix.search(row, 5, nprobe=nlist, impl=impl)
        assert I[0].tolist() == [1, 3, 7, 21, 50] and len(set(D[0].tolist())) == 1


def test_kmeans_pieces_c_vs_numpy():
    import ctypes

    rng = np.random.default_rng(3)
    n, d, k = 3000, 32, 20
    x = rng.standard_normal((n, d)).astype(np.float32)
    assign = rng.integers(0, k - 2, n).astype(np.int64)  # the last two clusters stay empty
    lib = oivf.clib()
    cent = np.zeros((k, d), dtype=np.float32)
    hassign = np.zeros(k, dtype=np.float32)
    P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    lib.orc_compute_centroids(ctypes.c_int64(n), P(x, ctypes.c_float), ctypes.c_int(d), ctypes.c_int64(k),
                              P(assign, ctypes.c_int64), P(cent, ctypes.c_float), P(hassign, ctypes.c_float))
    c2, h2 = oivf.compute_centroids_numpy(x, assign, k)
    assert np.array_equal(cent, c2) and np.array_equal(hassign, h2)
    nsplit = lib.orc_split_clusters(ctypes.c_int(d), ctypes.c_int64(k), ctypes.c_int64(n), P(hassign, ctypes.c_float),
                                    P(cent, ctypes.c_float))
    assert nsplit == 2 and (hassign > 0).all() and abs(hassign.sum() - n) < 1e-3
    assert np.abs(cent[k - 1]).sum() > 0 and np.abs(cent[k - 2]).sum() > 0


def test_kmeans_train_recovers_planted_clusters():
    d, nlist, n = 64, 16, 4000
    x = osynth.corpus(99, 0, n, d, nlist)
    cent, hist = oivf.kmeans_train(x, nlist, niter=10, return_history=True)
    assert cent.shape == (nlist, d) and len(hist) == 10
    planted = osynth.cluster_of(99, np.arange(n), nlist)
    got = oivf.assign_argmax_ip(x, cent)
    # purity of the learned partition w.r.t. the planted one
    purity = sum(np.bincount(planted[got == c], minlength=nlist).max() for c in range(nlist) if (got == c).any()) / n
    assert purity > 0.6
    with pytest.raises(AssertionError):
        oivf.kmeans_train(x[:5], nlist)


def test_subsampling_uses_rand_perm():
    d, k = 8, 4
    x = np.random.default_rng(0).standard_normal((k * 300, d)).astype(np.float32)
    cent = oivf.kmeans_train(x, k, niter=0, max_points_per_centroid=256)
    perm = oivf.rand_perm(x.shape[0], 1234)
    sub = x[perm[: k * 256]]
    perm2 = oivf.rand_perm(k * 256, 1235)
    assert np.array_equal(cent, sub[perm2[:k]])


def test_numpy_and_c_restatements_agree_fuzz():
    """hypothesis: the two independent restatements of the faiss search semantics (oracle/ivf.py in numpy,
    oracle/ivf_oracle.c) on small lattice indexes — tiny integer components so that exact score ties are common,
    empty lists, k larger than the probed vectors (-1 / -FLT_MAX padding), nprobe up to nlist, custom ids — must
    return identical (D, I) for the flat search, the coarse quantiser and the IVF search."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 2**31), d=st.sampled_from([2, 4, 8]), nlist=st.integers(1, 9), n=st.integers(0, 60),
           nq=st.integers(1, 5), k=st.integers(1, 12), nprobe=st.integers(1, 9), amp=st.integers(1, 3))
    def run(seed, d, nlist, n, nq, k, nprobe, amp):
        rng = np.random.default_rng(seed)
        nprobe = min(nprobe, nlist)
        x = (rng.integers(-amp, amp + 1, (n, d)) / 4.0).astype(np.float32)
        q = (rng.integers(-amp, amp + 1, (nq, d)) / 4.0).astype(np.float32)
        cent = (rng.integers(-amp, amp + 1, (nlist, d)) / 4.0).astype(np.float32)
        ids = rng.permutation(4 * n + 1)[:n].astype(np.int64)
        flat = oivf.FlatIP(d)
        flat.add(x)
        Dn, In = flat.search(q, k, impl="numpy")
        Dc, Ic = flat.search(q, k, impl="c")
        assert np.array_equal(Dn, Dc) and np.array_equal(In, Ic)
        ix = oivf.IVFFlat(d, nlist)
        ix.set_centroids(cent)
        ix.add(x, ids=ids)
        assert int(ix.list_sizes().sum()) == n
        dn, cn = ix.coarse(q, nprobe, impl="numpy")
        dc, cc = ix.coarse(q, nprobe, impl="c")
        assert np.array_equal(cn, cc) and np.array_equal(dn, dc)
        Dn, In = ix.search(q, k, nprobe=nprobe, impl="numpy")
        Dc, Ic = ix.search(q, k, nprobe=nprobe, impl="c")
        assert np.array_equal(Dn, Dc) and np.array_equal(In, Ic)
        # the result is the top-k by (score desc, id asc) of the probed lists' vectors, padded with (-FLT_MAX, -1)
        s64 = q.astype(np.float64) @ x.T.astype(np.float64) if n else np.zeros((nq, 0))
        lists = ix.assign(x) if n else np.zeros(0, np.int64)
        for qi in range(nq):
            probed = np.isin(lists, cn[qi])
            cand = sorted((-s64[qi, j], int(ids[j])) for j in np.nonzero(probed)[0])[:k]
            want = [c[1] for c in cand] + [-1] * (k - len(cand))
            assert In[qi].tolist() == want
            assert all(Dn[qi, len(cand):] == np.float32(-3.4028234663852886e38))

    run()
