"""Host-side logic of the "next" rows: parquet embedding shards and the tune sweep's
recall / Pareto arithmetic.  The index is played by the oracle (no GPU here)."""
import importlib
import json

import numpy as np

from oracle import ivf as oivf
from oracle import synth as osynth


def _pkg():
    return importlib.import_module("abstracts-search_b200")


class _OracleIndex:
    def __init__(self, d, nlist):
        self.ix = oivf.IVFFlat(d, nlist)
        self.d, self.nlist, self.nprobe = d, nlist, 1
        self.cp = _pkg().ClusteringParameters()

    def train(self, x):
        self.ix.train(x)

    def add(self, x):
        self.ix.add(x)

    def search(self, q, k):
        return self.ix.search(q, k, nprobe=min(self.nprobe, self.nlist))


def test_parquet_shards_roundtrip_and_fill(tmp_path):
    P = _pkg()
    d, nlist, n = 64, 8, 1000
    x = osynth.corpus(3, 0, n, d, nlist)  # lattice values are exact in float16 too
    ids = [f"https://openalex.org/W{i}" for i in range(n)]
    paths = P.store.write_shards(str(tmp_path / "data"), ids, x, shard_size=400, row_group_size=128)
    assert len(paths) == 3 and P.store.count_rows(str(tmp_path / "data")) == n
    got_ids, got = [], []
    for i, e in P.store.iter_row_groups(str(tmp_path / "data"), d):
        assert e.dtype == np.float32 and e.shape[0] <= 128
        got_ids += i
        got.append(e)
    assert got_ids == ids and np.array_equal(np.concatenate(got), x)
    ix = _OracleIndex(d, nlist)
    ix.ix.set_centroids(osynth.centroids(3, nlist, d))
    assert P.store.fill_index(ix, str(tmp_path / "data"), ids_parquet=str(tmp_path / "ids.parquet")) == n
    assert ix.ix.ntotal == n and P.faiss_io.read_ids_parquet(str(tmp_path / "ids.parquet")) == ids
    ref = oivf.IVFFlat(d, nlist)
    ref.set_centroids(osynth.centroids(3, nlist, d))
    ref.add(x)
    assert all(np.array_equal(a, b) for a, b in zip(ix.ix.ids, ref.ids))
    ix2 = _OracleIndex(d, nlist)
    assert P.store.train_index(ix2, str(tmp_path / "data"), max_rows=300) >= 300 and ix2.ix.is_trained


def test_tune_sweep_recall_pareto_and_params(tmp_path):
    P = _pkg()
    d, nlist, n, nq, k = 64, 16, 4000, 40, 10
    ix = _OracleIndex(d, nlist)
    ix.ix.set_centroids(osynth.centroids(5, nlist, d))
    ix.add(osynth.corpus(5, 0, n, d, nlist))
    q = osynth.queries(5, 0, nq, d, nlist, n)
    pts = P.tune.sweep(ix, q, k, nprobes=[1, 2, 4, 16], repeats=1)
    rec = [p["recall"] for p in pts]
    assert rec == sorted(rec) and rec[-1] == 1.0 and 0 < rec[0] <= 1.0  # more probes never lose recall
    assert ix.nprobe == 1  # restored
    assert P.tune.recall_at_k(np.array([[1, 2, 3]]), np.array([[3, 4, 1]])) == 2 / 3
    front = P.tune.pareto([{"nprobe": 1, "recall": 0.5, "qps": 100.0}, {"nprobe": 2, "recall": 0.7, "qps": 120.0},
                           {"nprobe": 4, "recall": 0.9, "qps": 60.0}, {"nprobe": 8, "recall": 0.9, "qps": 50.0}])
    assert [p["nprobe"] for p in front] == [2, 4]
    choice = P.tune.tune(ix, q, k, min_recall=0.99, nprobes=[1, 2, 4, 16], params_path=str(tmp_path / "params.json"),
                         untuned_path=str(tmp_path / "untuned.json"))
    assert choice["recall"] >= 0.99 and ix.nprobe == choice["nprobe"]
    assert json.load(open(tmp_path / "params.json"))["nprobe"] == choice["nprobe"]
    assert len(json.load(open(tmp_path / "untuned.json"))["sweep"]) == 4


def test_train_index_reads_only_the_drawn_row_groups(tmp_path, monkeypatch):
    """train_index must not decode the whole store — (file, row group) pairs come from the
    parquet footers, and only the seeded draw is read (3 of 8 groups for 300 rows of 128-row groups)."""
    P = _pkg()
    d, nlist, n = 64, 8, 1000
    x = osynth.corpus(3, 0, n, d, nlist)
    P.store.write_shards(str(tmp_path / "data"), [str(i) for i in range(n)], x, shard_size=400, row_group_size=128)
    groups = P.store.list_row_groups(str(tmp_path / "data"))
    assert len(groups) == 10 and sum(g[2] for g in groups) == n  # 3 shards: 4 + 4 + 2 row groups
    reads = []
    real = P.store._read_group
    monkeypatch.setattr(P.store, "_read_group", lambda path, g, dd, cols: (reads.append((path, g, tuple(cols))), real(path, g, dd, cols))[1])
    ix = _OracleIndex(d, nlist)
    seen = {}
    ix.train = lambda xs: seen.setdefault("x", xs.copy())
    rows = P.store.train_index(ix, str(tmp_path / "data"), max_rows=300, seed=7)
    assert rows == seen["x"].shape[0] >= 300 and len(reads) == 3 and all(c == ("embedding",) for _, _, c in reads)
    # the sample is the concatenation of exactly those groups, in draw order
    want = np.concatenate([real(p, g, d, ("embedding",))[1] for p, g, _ in reads])
    assert np.array_equal(seen["x"], want)
    assert P.store.train_index(ix, str(tmp_path / "data"), max_rows=300, seed=7) == rows  # seeded: same draw


def test_store_fuzz_ragged_shards_and_row_groups(tmp_path):
    """hypothesis: stores with ragged shard and row-group sizes (1-row groups, a short last shard, one group holding
    everything): the footer listing, the streaming iterator, fill_index (+ ids.parquet) and the seeded train draw
    must agree with the arrays that were written."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    P = _pkg()
    case = [0]

    class _Sink:
        def __init__(self, d):
            self.d, self.nlist, self.rows, self.trained = d, 4, [], None
            self.cp = P.ClusteringParameters()

        def add(self, x):
            assert x.dtype == np.float32 and x.shape[1] == self.d
            self.rows.append(x.copy())

        def train(self, x):
            self.trained = x.copy()

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(n=st.integers(1, 400), d=st.sampled_from([4, 64]), shard=st.integers(1, 500), group=st.integers(1, 300),
           cap=st.integers(1, 500), seed=st.integers(0, 2**31))
    def run(n, d, shard, group, cap, seed):
        rng = np.random.default_rng(seed)
        x = (rng.integers(-127, 128, (n, d)) / 128.0).astype(np.float32)  # exact in float16
        ids = [f"W{int(v)}" for v in rng.integers(0, 1 << 40, n)]
        case[0] += 1
        root = str(tmp_path / f"s{case[0]}")
        paths = P.store.write_shards(root, ids, x, shard_size=shard, row_group_size=group)
        assert len(paths) == -(-n // shard) and P.store.count_rows(root) == n
        groups = P.store.list_row_groups(root)
        assert sum(g[2] for g in groups) == n and all(0 < g[2] <= min(group, shard) for g in groups)
        got_ids, got = [], []
        for i, e in P.store.iter_row_groups(root, d):
            got_ids += i
            got.append(e)
        assert got_ids == ids and np.array_equal(np.concatenate(got), x)
        sink = _Sink(d)
        assert P.store.fill_index(sink, root, ids_parquet=root + "/../ids%d.parquet" % case[0]) == n
        assert np.array_equal(np.concatenate(sink.rows), x) and len(sink.rows) == len(groups)
        assert P.faiss_io.read_ids_parquet(root + "/../ids%d.parquet" % case[0]) == ids
        rows = P.store.train_index(sink, root, max_rows=cap, seed=seed % 1000)
        assert rows == sink.trained.shape[0] and min(cap, n) <= rows <= n
        # the sample consists of whole row groups of the store: every sampled row is a stored row
        stored = {r.tobytes() for r in x}
        assert all(r.tobytes() in stored for r in sink.trained)

    run()


def test_pareto_and_recall_properties():
    """hypothesis: `pareto` keeps exactly the operating points no other point dominates (recall and qps both at least
    as good, one strictly better) — one representative per duplicate — and `recall_at_k` is the intersection measure
    with -1 padding ignored on both sides."""
    from hypothesis import given, settings, strategies as st

    P = _pkg()
    pt = st.tuples(st.integers(0, 20), st.integers(0, 20))

    @settings(max_examples=300, deadline=None, derandomize=True)
    @given(pts=st.lists(pt, min_size=1, max_size=25))
    def run_pareto(pts):
        points = [{"nprobe": i, "recall": r / 20.0, "qps": float(q)} for i, (r, q) in enumerate(pts)]
        keep = P.tune.pareto(points)
        kept = {(p["recall"], p["qps"]) for p in keep}
        assert len(kept) == len(keep)
        assert [p["recall"] for p in keep] == sorted(p["recall"] for p in keep)

        def dominated(a, b):  # b dominates a
            return b[0] >= a[0] and b[1] >= a[1] and b != a

        allp = {(p["recall"], p["qps"]) for p in points}
        want = {a for a in allp if not any(dominated(a, b) for b in allp)}
        assert kept == want

    run_pareto()

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 2**31), nq=st.integers(1, 6), k=st.integers(1, 8), universe=st.integers(8, 40))
    def run_recall(seed, nq, k, universe):
        rng = np.random.default_rng(seed)
        I = np.stack([rng.permutation(universe)[:k] for _ in range(nq)]).astype(np.int64)
        T = np.stack([rng.permutation(universe)[:k] for _ in range(nq)]).astype(np.int64)
        pad = rng.random((nq, k)) < 0.2
        I[pad] = -1
        T[rng.random((nq, k)) < 0.2] = -1
        hits = sum(len(set(a[a >= 0]) & set(b[b >= 0])) for a, b in zip(I, T))
        assert P.tune.recall_at_k(I, T) == hits / max(1, int((T >= 0).sum()))
        assert P.tune.recall_at_k(T, T) == (1.0 if (T >= 0).any() else 0.0)

    run_recall()
