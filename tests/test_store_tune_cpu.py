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
