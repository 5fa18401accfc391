"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: list ownership, global id
numbering, the single packed all-gather and the merge order.  The per-rank scan is played by the
oracle here (there is no GPU); the product path uses the CUDA index and the device merge."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShard:
    """Oracle IVF that keeps only lists l with l % world == rank — the contract of
    absb_ivf_set_shard (default ids number ALL offered rows)."""

    def __init__(self, d, nlist, rank, world):
        from oracle import ivf as oivf

        self.ix = oivf.IVFFlat(d, nlist)
        self.d, self.nlist, self.rank, self.world = d, nlist, rank, world
        self.nprobe = 1
        self.rows_seen = 0

    @property
    def ntotal(self):
        return self.ix.ntotal

    def set_centroids(self, c):
        self.ix.set_centroids(c)

    def add(self, x):
        n = x.shape[0]
        ids = np.arange(self.rows_seen, self.rows_seen + n, dtype=np.int64)
        lists = self.ix.assign(x)
        lists = np.where(lists % self.world == self.rank, lists, -1)
        self.ix.add(x, ids=ids, list_ids=lists)
        self.rows_seen += n

    def add_core(self, x, ids, lists):
        lists = np.where(np.asarray(lists) % self.world == self.rank, lists, -1)
        self.ix.add(np.asarray(x), ids=np.asarray(ids), list_ids=lists)

    def search(self, x, k):
        _, Ic = self.ix.coarse(x, min(self.nprobe, self.nlist))
        return self.ix.search_preassigned(x, k, Ic)

    def coarse(self, x, nprobe):
        return self.ix.coarse(np.asarray(x), nprobe)

    def search_preassigned(self, x, k, Ic):
        return self.ix.search_preassigned(np.asarray(x), k, np.asarray(Ic))


class OracleOps:
    """CPU stand-ins (oracle arithmetic) for DeviceOps, so that the distributed build's host logic —
    slicing, the all-to-all routing, the all-reduce of sums, global ids — runs under gloo."""

    def __init__(self, shard):
        self.s = shard

    def to_device(self, a):
        import torch

        return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))

    def assign(self, x):
        import torch

        return torch.from_numpy(self.s.ix.assign(x.numpy()))

    def centroid_sums(self, x, assign):
        import torch

        k, d = self.s.nlist, self.s.d
        sums = np.zeros((k, d), dtype=np.float32)
        counts = np.zeros(k, dtype=np.float32)
        xn, an = x.numpy(), assign.numpy()
        for i in range(xn.shape[0]):
            sums[an[i]] += xn[i]
            counts[an[i]] += 1
        return torch.from_numpy(sums), torch.from_numpy(counts)

    def rand_perm(self, n, seed):
        from oracle import ivf as oivf

        return oivf.rand_perm(n, seed)

    def renorm(self, cent):
        import torch

        from oracle import ivf as oivf

        return torch.from_numpy(oivf.renorm_l2(cent.numpy()))

    def split_clusters(self, d, k, n, hassign, centroids):
        import ctypes

        from oracle import ivf as oivf

        return int(oivf.clib().orc_split_clusters(ctypes.c_int(d), ctypes.c_int64(k), ctypes.c_int64(n),
                                                  oivf._p(hassign, ctypes.c_float), oivf._p(centroids, ctypes.c_float)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("abstracts-search_b200")
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 64, 16, 3000, 11, 7, 5  # nq * k odd: the packed record needs its padding
    local = OracleShard(d, nlist, rank, world)
    local.set_centroids(osynth.centroids(7, nlist, d))
    sh = P.ShardedIndexIVFFlat(local, merge_fn=P.merge_partials_host)
    x = osynth.corpus(7, 0, n, d, nlist)
    sh.add(x[:1000])
    sh.add(x[1000:])
    sh.nprobe = nprobe
    total = sh.ntotal
    q = osynth.queries(7, 0, nq, d, nlist, n)
    D, I = sh.search(q, k)
    # the same batch SPREAD over the ranks: coarse where the slice lives, one all-gather of {queries | coarse ids}
    import torch

    q2 = osynth.queries(7, 100, 12, d, nlist, n)
    per = 12 // world
    Ds, Is = sh.search_spread(torch.from_numpy(np.ascontiguousarray(q2[rank * per:(rank + 1) * per])), k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), D=D, I=I, total=total, local=local.ntotal, Ds=Ds, Is=Is)
    dist.destroy_process_group()


def _cuts(world, n):
    return [0, 1800, n] if world == 2 else [0, 1800, 3100, n]


def _build_worker(rank, world, port, out_dir, spherical=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = importlib.import_module("abstracts-search_b200")
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 64, 16, 5000, 9, 6, 4
    x = osynth.corpus(11, 0, n, d, nlist)
    # uneven contiguous slices: rank 0 holds [0, 1800), the others share the rest
    cut = _cuts(world, n)
    mine = x[cut[rank]:cut[rank + 1]]
    local = OracleShard(d, nlist, rank, world)
    local.cp = P.ClusteringParameters()
    local.cp.max_points_per_centroid = 200  # 16 * 200 = 3200 < 5000: the subsampling branch
    local.cp.spherical = spherical
    local.device = "cpu"
    sh = P.ShardedIndexIVFFlat(local, merge_fn=P.merge_partials_host)
    ops = OracleOps(local)
    sh.train_distributed(mine, ops=ops)
    cent = local.ix.centroids.copy()
    # two distributed adds (global default ids continue across calls)
    h = (cut[rank + 1] - cut[rank]) // 2
    sh.add_distributed(mine[:h], ops=ops)
    sh.add_distributed(mine[h:], ops=ops)
    sh.nprobe = nprobe
    q = osynth.queries(11, 0, nq, d, nlist, n)
    D, I = sh.search(q, k)
    sizes = local.ix.list_sizes()
    np.savez(os.path.join(out_dir, f"b{rank}.npz"), D=D, I=I, cent=cent, sizes=sizes, total=sh.ntotal,
             ids0=local.ix.ids[rank], h=h)
    dist.destroy_process_group()


@pytest.mark.parametrize("spherical,world", [(False, 2), (True, 2), (True, 3)])
def test_distributed_train_and_add_world2_gloo(tmp_path, spherical, world):
    """train_distributed == single-process k-means (exact on the lattice corpus, where fp32 sums do not
    depend on order), with and without ClusteringParameters.spherical; add_distributed routes every row
    to its list owner with global ids."""
    mp.spawn(_build_worker, args=(world, _free_port(), str(tmp_path), spherical), nprocs=world, join=True)
    from oracle import ivf as oivf
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 64, 16, 5000, 9, 6, 4
    x = osynth.corpus(11, 0, n, d, nlist)
    cent = oivf.kmeans_train(x, nlist, max_points_per_centroid=200, spherical=spherical)
    if spherical:
        assert np.allclose(np.linalg.norm(cent, axis=1), 1.0, atol=1e-6)
    res = [np.load(tmp_path / f"b{r}.npz") for r in range(world)]
    for r in res:
        assert np.array_equal(r["cent"], cent), "distributed k-means differs from the single-process oracle"
        assert int(r["total"]) == n
    # the global batch order of the two add calls: [every rank's first half in rank order, then the second halves]
    cut = _cuts(world, n)
    h = [int(r["h"]) for r in res]
    order = np.concatenate([np.arange(cut[r], cut[r] + h[r]) for r in range(world)] +
                           [np.arange(cut[r] + h[r], cut[r + 1]) for r in range(world)])
    ref = oivf.IVFFlat(d, nlist)
    ref.set_centroids(cent)
    ref.add(x[order])  # default ids = position in the global batch order
    sizes = sum(r["sizes"] for r in res)
    assert np.array_equal(sizes, ref.list_sizes())
    for rank, r in enumerate(res):
        assert np.array_equal(r["ids0"], ref.ids[rank]), "list contents / order differ"
    D, I = ref.search(osynth.queries(11, 0, nq, d, nlist, n), k, nprobe=nprobe)
    for r in res:
        assert np.array_equal(r["I"], I) and np.array_equal(r["D"], D)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_world2_gloo(tmp_path, world):
    """world 3: 16 lists over 3 ranks (6 / 5 / 5), odd record sizes, a spread batch of 4 queries per rank."""
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import ivf as oivf
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 64, 16, 3000, 11, 7, 5
    ref = oivf.IVFFlat(d, nlist)
    ref.set_centroids(osynth.centroids(7, nlist, d))
    ref.add(osynth.corpus(7, 0, n, d, nlist))
    D, I = ref.search(osynth.queries(7, 0, nq, d, nlist, n), k, nprobe=nprobe)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for r in res:
        assert int(r["total"]) == n
        assert np.array_equal(r["I"], I) and np.array_equal(r["D"], D)
    assert sum(int(r["local"]) for r in res) == n and all(int(r["local"]) > 0 for r in res)
    D2, I2 = ref.search(osynth.queries(7, 100, 12, d, nlist, n), k, nprobe=nprobe)
    for r in res:
        assert np.array_equal(r["Is"], I2) and np.array_equal(r["Ds"], D2), "search_spread differs from the single index"
