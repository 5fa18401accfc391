"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: list ownership, global id
numbering, the single packed all-gather and the merge order.  The per-rank scan is played by the
oracle here (there is no GPU); the product path uses the CUDA index and the device merge."""
import os
import socket
import sys

import numpy as np
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

    def search(self, x, k):
        _, Ic = self.ix.coarse(x, min(self.nprobe, self.nlist))
        return self.ix.search_preassigned(x, k, Ic)


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

    d, nlist, n, nq, k, nprobe = 64, 16, 3000, 12, 7, 5
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
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), D=D, I=I, total=total, local=local.ntotal)
    dist.destroy_process_group()


def test_sharded_search_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import ivf as oivf
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 64, 16, 3000, 12, 7, 5
    ref = oivf.IVFFlat(d, nlist)
    ref.set_centroids(osynth.centroids(7, nlist, d))
    ref.add(osynth.corpus(7, 0, n, d, nlist))
    D, I = ref.search(osynth.queries(7, 0, nq, d, nlist, n), k, nprobe=nprobe)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for r in res:
        assert int(r["total"]) == n
        assert np.array_equal(r["I"], I) and np.array_equal(r["D"], D)
    assert sum(int(r["local"]) for r in res) == n and all(int(r["local"]) > 0 for r in res)
