"""IVF index sharded by inverted list over the ranks of one torch.distributed group.

One process per GPU.  Centroids are replicated, so every rank computes the identical coarse
top-nprobe for the whole query batch; rank r owns lists l with l % world == r and scans only those;
ONE all-gather of the per-rank partial (D, I) [n, k] follows, and every rank merges the world x k
candidates per query with the same (score desc, id asc) order a single-shard search uses.  This
stands in for faiss's IndexShards, which merges on host threads.

The collective is the only exchange on the search path.  `add` needs none: every rank is offered
the same rows (or regenerates them) and keeps the ones whose list it owns; default ids number the
offered rows globally, so ids do not depend on the world size.
"""
from __future__ import annotations

import numpy as np


def owner_of_list(list_ids, world: int):
    """Rank that owns each inverted list."""
    return list_ids % world


def merge_partials_host(D_all: np.ndarray, I_all: np.ndarray, k: int):
    """Host statement of the shard merge ([world, n, k] -> [n, k]); used for CPU-side checks of the
    collective plumbing.  The product path merges on the device (absb_merge_shards_dev)."""
    world, n, _ = D_all.shape
    D = np.full((n, k), -3.4028234663852886e38, dtype=np.float32)
    I = np.full((n, k), -1, dtype=np.int64)
    for q in range(n):
        s = D_all[:, q, :].reshape(-1)
        ids = I_all[:, q, :].reshape(-1)
        keep = ids >= 0
        s, ids = s[keep], ids[keep]
        order = np.lexsort((ids, -s.astype(np.float64)))[:k]
        D[q, : len(order)] = s[order]
        I[q, : len(order)] = ids[order]
    return D, I


class ShardedIndexIVFFlat:
    """`local` is this rank's IndexIVFFlat (already set_shard(rank, world)); `group` a
    torch.distributed process group (None = default)."""

    def __init__(self, local, group=None, merge_fn=None):
        import torch.distributed as dist

        self.local = local
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d = local.d
        self.nlist = local.nlist
        self.nprobe = getattr(local, "nprobe", 1)
        self._merge_fn = merge_fn  # injected by the CPU (gloo) tests; None = device merge
        self._gD = self._gI = None

    @property
    def ntotal(self) -> int:
        import torch
        import torch.distributed as dist

        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([self.local.ntotal], dtype=torch.int64, device=dev)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def train(self, x):
        # Clustering is deterministic given (x, seed): every rank trains the same centroids.
        self.local.train(x)

    def add(self, x):
        self.local.add(x)

    def add_with_ids(self, x, ids):
        self.local.add_with_ids(x, ids)

    def add_core(self, x, ids, list_ids):
        self.local.add_core(x, ids, list_ids)

    def search(self, x, k: int):
        """x: the full query batch on every rank.  Returns the merged (D, I) on every rank.

        The per-rank record {I [n,k] i64, D [n,k] f32} is packed into one byte buffer so that the
        exchange is ONE all-gather (n*k*12 bytes per rank: 61,440 B for 512 x 10)."""
        import torch
        import torch.distributed as dist

        self.local.nprobe = self.nprobe
        D, I = self.local.search(x, k)
        if self.world == 1:
            return D, I
        as_numpy = not hasattr(D, "is_cuda")
        if as_numpy:
            D, I = torch.from_numpy(D), torch.from_numpy(I)
            if dist.get_backend(self.group) == "nccl":
                D, I = D.cuda(), I.cuda()
        n = D.shape[0]
        rec = n * k * 12
        if self._gD is None or self._gD.numel() != self.world * rec or self._gD.device != D.device:
            self._gD = torch.empty(self.world * rec, dtype=torch.uint8, device=D.device)
            self._gI = torch.empty(rec, dtype=torch.uint8, device=D.device)
        mine = self._gI
        mine[: n * k * 8].view(torch.int64).copy_(I.reshape(-1))
        mine[n * k * 8:].view(torch.float32).copy_(D.reshape(-1))
        dist.all_gather_into_tensor(self._gD, mine, group=self.group)
        if self._merge_fn is not None:
            g = self._gD.cpu().view(self.world, rec)
            I_all = g[:, : n * k * 8].contiguous().view(torch.int64).view(self.world, n, k).numpy()
            D_all = g[:, n * k * 8:].contiguous().view(torch.float32).view(self.world, n, k).numpy()
            return self._merge_fn(D_all, I_all, k)
        from ctypes import c_void_p

        from ._lib import check, current_stream_ptr, lib, ptr

        Dm = torch.empty((n, k), dtype=torch.float32, device=D.device)
        Im = torch.empty((n, k), dtype=torch.int64, device=D.device)
        base = self._gD.data_ptr()
        check(lib().absb_merge_shards_dev(D.device.index or 0, self.world, n, k, c_void_p(base + n * k * 8),
                                          c_void_p(base), rec, ptr(Dm), ptr(Im), current_stream_ptr()))
        if as_numpy:
            return Dm.cpu().numpy(), Im.cpu().numpy()
        return Dm, Im
