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


class DeviceOps:
    """The per-rank arithmetic of the distributed build, on this rank's GPU through the C ABI.
    (The gloo CPU tests substitute an oracle-backed object with the same five methods.)"""

    def __init__(self, local):
        self.local = local

    def to_device(self, a):
        import torch

        return a if hasattr(a, "is_cuda") else torch.from_numpy(np.ascontiguousarray(a)).cuda(self.local.device)

    def assign(self, x):
        return self.local.assign(x)

    def centroid_sums(self, x, assign):
        import torch

        from ._lib import check, current_stream_ptr, lib, ptr

        ix = self.local
        sums = torch.empty((ix.nlist, ix.d), dtype=torch.float32, device=x.device)
        counts = torch.empty((ix.nlist,), dtype=torch.float32, device=x.device)
        check(lib().absb_ivf_centroid_sums_dev(ix._h, x.shape[0], ptr(x), ptr(assign.contiguous()), ptr(sums), ptr(counts),
                                               current_stream_ptr()))
        return sums, counts

    def rand_perm(self, n: int, seed: int) -> np.ndarray:
        from ._lib import check, lib, ptr

        out = np.empty((n,), dtype=np.int32)
        check(lib().absb_rand_perm(n, seed, ptr(out)))
        return out.astype(np.int64)

    def renorm(self, cent):
        """fvec_renorm_L2 of the centroid rows, in place (ClusteringParameters.spherical)."""
        from ._lib import check, current_stream_ptr, lib, ptr

        check(lib().absb_renorm_rows_dev(cent.device.index or 0, cent.shape[0], cent.shape[1], ptr(cent), current_stream_ptr()))
        return cent

    def split_clusters(self, d: int, k: int, n: int, hassign: np.ndarray, centroids: np.ndarray) -> int:
        from ctypes import byref, c_int64

        from ._lib import check, lib, ptr

        ns = c_int64()
        check(lib().absb_kmeans_split_clusters(d, k, n, ptr(hassign), ptr(centroids), byref(ns)))
        return ns.value


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
        self._px = None  # NVLink peer exchange (peer.py); None = one NCCL all-gather

    def use_peer_exchange(self, px=None, max_results: int = 512 * 10, strict: bool = True) -> bool:
        """Route the exchange of the search path over NVLink peer memory instead of NCCL (collective
        when `px` is None: creates and connects a PeerExchange sized for n*k <= max_results).
        strict=False keeps the NCCL all-gather (returns False) where peer mapping is unavailable."""
        from .peer import PeerExchange

        if px is None:
            px = PeerExchange.over_group(self.local.device, max_results * 12 + 16, self.group, strict=strict)
            if px is None:
                return False
        assert px.world == self.world and px.rank == self.rank
        self._px = px
        return True

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

    # ---- distributed build (`add` = all-to-all of rows, `train` = all-reduce of sums) --
    def _ops(self, ops):
        return ops if ops is not None else DeviceOps(self.local)

    def _slice_offsets(self, n_local: int, device):
        """Rows are spread over ranks in rank order: rank r holds global rows [off[r], off[r+1])."""
        import torch
        import torch.distributed as dist

        sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
        sizes[self.rank] = n_local
        dist.all_reduce(sizes, group=self.group)
        off = torch.zeros(self.world + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(sizes.cpu(), 0)
        return off

    def add_distributed(self, x_local, ids_local=None, ops=None):
        """Index.add for rows that are SPREAD over the ranks (rank r holds the r-th contiguous slice
        of the global batch): every rank assigns only its own rows (1/W of the coarse GEMM), then one
        all-to-all routes each (vector, id, list) to the rank that owns the list.  The lists end up
        identical — contents and order — to a single index fed the concatenated batch, and default
        ids number the rows globally."""
        import torch
        import torch.distributed as dist

        ops = self._ops(ops)
        x = ops.to_device(x_local)
        n, W = x.shape[0], self.world
        dev = x.device
        off = self._slice_offsets(n, dev)
        base = getattr(self, "_rows_seen", 0)
        if ids_local is None:
            ids = torch.arange(base + int(off[self.rank]), base + int(off[self.rank]) + n, dtype=torch.int64, device=dev)
        else:
            ids = ops.to_device(np.ascontiguousarray(ids_local, dtype=np.int64) if not hasattr(ids_local, "is_cuda") else ids_local)
        self._rows_seen = base + int(off[-1])
        lists = ops.assign(x)
        lists = lists if hasattr(lists, "device") and not isinstance(lists, np.ndarray) else torch.from_numpy(lists)
        lists = lists.to(dev).to(torch.int64)
        owner = owner_of_list(lists, W)
        order = torch.argsort(owner, stable=True)
        send = torch.bincount(owner, minlength=W).to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        s_list, r_list = send.cpu().tolist(), recv.cpu().tolist()
        nr = int(sum(r_list))

        def exchange(t):
            out = torch.empty((nr,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_to_all_single(out, t[order].contiguous(), r_list, s_list, group=self.group)
            return out

        xr, idr, lr = exchange(x), exchange(ids), exchange(lists)
        self.local.add_core(xr if xr.is_cuda else xr.numpy(), idr if idr.is_cuda else idr.numpy(),
                            lr if lr.is_cuda else lr.numpy())
        return nr

    def train_distributed(self, x_local, ops=None):
        """Index.train (faiss Clustering defaults: subsample to 256 x nlist rows with rand_perm(seed),
        initial centroids = rows rand_perm(seed + 1)[:nlist] of the sample, niter x {assign, mean,
        split empty clusters}) with the training rows spread over the ranks.  Per iteration every
        rank assigns its slice and reduces it to per-list sums; ONE all-reduce of [nlist, d] sums +
        counts follows; mean and split_clusters are replicated (deterministic)."""
        import torch
        import torch.distributed as dist

        ops = self._ops(ops)
        cp = getattr(self.local, "cp", None)
        niter = cp.niter if cp else 10
        max_ppc = cp.max_points_per_centroid if cp else 256
        seed = cp.seed if cp else 1234
        spherical = bool(getattr(cp, "spherical", False))
        k, d = self.nlist, self.d
        x = ops.to_device(x_local)
        dev = x.device
        off = self._slice_offsets(x.shape[0], dev)
        lo, hi, n = int(off[self.rank]), int(off[self.rank + 1]), int(off[-1])
        if n < k:
            raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
        # global sample order -> the rows of it this rank holds, kept in sample order
        if n > k * max_ppc:
            perm = ops.rand_perm(n, seed)[: k * max_ppc]
        else:
            perm = np.arange(n, dtype=np.int64)
        ns = len(perm)
        mine = np.nonzero((perm >= lo) & (perm < hi))[0]  # sample positions held here
        xs = x[torch.from_numpy(perm[mine] - lo).to(dev)].contiguous()
        # initial centroids: sample rows perm2[:k]; every rank fills the rows it holds, one all-reduce
        if ns == k:
            pick = np.arange(k, dtype=np.int64)
        else:
            pick = ops.rand_perm(ns, seed + 1)[:k]
        pos_of = np.full(ns, -1, dtype=np.int64)
        pos_of[mine] = np.arange(len(mine))
        src = pos_of[pick]
        cent = torch.zeros((k, d), dtype=torch.float32, device=dev)
        have = np.nonzero(src >= 0)[0]
        if len(have):
            cent[torch.from_numpy(have).to(dev)] = xs[torch.from_numpy(src[have]).to(dev)]
        dist.all_reduce(cent, group=self.group)
        if ns == k:
            self.local.set_centroids(cent if cent.is_cuda else cent.numpy())
            return
        if spherical:
            cent = ops.renorm(cent)  # post_process_centroids before the first assignment
        self.local.set_centroids(cent if cent.is_cuda else cent.numpy())
        for _ in range(niter):
            assign = ops.assign(xs)
            sums, counts = ops.centroid_sums(xs, assign if hasattr(assign, "device") and not isinstance(assign, np.ndarray)
                                             else torch.from_numpy(assign).to(dev))
            dist.all_reduce(sums, group=self.group)
            dist.all_reduce(counts, group=self.group)
            inv = torch.where(counts > 0, 1.0 / counts, torch.zeros_like(counts))
            cent = sums * inv[:, None]
            hassign = counts.cpu().numpy().astype(np.float32)
            if (hassign == 0).any():
                c_h = np.ascontiguousarray(cent.cpu().numpy())
                ops.split_clusters(d, k, ns, hassign, c_h)
                cent = torch.from_numpy(c_h).to(dev)
            if spherical:
                cent = ops.renorm(cent.contiguous())
            self.local.set_centroids(cent if cent.is_cuda else cent.numpy())

    def _search_peer(self, x, k: int):
        """search() with the exchange fused into the kernels on both sides (csrc/peer.cuh): no
        NCCL call, no packed staging buffer, no separate copies."""
        import torch

        from ._lib import check, current_stream_ptr, lib, ptr

        px = self._px
        n = x.shape[0]
        assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.d
        if ((n * k * 8 + 15) & ~15) + n * k * 4 > px.slot_bytes:
            raise ValueError(f"{n} x {k} results exceed the exchange slot ({px.slot_bytes} bytes); call use_peer_exchange(max_results=...)")
        # a wait of an EARLIER search that gave up on a dead peer (20 s) left its mark in mapped host
        # memory: reading it costs no synchronisation, so every search checks it
        if px.status() != 0:
            raise RuntimeError("peer exchange timed out waiting for another rank; the merged results of that search "
                               "were returned as missing (-1)")
        with torch.cuda.device(x.device):
            st = current_stream_ptr()
            # single-pass and two-stage scans alike end in a merge kernel that stores this shard's top-k
            # into every rank's buffer (merge_partials_push): no pack / push kernel of its own
            check(lib().absb_ivf_search_push_dev(self.local._h, px._h, n, ptr(x), k, self.nprobe, st))
            Dm = torch.empty((n, k), dtype=torch.float32, device=x.device)
            Im = torch.empty((n, k), dtype=torch.int64, device=x.device)
            check(lib().absb_peer_merge_shards_dev(px._h, n, k, ptr(Dm), ptr(Im), st))
        return Dm, Im

    def search_spread(self, x_local, k: int, px_queries=None):
        """Index.search for a query batch that is SPREAD over the ranks (rank r holds the r-th contiguous slice,
        e.g. the queries it has just encoded; every rank the same number): each rank computes the coarse
        top-nprobe of ITS slice only (1 / W of the coarse GEMM and of the centroid select — centroids are
        replicated, so the result is the one every rank would compute), ONE all-gather moves the record
        {embeddings | coarse ids} of every rank (`px_queries`: a PeerExchange with slots of
        n_local * (4 d + 8 nprobe) bytes, else NCCL), then every rank scans its own lists for the whole batch and
        the usual exchange + merge of the partial top-k follows.  Same (D, I) as `search` on the gathered batch."""
        import torch
        import torch.distributed as dist

        self.local.nprobe = self.nprobe
        n_loc, d, npb, W = x_local.shape[0], self.d, int(min(self.nprobe, self.nlist)), self.world
        assert x_local.dtype == torch.float32 and x_local.is_contiguous() and x_local.shape[1] == d
        on_gpu = x_local.is_cuda
        Ic = self.local.coarse(x_local if on_gpu else x_local.numpy(), npb)[1]
        Ic = torch.as_tensor(Ic).to(x_local.device).to(torch.int64).contiguous()
        eb, cb = n_loc * d * 4, n_loc * npb * 8
        rec = torch.empty(eb + cb, dtype=torch.uint8, device=x_local.device)
        rec[:eb].view(torch.float32).copy_(x_local.reshape(-1))
        rec[eb:].view(torch.int64).copy_(Ic.reshape(-1))
        if px_queries is not None and W > 1:
            g = px_queries.allgather(rec)
        else:
            g = torch.empty((W, eb + cb), dtype=torch.uint8, device=x_local.device)
            if W > 1:
                dist.all_gather_into_tensor(g.view(-1), rec, group=self.group)
            else:
                g[0].copy_(rec)
        x_all = g[:, :eb].contiguous().view(torch.float32).view(W * n_loc, d)
        Ic_all = g[:, eb:].contiguous().view(torch.int64).view(W * n_loc, npb)
        self.last_queries = x_all  # the gathered batch (e.g. for a later exact re-check)
        if on_gpu:
            return self.search_preassigned(x_all, k, Ic_all)
        return self.search_preassigned(x_all.numpy(), k, Ic_all.numpy())

    def search_preassigned(self, x, k: int, Ic):
        """`search` with the coarse result given for the whole batch (IndexIVF::search_preassigned on every shard)."""
        import torch

        from ._lib import check, current_stream_ptr, lib, ptr

        n = x.shape[0]
        if self._px is not None and self.world > 1 and hasattr(x, "is_cuda") and x.is_cuda:
            px = self._px
            if ((n * k * 8 + 15) & ~15) + n * k * 4 > px.slot_bytes:
                raise ValueError(f"{n} x {k} results exceed the exchange slot ({px.slot_bytes} bytes)")
            if px.status() != 0:
                raise RuntimeError("peer exchange timed out waiting for another rank")
            with torch.cuda.device(x.device):
                st = current_stream_ptr()
                check(lib().absb_ivf_search_preassigned_push_dev(self.local._h, px._h, n, ptr(x), k, Ic.shape[1], ptr(Ic), st))
                Dm = torch.empty((n, k), dtype=torch.float32, device=x.device)
                Im = torch.empty((n, k), dtype=torch.int64, device=x.device)
                check(lib().absb_peer_merge_shards_dev(px._h, n, k, ptr(Dm), ptr(Im), st))
            return Dm, Im
        D, I = self.local.search_preassigned(x, k, Ic)
        return self._exchange_and_merge(D, I, k)

    def search(self, x, k: int):
        """x: the full query batch on every rank.  Returns the merged (D, I) on every rank.

        The per-rank record {I [n,k] i64, D [n,k] f32} is packed into one byte buffer so that the
        exchange is ONE all-gather (n*k*12 bytes per rank, segments padded to 16: 61,440 B for 512 x 10)."""
        import torch
        import torch.distributed as dist

        self.local.nprobe = self.nprobe
        if self._px is not None and self.world > 1 and hasattr(x, "is_cuda") and x.is_cuda:
            return self._search_peer(x, k)
        D, I = self.local.search(x, k)
        return self._exchange_and_merge(D, I, k)

    def _exchange_and_merge(self, D, I, k: int):
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return D, I
        as_numpy = not hasattr(D, "is_cuda")
        if as_numpy:
            D, I = torch.from_numpy(D), torch.from_numpy(I)
            if dist.get_backend(self.group) == "nccl":
                D, I = D.cuda(), I.cuda()
        n = D.shape[0]
        # record = {I [n,k] i64 | D [n,k] f32}, both segments padded to 16 bytes so that every rank's
        # block of the gathered buffer stays 8-byte aligned for the merge kernel's int64 loads (odd n*k)
        i_bytes = (n * k * 8 + 15) & ~15
        rec = i_bytes + ((n * k * 4 + 15) & ~15)
        if self._gD is None or self._gD.numel() != self.world * rec or self._gD.device != D.device:
            self._gD = torch.empty(self.world * rec, dtype=torch.uint8, device=D.device)
            self._gI = torch.zeros(rec, dtype=torch.uint8, device=D.device)
        mine = self._gI
        mine[: n * k * 8].view(torch.int64).copy_(I.reshape(-1))
        mine[i_bytes: i_bytes + n * k * 4].view(torch.float32).copy_(D.reshape(-1))
        dist.all_gather_into_tensor(self._gD, mine, group=self.group)
        if self._merge_fn is not None:
            g = self._gD.cpu().view(self.world, rec)
            I_all = g[:, : n * k * 8].contiguous().view(torch.int64).view(self.world, n, k).numpy()
            D_all = g[:, i_bytes: i_bytes + n * k * 4].contiguous().view(torch.float32).view(self.world, n, k).numpy()
            return self._merge_fn(D_all, I_all, k)
        from ctypes import c_void_p

        from ._lib import check, current_stream_ptr, lib, ptr

        Dm = torch.empty((n, k), dtype=torch.float32, device=D.device)
        Im = torch.empty((n, k), dtype=torch.int64, device=D.device)
        base = self._gD.data_ptr()
        check(lib().absb_merge_shards_dev(D.device.index or 0, self.world, n, k, c_void_p(base + i_bytes),
                                          c_void_p(base), rec, ptr(Dm), ptr(Im), current_stream_ptr()))
        if as_numpy:
            return Dm.cpu().numpy(), Im.cpu().numpy()
        return Dm, Im
