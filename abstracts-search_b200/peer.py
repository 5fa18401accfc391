"""NVLink peer-memory exchange (csrc/peer.cuh): the all-gathers of the list-sharded query path as
remote stores + flags between the GPUs of one box, instead of NCCL calls.

One process per GPU.  `PeerExchange.over_group` is collective: every rank allocates its buffer,
the 64-byte CUDA IPC handles travel once over torch.distributed (any backend — it is plumbing),
and every rank maps every other rank's buffer.  After that the query path makes no collective
call at all: `ShardedIndexIVFFlat.search` runs `absb_ivf_search_push_dev` (the kernel that merges
this shard's partial top-k stores it into every rank's buffer) and `absb_peer_merge_shards_dev`
(waits for all ranks inside the kernel, then merges world x k candidates per query).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int, c_void_p

from ._lib import check, current_stream_ptr, lib, ptr


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, address: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 2}
        self._owner = owner


class PeerExchange:
    def __init__(self, device: int, rank: int, world: int, slot_bytes: int):
        self.device, self.rank, self.world = int(device), int(rank), int(world)
        self.slot_bytes = (int(slot_bytes) + 255) & ~255
        self._h = c_void_p()
        check(lib().absb_peer_create(self.device, self.rank, self.world, int(slot_bytes), byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().absb_peer_destroy(h)
            except Exception:
                pass

    # ---- wiring ---------------------------------------------------------------------------
    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(lib().absb_peer_ipc_handle(self._h, buf))
        return buf.raw

    def local_ptr(self) -> int:
        p = c_void_p()
        check(lib().absb_peer_local_ptr(self._h, byref(p)))
        return int(p.value)

    def connect(self, handles: list[bytes]) -> None:
        assert len(handles) == self.world and all(len(h) == 64 for h in handles)
        check(lib().absb_peer_connect(self._h, b"".join(handles)))

    @classmethod
    def over_group(cls, device: int, slot_bytes: int, group=None, strict: bool = True):
        """Collective over a torch.distributed group whose ranks sit on the GPUs of one box.
        strict=False: returns None on EVERY rank when any rank could not map its peers (no peer
        access between the GPUs, CUDA IPC unavailable) so that callers can keep the NCCL exchange."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        px, err = None, None
        try:
            px = cls(device, rank, world, slot_bytes)
            mine = px.handle()
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            mine, err = None, repr(e)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                px.connect(handles)
            except Exception as e:  # noqa: BLE001
                err = repr(e)
        elif err is None:
            err = "a peer could not create its buffer"
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)  # also: nobody pushes before everybody has mapped everybody
        bad = [f"rank {r}: {e}" for r, e in enumerate(errs) if e is not None]
        if bad:
            if strict:
                raise RuntimeError("NVLink peer exchange unavailable: " + "; ".join(bad))
            return None
        return px

    @classmethod
    def emulate(cls, device: int, world: int, slot_bytes: int) -> list["PeerExchange"]:
        """`world` ranks inside ONE process on one GPU (tests): buffers are wired by raw pointer."""
        pxs = [cls(device, r, world, slot_bytes) for r in range(world)]
        table = (c_void_p * world)(*[c_void_p(p.local_ptr()) for p in pxs])
        for p in pxs:
            check(lib().absb_peer_connect_ptrs(p._h, table))
        return pxs

    # ---- exchanges ------------------------------------------------------------------------
    def _view(self, address: int, nbytes_per_rank: int):
        import torch

        raw = torch.as_tensor(_RawCuda(address, self.world * self.slot_bytes, self), device=f"cuda:{self.device}")
        return raw.view(self.world, self.slot_bytes)[:, :nbytes_per_rank]

    def push(self, t) -> None:
        assert t.is_cuda and t.is_contiguous()
        check(lib().absb_peer_push_dev(self._h, ptr(t), t.numel() * t.element_size(), current_stream_ptr()))

    def wait(self, nbytes_per_rank: int):
        out = c_void_p()
        check(lib().absb_peer_wait_dev(self._h, byref(out), current_stream_ptr()))
        return self._view(int(out.value), nbytes_per_rank)

    def allgather(self, t):
        """[world, *t.shape] view of this epoch's ring entry (valid until the next-but-one
        exchange on this object); `t` is this rank's contribution, same shape on every rank."""
        assert t.is_cuda and t.is_contiguous()
        nbytes = t.numel() * t.element_size()
        out = c_void_p()
        check(lib().absb_peer_allgather_dev(self._h, ptr(t), nbytes, byref(out), current_stream_ptr()))
        g = self._view(int(out.value), nbytes)
        if nbytes == self.slot_bytes:
            return g.view(t.dtype).view((self.world,) + tuple(t.shape))
        return g.contiguous().view(t.dtype).view((self.world,) + tuple(t.shape))

    def status(self) -> int:
        s = c_int()
        check(lib().absb_peer_status(self._h, byref(s)))
        return s.value
