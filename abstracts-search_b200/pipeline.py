"""The app.py query loop — `model.encode(query, prompt_name="s2p_query")` followed by `index.search(emb, k)`
(/root/reference/README.md:16,28) — for a STREAM of query batches, software-pipelined on two CUDA streams:

    stream E (high priority)   encode(i+1)  [H2D ids] -> 28 Qwen2 layers -> pool/Dense/normalise -> all-gather
    stream S                   search(i)    coarse GEMM -> plan -> fine scan -> merge -> exchange [-> D2H]

With full-size kernels on both streams the overlap is what the block scheduler finds at kernel boundaries
(the scan moves in as the GEMM CTAs of the other stream drain) plus the host<->device copies; measured gain
3-6% of a step.  `coresident=True` makes the two kernels SHARE SMs for the whole scan: the scan runs in its
co-resident shape (one 8-warp, 64 KB, <= 96-register CTA per SM: `IndexIVFFlat.set_scan_impl(2)`) and the
encoder's GEMM / attention CTAs leave room for it (`absb_gemm_set_smem_budget`: 4-5 operand stages instead of
5-7, 104 registers, max-shared carveout everywhere).  On B200 the kernels then do run concurrently
(profiles/r02_overlap_timeline.md) but the small-M GEMMs are HBM-latency-bound themselves and slow down 1.6x
under the scan's traffic while the 8-warp scan needs 4.2 ms instead of 1.8 ms: no net gain, hence off by
default.  Results are bit-identical either way — same arithmetic, only the schedule differs.

Inputs are token ids (the offline stand-in for tokenizer output): CUDA tensors for a device-resident
pipeline, or pinned host tensors — then every batch's ids are copied host->device on stream E and its
(D, I) device->host on stream S inside the pipeline, and `result()` hands back numpy arrays.
"""
from __future__ import annotations

from ._lib import check, lib

COEXIST_GEMM_SMEM = 161 * 1024  # bytes one GEMM CTA may take next to a 64 KB scan CTA (+ 2 x 1 KB reserved)


class QueryPipeline:
    def __init__(self, encoder, index, k: int = 10, nprobe: int | None = None, batch: int = 512, tokens: int = 32,
                 sharded=None, px_emb=None, coresident: bool = False, depth: int = 2):
        """`index`: this GPU's IndexIVFFlat; `sharded`: its ShardedIndexIVFFlat when the index is list-sharded over
        ranks (then `batch` is the WHOLE-JOB batch, this rank encodes batch / world of it and `px_emb` — a
        PeerExchange — or NCCL all-gathers the embeddings)."""
        import torch

        self.torch = torch
        self.enc, self.ix, self.sh, self.px_emb = encoder, index, sharded, px_emb
        self.k, self.batch, self.tokens, self.depth = int(k), int(batch), int(tokens), int(depth)
        if nprobe is not None:
            index.nprobe = int(nprobe)
            if sharded is not None:
                sharded.nprobe = int(nprobe)
        self.world = sharded.world if sharded is not None else 1
        self.rank = sharded.rank if sharded is not None else 0
        assert self.batch % self.world == 0
        self.per = self.batch // self.world
        self.device = encoder.device
        d = encoder.config.embed_dim
        with torch.cuda.device(self.device):
            self.s_enc = torch.cuda.Stream(priority=-1)  # tensor-bound encoder first in line for SM slots
            self.s_srch = torch.cuda.Stream(priority=0)
            self.emb = [torch.empty((self.batch, d), dtype=torch.float32, device=self.device) for _ in range(depth)]
            self.ids_d = [torch.empty((self.per, self.tokens), dtype=torch.int64, device=self.device) for _ in range(depth)]
            self.mask_d = [torch.empty((self.per, self.tokens), dtype=torch.int32, device=self.device) for _ in range(depth)]
            self.D_h = [torch.empty((self.batch, self.k), dtype=torch.float32).pin_memory() for _ in range(depth)]
            self.I_h = [torch.empty((self.batch, self.k), dtype=torch.int64).pin_memory() for _ in range(depth)]
            self.ev_enc = [torch.cuda.Event() for _ in range(depth)]
            self.ev_srch = [torch.cuda.Event() for _ in range(depth)]
        self.out = [None] * depth
        self.host_io = [False] * depth
        self.n_enc = 0   # batches whose encode has been issued
        self.n_srch = 0  # batches whose search has been issued
        self.coresident = bool(coresident)
        self._saved = None
        if self.coresident:
            self.enable_coresidency()

    # ---- SM sharing -----------------------------------------------------------------------------
    def enable_coresidency(self):
        check(lib().absb_gemm_set_smem_budget(COEXIST_GEMM_SMEM))
        self.ix.set_scan_impl(2)

    def disable_coresidency(self):
        check(lib().absb_gemm_set_smem_budget(0))
        self.ix.set_scan_impl(1, 4, 3, 1)
        self.ix.set_tunables(scan_ctas_per_sm=0)

    # ---- stages ---------------------------------------------------------------------------------
    def _issue_encode(self, ids, mask):
        torch = self.torch
        i = self.n_enc
        s = i % self.depth
        host = not ids.is_cuda
        self.host_io[s] = host
        with torch.cuda.stream(self.s_enc):
            self.s_enc.wait_event(self.ev_srch[s])  # the search that read this slot `depth` batches ago is done
            if host:
                self.ids_d[s].copy_(ids, non_blocking=True)
                self.mask_d[s].copy_(mask, non_blocking=True)
                ids, mask = self.ids_d[s], self.mask_d[s]
            e = self.enc.encode_tokens(ids, mask, normalize_embeddings=True)
            if self.world > 1:
                if self.px_emb is not None:
                    self.emb[s].copy_(self.px_emb.allgather(e).view(self.batch, -1))
                else:
                    import torch.distributed as dist

                    dist.all_gather_into_tensor(self.emb[s], e, group=self.sh.group)
            else:
                self.emb[s].copy_(e)
            self.ev_enc[s].record(self.s_enc)
        self.n_enc += 1

    def _issue_search(self):
        torch = self.torch
        j = self.n_srch
        s = j % self.depth
        with torch.cuda.stream(self.s_srch):
            self.s_srch.wait_event(self.ev_enc[s])
            D, I = self.sh.search(self.emb[s], self.k) if self.sh is not None else self.ix.search(self.emb[s], self.k)
            if self.host_io[s]:
                self.D_h[s].copy_(D, non_blocking=True)
                self.I_h[s].copy_(I, non_blocking=True)
            self.out[s] = (D, I)
            self.ev_srch[s].record(self.s_srch)
        self.n_srch += 1
        return j

    # ---- public ---------------------------------------------------------------------------------
    def submit(self, ids, mask):
        """Feed one batch of this rank's token ids [batch / world, tokens] (+ int32 mask).  Issues its encode and, if
        an earlier batch is waiting, that batch's search — so that encode(i+1) and search(i) are in flight
        together.  Returns the ticket of a batch whose search was issued, or None."""
        self._issue_encode(ids, mask)
        if self.n_enc - self.n_srch > 1:
            return self._issue_search()
        return None

    def flush(self):
        """Issue the searches of the batches still waiting; returns their tickets."""
        out = []
        while self.n_srch < self.n_enc:
            out.append(self._issue_search())
        return out

    def result(self, ticket: int):
        """(D, I) of a batch whose search has been issued: numpy arrays (after a host sync on that batch only)
        when the batch came from host buffers, CUDA tensors (valid until `depth` more batches have been
        submitted; ordered after stream S) otherwise."""
        assert ticket < self.n_srch and ticket >= self.n_srch - self.depth, "result already overwritten"
        s = ticket % self.depth
        if self.host_io[s]:
            self.ev_srch[s].synchronize()
            return self.D_h[s].numpy().copy(), self.I_h[s].numpy().copy()
        self.torch.cuda.current_stream().wait_event(self.ev_srch[s])
        return self.out[s]

    def join(self):
        """Make the current stream wait for everything issued so far."""
        cur = self.torch.cuda.current_stream()
        cur.wait_stream(self.s_enc)
        cur.wait_stream(self.s_srch)

    def start(self):
        """Order both side streams after the work already queued on the current stream."""
        cur = self.torch.cuda.current_stream()
        self.s_enc.wait_stream(cur)
        self.s_srch.wait_stream(cur)

    def run(self, batches):
        """Generator over (D, I) of every (ids, mask) batch, pipelined."""
        self.start()
        for ids, mask in batches:
            t = self.submit(ids, mask)
            if t is not None:
                yield self.result(t)
        for t in self.flush():
            yield self.result(t)
