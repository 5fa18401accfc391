#!/usr/bin/env python
"""bench.py — queries/sec of the abstracts-search hot path (encode + IVF search, k=10) on B200.

One "step" = one 512-query batch through the whole path:
    stella_en_1.5B_v5 encode of 512 x 32-token queries (data-parallel over ranks)
    -> all-gather of the embeddings (N > 1)
    -> IVF65536,Flat search, nprobe 32, k 10, over this rank's inverted lists
    -> ONE all-gather of the per-shard partial top-k + merge (N > 1).

Workload (config.workload): BASELINE.json's metric is quoted on 207M x 1024 sharded by inverted list
over 8 GPUs (configs[3]); that index is 848 GB and does not fit fewer than 5 GPUs, so every GPU
holds the shard it holds in that configuration — 207M / 8 = 25,875,000 rows (106 GB) — and the total
grows with N ("scaling": "weak"): N = 8 is exactly the 207M x 1024 index of the metric, N = 1 is a
25.9M x 1024 IVF65536 index, larger than the 10M single-GPU search config (configs[2]; run it with
--rows-per-gpu 10000000).  The per-GPU scan work per query (32 probes x ~395 vectors x 4,104 B) is
the same at every N; the encode work per GPU shrinks as 1/N.

    value     device-resident inputs, CUDA events on the launching stream, max over ranks
    e2e       the same step through the numpy (host-buffer) API a faiss/sentence-transformers user
              calls: pinned host token ids -> encode -> host embeddings -> search -> host (D, I)
    roofline  the dominant kernel of the step, timed live with CUDA events inside the timed region
    cpu_baseline  the oracle port (torch-CPU fp32 encoder + C/OpenMP IVF) on a bounded sample

`--impl reference` times the CPU path alone (rank 0) — the reference's own CPU stack
(sentence-transformers + faiss-cpu) is not installable offline, so the oracle port stands in
(cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec (encode+IVF search, k=10, nprobe=32, IVF65536 over 1024-d fp32)"
UNIT = "queries/s"
SEED = 1234


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="search", choices=["search", "encode", "build", "oa_jsonl"],
                    help="search = the headline metric (encode+IVF search); encode = BASELINE configs[1], bulk embedding; "
                         "build = BASELINE configs[4], Index.add / Index.train rates; oa_jsonl = the host-side OpenAlex "
                         "JSON-lines front end (SURVEY 8f row 4), MB/s next to the reference program itself")
    ap.add_argument("--oa-records", type=int, default=40000, help="oa_jsonl workload: synthetic records per step")
    ap.add_argument("--add-rows", type=int, default=1 << 20, help="build workload: rows per add() per GPU")
    ap.add_argument("--train-rows", type=int, default=1 << 21, help="build workload: k-means sample rows per GPU")
    ap.add_argument("--encode-batch", type=int, default=32)
    ap.add_argument("--seq-len", type=int, default=256)
    ap.add_argument("--rows-per-gpu", type=int, default=25_875_000)
    ap.add_argument("--nlist", type=int, default=65536)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=512, help="queries per step (whole job)")
    ap.add_argument("--query-tokens", type=int, default=32)
    ap.add_argument("--coarse-impl", type=int, default=1, help="0 fp32 FFMA GEMM, 1 tcgen05 split-bf16 GEMM")
    ap.add_argument("--scan-chunk", type=int, default=512, help="vectors per scan work item")
    ap.add_argument("--scan-order", type=int, default=1, help="1 list-major work queue (default), 0 query-major")
    ap.add_argument("--gemm-variant", type=int, default=0, help="tcgen05 GEMM tile shape (0 auto; see absb_gemm_set_variant)")
    ap.add_argument("--gemm-ksplit", type=int, default=0,
                    help="k-slices of the residual-add GEMMs: 0 / 1 = never split (default), n = forced (experiment)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: all-gathers of the query path as NVLink peer-memory stores fused into our kernels "
                         "(default) or as NCCL calls")
    ap.add_argument("--pipeline", type=int, default=0,
                    help="1: QueryPipeline — consecutive batches software-pipelined on two streams (encode of batch i+1 in flight "
                         "with the search of batch i); the timed region covers fill and drain.  0 (default): encode then search on "
                         "one stream, every kernel of the forward chained by programmatic dependent launch — measured faster "
                         "(7.94 vs 8.18 ms per step at the 8-GPU per-GPU load, profiles/r02_overlap_timeline.md)")
    ap.add_argument("--coresident", action="store_true",
                    help="pipeline with the SM-sharing shapes (one register-capped 8-warp scan CTA per SM beside GEMM CTAs with a "
                         "smaller operand ring).  Measured on B200 (profiles/r02_overlap_timeline.md): the kernels do run at the "
                         "same time, but each slows the other down by as much as the overlap saves, so it is off by default")
    ap.add_argument("--scan-impl", type=int, default=1, help="fine scan of the serial path: 1 shared-memory ring (cp.async.bulk), 0 registers")
    ap.add_argument("--ring", default="", help="ring geometry warps,depth,stage_vecs of --scan-impl 1 (default 4,3,1)")
    ap.add_argument("--two-stage", type=int, default=64,
                    help="two-stage fine scan: fp16 shadow codes (+50%% index memory) give a shortlist of this many "
                         "candidates (32/64/128), fp32 codes their exact scores; identical results (0 = single-pass scan)")
    ap.add_argument("--scan-ctas", type=int, default=-1, help="resident scan CTAs per SM (<= 0: occupancy query)")
    ap.add_argument("--no-compact", action="store_true", help="skip Index.compact() after the synthetic fill")
    ap.add_argument("--no-kernel-events", action="store_true", help="do not time individual kernels in the timed region")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--corpus", default="unit", choices=["unit", "lattice"],
                    help="synthetic index contents: unit = real-valued L2-normalised fp32 rows (what `index fill` stores; not "
                         "exactly representable in fp16, so the two-stage scan's error bound is genuinely exercised), "
                         "lattice = integer/128 components (every fp32 score exact in any summation order: scores as well as "
                         "ids are compared bit for bit with the oracle)")
    ap.add_argument("--parity-queries", type=int, default=-1,
                    help="queries of the untimed oracle parity leg (half taken from the step's GPU-encoded embeddings, half "
                         "synthetic perturbed corpus rows); -1 = auto (<= 32, bounded by the rows the oracle must regenerate), 0 = off")
    ap.add_argument("--skip-secondary", action="store_true",
                    help="skip the secondary metrics (bulk encode, Index.add / k-means rates, 10M-row search) measured before "
                         "the main index is built")
    ap.add_argument("--secondary-rows", type=int, default=10_000_000, help="rows per GPU of the BASELINE configs[2] search index")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": float(j["hbm_gbs"]), "tf_burst": float(j["bf16_tflops"]),
                "tf_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU path (oracle port) — used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------
class CpuPath:
    """torch-CPU fp32 stella encoder + C/OpenMP IVF over the probed lists of a bounded query sample."""

    def __init__(self, total_rows: int, nlist: int, nprobe: int, k: int, nq: int, tokens: int, d: int = 1024,
                 corpus: str = "unit"):
        import torch

        from oracle import encoder as oenc
        from oracle import ivf as oivf
        from oracle import synth as osynth

        P = importlib.import_module("abstracts-search_b200.encoder")
        self.cfg = P.STELLA_1_5B
        self.oenc, self.k, self.nprobe, self.nq = oenc, k, nprobe, nq
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        g = torch.Generator().manual_seed(0)
        t0 = time.time()
        self.sd = {}
        for name, shape in self.cfg.param_shapes().items():
            t = torch.empty(shape, dtype=torch.float32)
            if name.endswith("layernorm.weight") or name == "norm.weight":
                t.normal_(1.0, 0.05, generator=g)
            else:
                t.normal_(0.0, 0.02, generator=g)
            self.sd[name] = t.numpy()
        rng = np.random.default_rng(4321)
        self.ids = rng.integers(0, self.cfg.vocab_size, (nq, tokens)).astype(np.int64)
        self.mask = np.ones_like(self.ids)
        # index: the real centroids, and only the inverted lists this sample probes (what faiss-cpu
        # would touch), rebuilt from the counter-based generator
        emb = oenc.forward_plain(self.cfg, self.sd, self.ids, self.mask, normalize=True)
        cent = np.empty((nlist, d), dtype=np.float32)
        for l0 in range(0, nlist, 8192):
            cent[l0:l0 + 8192] = osynth.centroids(SEED, nlist, d, l0, min(8192, nlist - l0))
        self.o = oivf.IVFFlat(d, nlist)
        self.o.set_centroids(cent)
        _, Ic = self.o.coarse(emb, nprobe, impl="c")
        oracle_fill_lists(self.o, np.unique(Ic), total_rows, nlist, d, corpus)
        self.vectors_per_query = float(self.o.list_sizes()[Ic].sum()) / nq
        self.setup_s = time.time() - t0

    def step(self):
        emb = self.oenc.forward_plain(self.cfg, self.sd, self.ids, self.mask, normalize=True)
        _, Ic = self.o.coarse(emb, self.nprobe, impl="c")
        return self.o.search_preassigned(emb, self.k, Ic, impl="c")

    def describe(self):
        return (f"{self.nq} queries x {self.ids.shape[1]} tokens per step: torch-CPU fp32 Qwen2-1.5B forward + C/OpenMP "
                f"coarse over {self.o.nlist} centroids + scan of the probed lists "
                f"({self.vectors_per_query:.0f} vectors/query)")


def oracle_fill_lists(o, want_lists, total_rows: int, nlist: int, d: int, corpus: str) -> int:
    """Rebuild exactly the inverted lists `want_lists` of the synthetic index inside the oracle index `o`
    from oracle/synth.py alone (rows in ascending order = the insertion order of the bench's fill): the
    oracle never sees product memory.  Returns the number of rows regenerated."""
    from oracle import synth as osynth

    want = np.unique(np.asarray(want_lists, dtype=np.int64))
    want = want[want >= 0]
    rows_l, lists_l = [], []
    for r0 in range(0, total_rows, 1 << 22):
        rows = np.arange(r0, min(r0 + (1 << 22), total_rows), dtype=np.int64)
        c = osynth.cluster_of(SEED, rows, nlist)
        m = np.isin(c, want)
        rows_l.append(rows[m])
        lists_l.append(c[m].astype(np.int64))
    rows, lists = np.concatenate(rows_l), np.concatenate(lists_l)
    gen = osynth.corpus_rows_unit if corpus == "unit" else osynth.corpus_rows
    for r0 in range(0, len(rows), 32768):
        r, l = rows[r0:r0 + 32768], lists[r0:r0 + 32768]
        o.add(gen(SEED, r, d, nlist), ids=r, list_ids=l)
    o._as_csr()
    return int(len(rows))


def auto_cpu_sample(total_rows: int, nlist: int, nprobe: int) -> int:
    # 16 queries, fewer only if their probed lists would exceed ~1.7M vectors (7 GB, ~1 min to regenerate)
    per_q = max(1.0, total_rows / nlist * nprobe)
    return int(max(2, min(16, 1_700_000 // per_q)))


def time_cpu(cp: CpuPath, steps: int, warmup: int):
    for _ in range(warmup):
        cp.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        cp.step()
    dt = time.perf_counter() - t0
    return cp.nq * steps / dt, dt / steps


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    total_rows = args.rows_per_gpu * world
    nq = args.cpu_sample or auto_cpu_sample(total_rows, args.nlist, args.nprobe)
    cp = CpuPath(total_rows, args.nlist, args.nprobe, args.k, nq, args.query_tokens, corpus=args.corpus)
    qps, s_per_step = time_cpu(cp, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cp.cores, "kind": "port", "sample": cp.describe()},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world: int):
    total = args.rows_per_gpu * world
    return {
        "workload": (f"end-to-end encode+search (BASELINE configs[3]): stella_en_1.5B_v5 encode of {args.batch} x "
                     f"{args.query_tokens}-token queries + IVF{args.nlist},Flat search k={args.k} nprobe={args.nprobe} over "
                     f"{total} x 1024 fp32 rows sharded by inverted list over {world} GPU(s) "
                     f"({args.rows_per_gpu} rows = {args.rows_per_gpu * 4104 / 1e9:.0f} GB per GPU; 8 GPUs = the 207M index)"),
        "rows_total": total, "rows_per_gpu": args.rows_per_gpu, "nlist": args.nlist, "nprobe": args.nprobe, "k": args.k,
        "corpus": ("unit: real-valued L2-normalised fp32 rows from the counter-based generator (oracle/synth.py regenerates "
                   "them bit for bit)" if args.corpus == "unit" else "lattice: integer/128 components, fp32 scores exact"),
        "queries_per_step": args.batch, "query_tokens": args.query_tokens,
        "parallelism": (f"encode dp{world} + coarse quantiser dp{world} (each rank: its own queries); 1 all-gather of {{embeddings | "
                        f"coarse ids}}; index sharded by list x{world}; 1 exchange of partial top-k"),
        "l2": "inputs larger than L2: >= 20 GB of list codes and 3.1 GB of weights stream per step (L2 = 126 MB)",
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = os.path.join(tempfile.gettempdir(), f"absb_clocks_{os.getpid()}.csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def apply_scan_impl(ix, args):
    w, dpt, sv = (int(v) for v in (getattr(args, "ring", "") or "4,3,1").split(","))
    ix.set_scan_impl(getattr(args, "scan_impl", 1), w, dpt, sv)
    ix.set_tunables(scan_chunk=getattr(args, "scan_chunk", -1), scan_ctas_per_sm=max(0, getattr(args, "scan_ctas", -1)))


def build_shard(P, torch, args, rank: int, world: int, dev: int, rows_per_gpu: int | None = None):
    d, nlist = 1024, args.nlist
    total = (rows_per_gpu or args.rows_per_gpu) * world
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT, device=dev)
    ix.set_tunables(scan_chunk=args.scan_chunk, coarse_impl=args.coarse_impl, scan_ctas_per_sm=args.scan_ctas)
    ix.set_scan_order(bool(args.scan_order))
    apply_scan_impl(ix, args)
    if args.two_stage:
        ix.set_two_stage(args.two_stage)
    if world > 1:
        ix.set_shard(rank, world)
    ix.set_centroids(P.synth.centroids(SEED, nlist, d, device=dev))
    keep = 1 << 20  # rows materialised per add
    span = keep * world
    xbuf = torch.empty((int(keep * 1.05) + 4096, d), dtype=torch.float32, device=f"cuda:{dev}")
    t0 = time.time()
    for r0 in range(0, total, span):
        n = min(span, total - r0)
        lists = P.synth.cluster_of(SEED, r0, n, nlist, device=dev)
        if world > 1:
            sel = torch.nonzero(lists % world == rank).squeeze(1)
            rows = (sel + r0).contiguous()
            lists = lists[sel].contiguous()
        else:
            rows = torch.arange(r0, r0 + n, dtype=torch.int64, device=f"cuda:{dev}")
        x = P.synth.corpus_rows(SEED, rows, d, nlist, out=xbuf, unit=args.corpus == "unit")
        ix.add_core(x, rows, lists)
    torch.cuda.synchronize()
    del xbuf
    torch.cuda.empty_cache()
    if not args.no_compact:
        # 1M-row adds leave every list one page per add, scattered over the pool: make lists contiguous
        # (in place) so that the scan's work items are scan_chunk vectors long instead of one page
        ix.compact()
    return ix, time.time() - t0


# ------------------------------------------------------------------------------------------------
# oracle parity at the configuration the metric is quoted on (untimed; rank 0 checks, all ranks search)
# ------------------------------------------------------------------------------------------------
def auto_parity_queries(total_rows: int, nlist: int, nprobe: int) -> int:
    # bounded by the rows the oracle has to regenerate (~2.2M rows = 9 GB of fp32 at the 207M index)
    per_q = max(1.0, total_rows / nlist * nprobe)
    return int(max(4, min(32, (2_200_000 // per_q) // 2 * 2)))


def parity_sample(args, P, torch, dist, rank, world, ix, search_fn, emb_step, device):
    """The north star's "top-k ids bit-exact at matched nprobe" (reference call sites
    /root/reference/Makefile:31-32, README.md:16,28) checked INSIDE the bench run, on the index the
    number is measured on: a sample of queries — half of them embeddings the GPU encoder produced in
    the timed step, half synthetic perturbed corpus rows (they have true near neighbours) — goes through
    the product's whole search (coarse + fine scan + exchange + merge, every rank taking part) once per
    scan mode, and through the oracle (C restatement: coarse over all centroids, scan of the probed
    lists regenerated from oracle/synth.py, k-best).  ids must be identical; a query whose ids differ is
    accepted only as a near-tie (at every rank the two ids' fp64 scores agree within the fp32 rounding
    bound, 2e-6 — counted, not hidden); scores within 2e-5, and bit-identical on the lattice corpus for the
    lattice queries."""
    n_total = args.parity_queries if args.parity_queries > 0 else auto_parity_queries(args.rows_per_gpu * world, args.nlist, args.nprobe)
    n_enc = min(n_total // 2, emb_step.shape[0])
    n_syn = n_total - n_enc
    d, k, nprobe, nlist = 1024, args.k, args.nprobe, args.nlist
    total_rows = args.rows_per_gpu * world
    unit = args.corpus == "unit"
    q_syn = P.synth.queries(SEED, 1000, n_syn, d, nlist, total_rows, device=int(device.split(":")[1]), unit=unit)
    q_dev = torch.cat([emb_step[:n_enc], q_syn], 0).contiguous()
    modes = ([("two_stage_%d" % args.two_stage, args.two_stage)] if args.two_stage else []) + [("single_pass", 0)]
    got = {}
    for name, ts in modes:
        ix.set_two_stage(ts)  # shadow codes stay; 0 = single-pass fp32 scan over the same lists
        D, I = search_fn(q_dev)
        got[name] = (D.cpu().numpy(), I.cpu().numpy())
    ix.set_two_stage(args.two_stage)
    _, Ic_gpu = ix.coarse(q_dev, nprobe)
    Ic_gpu = Ic_gpu.cpu().numpy()
    if rank != 0:
        return None
    from oracle import ivf as oivf
    from oracle import synth as osynth

    t0 = time.time()
    q = q_dev.cpu().numpy()
    cent = np.empty((nlist, d), dtype=np.float32)
    for l0 in range(0, nlist, 8192):
        cent[l0:l0 + 8192] = osynth.centroids(SEED, nlist, d, l0, min(8192, nlist - l0))
    o = oivf.IVFFlat(d, nlist)
    o.set_centroids(cent)
    _, Ic = o.coarse(q, nprobe, impl="c")
    # coarse ties / near-ties: a query whose nprobe-th and (nprobe+1)-th centroid scores are closer than
    # the fp32 rounding bound may legitimately probe a different last list; it is compared on the
    # oracle's own probe set only if the GPU chose the same one
    same_probe = np.array([np.array_equal(np.sort(a), np.sort(b)) for a, b in zip(Ic, Ic_gpu)])
    coarse_order_equal = bool(np.array_equal(Ic, Ic_gpu))
    rows_regen = oracle_fill_lists(o, Ic, total_rows, nlist, d, args.corpus)
    Do, Io = o.search_preassigned(q, k, Ic, impl="c")
    # fp64 scores of each query's probed vectors: the referee for near-ties
    off, codes, ids = o._as_csr()
    TOL = 2e-6  # fp32 rounding of a 1024-term inner product of unit vectors (observed: 2.4e-7)
    s64 = []
    for i in range(len(q)):
        sc = np.concatenate([codes[off[l]:off[l + 1]].astype(np.float64) @ q[i].astype(np.float64) for l in Ic[i]])
        idv = np.concatenate([ids[off[l]:off[l + 1]] for l in Ic[i]])
        s64.append(dict(zip(idv.tolist(), sc.tolist())))
    out = {"queries": int(len(q)), "from_gpu_encoder": int(n_enc), "synthetic_perturbed_rows": int(n_syn),
           "rows_total": int(total_rows), "oracle_rows_regenerated": rows_regen, "nprobe": nprobe, "k": k,
           "corpus": args.corpus, "coarse_ids_equal_in_order": coarse_order_equal,
           "queries_with_other_probe_set": int((~same_probe).sum()), "near_tie_tolerance": TOL,
           "oracle": "oracle/ivf_oracle.c (orc_coarse + orc_ivf_scan) over lists regenerated by oracle/synth.py", "modes": {}}
    ids_ok, scores_ok, bit_ok = True, True, True
    for name, (Dg, Ig) in got.items():
        exact = np.array([np.array_equal(a, b) for a, b in zip(Ig, Io)])
        # a query whose ids differ is acceptable only as a near-tie: at every rank the fp64 score of the id
        # the GPU returned equals the fp64 score of the oracle's id within the fp32 rounding bound
        excused = np.zeros(len(q), dtype=bool)
        for i in np.nonzero(~exact & same_probe)[0]:
            try:
                excused[i] = all(abs(s64[i][int(a)] - s64[i][int(b)]) <= TOL for a, b in zip(Ig[i], Io[i]))
            except KeyError:  # an id outside the probed lists: never acceptable
                excused[i] = False
        ok = exact | excused | ~same_probe
        diff = float(np.max(np.abs(Dg[same_probe] - Do[same_probe]))) if same_probe.any() else 0.0
        bit = bool(np.array_equal(Dg[same_probe], Do[same_probe]))
        out["modes"][name] = {"ids_equal": bool(ok.all()), "queries_ids_identical": int((exact & same_probe).sum()),
                              "queries_near_tie_reordered": int(excused.sum()), "queries_compared": int(same_probe.sum()),
                              "scores_max_abs_diff": diff, "scores_bit_identical": bit}
        ids_ok &= bool(ok.all())
        scores_ok &= diff <= 2e-5
        bit_ok &= bit
    syn = np.arange(len(q)) >= n_enc
    if args.corpus == "lattice" and (same_probe & syn).any():
        # lattice rows x lattice queries: every fp32 partial sum is exact -> scores must match bit for bit
        m = same_probe & syn
        out["lattice_queries_scores_bit_identical"] = bool(all(np.array_equal(Dg[m], Do[m]) and np.array_equal(Ig[m], Io[m])
                                                               for Dg, Ig in got.values()))
        scores_ok &= out["lattice_queries_scores_bit_identical"]
    if len(got) == 2:
        (Da, Ia), (Db, Ib) = got.values()
        out["two_stage_equals_single_pass_bitwise"] = bool(np.array_equal(Ia, Ib) and np.array_equal(Da, Db))
        ids_ok &= out["two_stage_equals_single_pass_bitwise"]
    out.update({"ids_equal": ids_ok, "scores_equal": scores_ok, "scores_bit_identical": bit_ok, "seconds": time.time() - t0})
    return out


# ------------------------------------------------------------------------------------------------
# secondary metrics, measured in the same run before the main index takes the memory
# ------------------------------------------------------------------------------------------------
def secondary_metrics(args, P, torch, dist, enc, rank, world, dev):
    """The other rates the north star asks for at every N (embeddings/s, index build, the 10M-row
    single-GPU search config), each a few hundred ms of GPU time:
      encode_b32_s256   BASELINE configs[1]: bulk encode, b = 32 x 256-token abstracts per GPU per step (DP)
      add / kmeans      BASELINE configs[4]: Index.add (assign GEMM + append; rows spread over ranks and routed to
                        the list owners by ONE NCCL all-to-all at N > 1) and k-means (ONE all-reduce of the
                        centroid sums per iteration at N > 1); the resulting list sizes are checked
                        against the histogram oracle/synth.py predicts for the same rows
      search_10M        BASELINE configs[2]: IVF65536,Flat search over 10M rows per GPU, nprobe 32, k 10."""
    device = f"cuda:{dev}"
    d, nlist = 1024, args.nlist
    out = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(steps):
            fn(warm + i)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)) / steps

    # ---- bulk encode, b = 32 x 256 tokens (and 512, stella's max_seq_length) ----------------------
    g = torch.Generator().manual_seed(99 + rank)
    for S in (256, 512):
        B = 32
        ids = torch.randint(0, P.STELLA_1_5B.vocab_size, (4, B, S), generator=g, dtype=torch.int64).to(device)
        mask = torch.ones((B, S), dtype=torch.int32, device=device)
        ms = timed(lambda i: enc.encode_tokens(ids[i % 4], mask, normalize_embeddings=True), 5, 3)
        fl = enc.last_stats()["flops"]
        out[f"encode_b32_s{S}"] = {"embeddings_per_s": B * world / (ms * 1e-3), "ms_per_step": ms, "unit": "embeddings/s",
                                   "tflops_per_gpu": fl / (ms * 1e-3) / 1e12,
                                   "frac_of_sustained_bf16": fl / (ms * 1e-3) / 1e12 / peaks()["tf_sustained"],
                                   "parallelism": f"dp{world}, no collective"}
        del ids, mask
    # ---- index build: k-means iterations and Index.add ------------------------------------------
    n_add = 1 << 20
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT, device=dev)
    ix.set_tunables(coarse_impl=args.coarse_impl)
    sh = None
    if world > 1:
        ix.set_shard(rank, world)
        sh = P.ShardedIndexIVFFlat(ix)
    unit = args.corpus == "unit"
    xt = P.synth.corpus(SEED + 1, rank * n_add, n_add, d, nlist, device=dev, unit=unit)
    ix.cp.niter = 2
    ix.cp.max_points_per_centroid = 1 << 30
    barrier()
    t0 = time.perf_counter()
    if sh is not None:
        sh.train_distributed(xt)
    else:
        ix.train(xt)
    barrier()
    train_s = max_over_ranks(time.perf_counter() - t0)
    cent = ix.get_centroids()
    out["kmeans"] = {"rows_x_iters_per_s": n_add * world * 2 / train_s, "seconds": train_s, "rows_per_gpu": n_add, "niter": 2,
                     "spherical": bool(ix.cp.spherical), "centroids_finite": bool(np.isfinite(cent).all()),
                     "centroid_norm_min_max": [float(np.linalg.norm(cent, axis=1).min()), float(np.linalg.norm(cent, axis=1).max())],
                     "collective": "1 all-reduce of [nlist, d] sums + counts per iteration (NCCL)" if world > 1 else "none"}
    del xt, cent
    ix.set_centroids(P.synth.centroids(SEED, nlist, d, device=dev))  # the generating centres: known assignment
    steps_add, warm_add = 4, 1
    xb = [P.synth.corpus(SEED, (s_ * world + rank) * n_add, n_add, d, nlist, device=dev, unit=unit) for s_ in range(steps_add + warm_add)]

    def add_step(i):
        if sh is not None:
            sh.add_distributed(xb[i])
        else:
            ix.add(xb[i])

    # every add() is timed on its own: the first adds of a process also pay the driver's mapping of fresh device
    # memory for the list slabs (4.3 GB per 1M-row add), which varies from box to box; the median is reported
    torch.cuda.empty_cache()
    per_step = []
    for i in range(warm_add + steps_add):
        per_step.append(timed(lambda _i, i=i: add_step(i), 1, 0))
    ms = float(np.median(per_step[warm_add:]))
    sizes = torch.from_numpy(ix.list_sizes()).to(device)
    if world > 1:
        dist.all_reduce(sizes)
    hist_ok = None
    if rank == 0:
        from oracle import synth as osynth

        rows = np.arange(0, (steps_add + warm_add) * world * n_add, dtype=np.int64)
        want = np.bincount(osynth.cluster_of(SEED, rows, nlist), minlength=nlist)
        hist_ok = bool(np.array_equal(want, sizes.cpu().numpy()))
    out["add"] = {"rows_per_s": n_add * world / (ms * 1e-3), "ms_per_step": ms, "ms_per_step_each": [round(v, 1) for v in per_step],
                  "rows_per_step_per_gpu": n_add,
                  "assign_tflops_per_gpu": n_add * 6 * 2.0 * nlist * d / (ms * 1e-3) / 1e12,
                  "list_sizes_equal_generator_histogram": hist_ok,
                  "collective": "1 all-to-all of (vector, id, list) to the list owners (NCCL)" if world > 1 else "none"}
    del xb, ix, sh, sizes
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    # ---- BASELINE configs[2]: IVF65536,Flat search over 10M rows per GPU --------------------------
    ix, _ = build_shard(P, torch, args, rank, world, dev, rows_per_gpu=args.secondary_rows)
    ix.nprobe = args.nprobe
    sh = P.ShardedIndexIVFFlat(ix) if world > 1 else None
    if sh is not None:
        sh.nprobe = args.nprobe
        if args.exchange == "peer":
            sh.use_peer_exchange(max_results=args.batch * args.k, strict=False)
    total = args.secondary_rows * world
    q = P.synth.queries(SEED, 0, args.batch, d, nlist, total, device=dev, unit=unit)  # perturbed rows of THIS corpus
    fn = (lambda i: sh.search(q, args.k)) if sh is not None else (lambda i: ix.search(q, args.k))
    ms = timed(fn, 10, 3)
    ix.set_profile(2)
    timed(fn, 5, 0)
    prof = ix.get_profile()
    ix.set_profile(0)
    st = ix.last_stats()
    moved = st["vectors"] * 2048 + args.batch * args.two_stage * 4104 if args.two_stage else st["bytes"]
    scan_ms = prof["scan_ms"] / 5
    out["search_10M"] = {"qps": args.batch / (ms * 1e-3), "ms_per_step": ms, "rows_total": total, "unit": "queries/s",
                         "vectors_scanned_per_query_per_gpu": st["vectors"] / args.batch,
                         "scan_ms": scan_ms, "scan_gbs_bytes_moved": moved / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None,
                         "scan_frac_of_hbm": moved / (scan_ms * 1e-3) / 1e9 / peaks()["hbm_gbs"] if scan_ms > 0 else None,
                         "coarse_gemm_ms": prof["coarse_gemm_ms"] / 5, "other_ms": prof["other_ms"] / 5,
                         "two_stage_shortlist": args.two_stage, "search only (queries resident, no encoder)": True}
    del ix, sh, q
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    P = importlib.import_module("abstracts-search_b200")
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    dev = local_rank
    device = f"cuda:{dev}"
    pk = peaks()
    assert args.batch % world == 0
    nq, S, k = args.batch, args.query_tokens, args.k
    per = nq // world

    if args.gemm_variant:
        importlib.import_module("abstracts-search_b200.encoder").gemm_set_variant(args.gemm_variant)
    if args.gemm_ksplit:
        importlib.import_module("abstracts-search_b200.encoder").gemm_set_ksplit(args.gemm_ksplit)
    enc = P.Encoder(config=P.STELLA_1_5B, device=device, random_init_seed=0)
    secondary = None
    if not args.skip_secondary:
        secondary = secondary_metrics(args, P, torch, dist, enc, rank, world, dev)
    # The fp16 shadow codes of the two-stage scan take +50% index memory (159 GB of 180 GB for the
    # 25.9M-row shard).  If any rank cannot allocate them, every rank rebuilds its shard without them
    # and the step uses the single-pass scan: same results, config.two_stage_shortlist says which ran.
    ix, build_s, err = None, 0.0, None
    try:
        ix, build_s = build_shard(P, torch, args, rank, world, dev)
    except (MemoryError, RuntimeError) as e:
        if not args.two_stage:
            raise
        err = repr(e)
    if args.two_stage:
        failed = [err]
        if world > 1:
            failed = [None] * world
            dist.all_gather_object(failed, err)
        if any(f is not None for f in failed):
            if rank == 0:
                print(f"two-stage index build failed ({[f for f in failed if f][0]}); rebuilding single-pass", file=sys.stderr)
            ix = None
            import gc

            gc.collect()
            torch.cuda.empty_cache()
            args.two_stage = 0
            ix, build_s = build_shard(P, torch, args, rank, world, dev)
    ix.nprobe = args.nprobe
    sh = P.ShardedIndexIVFFlat(ix) if world > 1 else None
    exchange = "none (single GPU)"
    px_emb = None
    if sh is not None:
        sh.nprobe = args.nprobe
        exchange = "nccl all-gather x2"
        if args.exchange == "peer" and sh.use_peer_exchange(max_results=args.batch * args.k, strict=False):
            # slot = one rank's {embeddings | coarse ids} record of ShardedIndexIVFFlat.search_spread
            px_emb = P.PeerExchange.over_group(dev, (args.batch // world) * (1024 * 4 + 8 * args.nprobe), strict=False)
            exchange = ("NVLink peer-memory stores fused into the merge kernels (no NCCL on the query path)"
                        if px_emb is not None else "peer-memory top-k exchange + nccl embedding all-gather")

    # inputs: pinned host token ids (e2e) and their device copies (value)
    g = torch.Generator().manual_seed(4321)
    ids_h = torch.randint(0, P.STELLA_1_5B.vocab_size, (nq, S), generator=g, dtype=torch.int64).pin_memory()
    mask_h = torch.ones((nq, S), dtype=torch.int32).pin_memory()
    ids_np, mask_np = ids_h.numpy(), mask_h.numpy()
    ids_d, mask_d = ids_h.to(device), mask_h.to(device)
    lo, hi = rank * per, (rank + 1) * per
    emb_all = torch.empty((nq, 1024), dtype=torch.float32, device=device)

    last = {}

    def step_dev():
        e = enc.encode_tokens(ids_d[lo:hi], mask_d[lo:hi], normalize_embeddings=True)
        if world > 1:
            # coarse top-nprobe of this rank's queries only; ONE all-gather of {embeddings | coarse ids}
            out = sh.search_spread(e, k, px_queries=px_emb)
            last["emb"] = sh.last_queries
            return out
        last["emb"] = e
        return ix.search(e, k)

    # ---- two-stream software pipeline over consecutive batches (the product's QueryPipeline) -----------
    pipe = None
    if args.pipeline:
        pipe = P.QueryPipeline(enc, ix, k=k, nprobe=args.nprobe, batch=nq, tokens=S, sharded=sh, px_emb=px_emb,
                               coresident=args.coresident)

        def run_pipelined(steps, host=False):
            """`steps` batches end to end; encode(i + 1) is in flight with search(i).  host=True: pinned host
            token ids in, numpy (D, I) out, the copies issued inside the pipeline."""
            out = None
            pipe.start()
            a, b = (ids_h[lo:hi], mask_h[lo:hi]) if host else (ids_d[lo:hi], mask_d[lo:hi])
            for _ in range(steps):
                t = pipe.submit(a, b)
                if t is not None:
                    out = pipe.result(t)
            for t in pipe.flush():
                out = pipe.result(t)
            pipe.join()
            return out

    def step_e2e():
        if pipe is not None:
            return run_pipelined(1, host=True)
        e = enc.encode_tokens(ids_np[lo:hi], mask_np[lo:hi], normalize_embeddings=True)  # host in, host out
        if world > 1:
            e_d = torch.from_numpy(e).to(device, non_blocking=True)
            D, I = sh.search_spread(e_d, k, px_queries=px_emb)
            return D.cpu().numpy(), I.cpu().numpy()
        return ix.search(e, k)  # numpy in, numpy out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up + correctness of the step's plumbing ------------------------------------------
    for _ in range(max(args.warmup, 3)):
        D, I = step_dev()
    torch.cuda.synchronize()
    st = ix.last_stats()
    es = enc.last_stats()
    Iw = I.cpu().numpy()
    assert (Iw >= 0).all() and Iw.shape == (nq, k), "search returned missing results on a full index"
    if world > 1:
        # the step's data-parallel coarse + gathered coarse ids against the replicated coarse on every rank
        Dr_, Ir_ = sh.search(last["emb"].clone(), k)
        assert np.array_equal(Ir_.cpu().numpy(), Iw) and np.array_equal(Dr_.cpu().numpy(), D.cpu().numpy()), \
            "search_spread differs from the replicated-coarse search"
    _, Ic = ix.coarse(last["emb"] if world > 1 else enc.encode_tokens(ids_d, mask_d, True), args.nprobe)
    distinct_lists = int(torch.unique(Ic).numel())
    # + coarse of the local slice (split + GEMM + select), record pack copies, all-gather (peer: push + wait), unpack copies, shard merge
    launches_per_step = int(es["launches"] + st["launches"] + ((3 + 2 + (2 if px_emb is not None else 1) + 2 + 1) if world > 1 else 0))

    # ---- timed region: value -------------------------------------------------------------------
    # Pass A (clean): K steps, CUDA events around the whole region on the launching stream -> value.
    # Pass B (instrumented): the same K steps again with the library recording a CUDA-event pair
    # around every kernel on that stream -> per-kernel durations for the roofline.  The ~230 extra
    # event records per step cost up to 8% at N=8 (short kernels), so they stay out of `value`.
    def timed(steps, pipelined=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if pipelined:
            Dp, Ip = run_pipelined(steps)  # both side streams are joined into the current one before ev1
        else:
            for _ in range(steps):
                step_dev()
        ev1.record()
        barrier()
        if pipelined:
            assert np.array_equal(Ip.cpu().numpy(), Iw), "pipelined result differs from the serial result"
        return max_over_ranks(ev0.elapsed_time(ev1))

    sampler = ClockSampler(dev) if rank == 0 else None
    if args.pipeline:
        timed(3, pipelined=True)  # warm the side streams
    ms_total = timed(args.steps, pipelined=bool(args.pipeline))
    kernel_events = not args.no_kernel_events
    enc_prof = ix_prof = None
    ms_instrumented = ms_serial = None
    if pipe is not None and pipe.coresident:
        pipe.disable_coresidency()  # the per-kernel numbers below are those of each kernel ALONE on the GPU (serial step)
        apply_scan_impl(ix, args)
    if args.pipeline:
        ms_serial = timed(args.steps) / args.steps  # the same K steps, encode then search on one stream
    if kernel_events:
        enc.set_profile(2)
        ix.set_profile(2)
        ms_instrumented = timed(args.steps) / args.steps
        enc_prof, ix_prof = enc.get_profile(), ix.get_profile()
        enc.set_profile(0)
        ix.set_profile(0)
    clocks = sampler.stop() if sampler else None
    st = ix.last_stats()
    ms_per_step = ms_total / args.steps
    value = nq / (ms_per_step * 1e-3)
    if pipe is not None and pipe.coresident:
        pipe.enable_coresidency()

    # ---- e2e: host buffers through the public API ------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        for _ in range(2):
            De, Ie = step_e2e()
        assert np.array_equal(Ie, Iw), "host-API result differs from the device-API result"
        barrier()
        t0 = time.perf_counter()
        if pipe is not None:
            De, Ie = run_pipelined(args.steps, host=True)  # K batches: pinned host ids in, numpy (D, I) out, every batch
            assert np.array_equal(Ie, Iw)
        else:
            for _ in range(args.steps):
                step_e2e()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
        if pipe is not None:
            h2d, d2h = per * S * (8 + 4), nq * k * 12
            api = ("QueryPipeline.submit(pinned host token ids) / .result() -> numpy (D, I): per batch H2D of the ids + mask on "
                   "the encode stream, absb_enc_forward_dev, all-gather, absb_ivf_search[_push]_dev, D2H of (D, I) on the search stream")
        else:
            h2d = per * S * (8 + 4) + (nq * 1024 * 4 if world == 1 else per * 1024 * 4)
            d2h = per * 1024 * 4 + nq * k * 12
            api = "Encoder.encode_tokens(numpy) -> IndexIVFFlat.search(numpy): absb_enc_forward + absb_ivf_search"
        e2e = {"value": nq / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "api": api}

    # ---- roofline of the dominant kernel -------------------------------------------------------
    rooflines = {}
    if kernel_events:
        scan_step_ms = ix_prof["scan_ms"] / args.steps
        hbm_src = pk["source"] + " copy bandwidth"
        scan_name = "ivf_scan_ring_kernel" if args.scan_impl >= 1 else "ivf_scan_kernel"
        if args.two_stage:
            # the fp16 shortlist pass alone: its algorithmic bytes are the fp16 shadow codes of every probed vector
            ms16 = ix_prof["scan16_ms"] / max(1, ix_prof["scan16_launches"])
            b16 = st["vectors"] * 2048
            g16 = b16 / (ms16 * 1e-3) / 1e9 if ms16 > 0 else 0.0
            rooflines[scan_name + "<fp16 shadow codes>"] = {
                "bound": "hbm", "achieved": g16, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": g16 / pk["hbm_gbs"], "traffic": None,
                "peak_source": hbm_src, "ms_per_launch": ms16, "ms_per_step": ms16 * ix_prof["scan16_launches"] / args.steps,
                "algorithmic_bytes_per_launch": b16, "vectors_per_step": st["vectors"],
                "note": "stage 1 of the two-stage scan: 2,048 B per probed vector; repeated probes of a list within a batch hit L2 "
                        "(list-major queue), so DRAM traffic is below the algorithmic bytes and frac can exceed 1"}
            # all three scan launches of a step together against the IndexIVFFlat algorithmic bytes (4,104 B per vector)
            moved = b16 + nq * args.two_stage * (1024 * 4 + 8)
            rooflines["two_stage_scan (fp16 pass + fp32 re-score + fallback)"] = {
                "bound": "hbm", "achieved": moved / (scan_step_ms * 1e-3) / 1e9 if scan_step_ms > 0 else 0.0, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": moved / (scan_step_ms * 1e-3) / 1e9 / pk["hbm_gbs"] if scan_step_ms > 0 else 0.0,
                "traffic": None, "peak_source": hbm_src, "ms_per_launch": scan_step_ms, "ms_per_step": scan_step_ms,
                "bytes_moved_per_step": moved, "ivfflat_algorithmic_bytes_per_step": st["bytes"],
                "ivfflat_equivalent_gbs": st["bytes"] / (scan_step_ms * 1e-3) / 1e9 if scan_step_ms > 0 else 0.0,
                "launches_per_step": ix_prof["scan_launches"] // max(1, args.steps)}
            # the single-pass fp32 scan (IndexIVFFlat's own bytes) over the same work items, measured in the same run
            ix.set_two_stage(0)
            q_sp = last["emb"]
            srch = (lambda: sh.search(q_sp, k)) if world > 1 else (lambda: ix.search(q_sp, k))
            for _ in range(2):
                srch()
            ix.set_profile(2)
            for _ in range(5):
                srch()
            sp_prof = ix.get_profile()
            ix.set_profile(0)
            st_sp = ix.last_stats()
            ix.set_two_stage(args.two_stage)
            ms_sp = sp_prof["scan_ms"] / max(1, sp_prof["scan_launches"])
            g_sp = st_sp["bytes"] / (ms_sp * 1e-3) / 1e9 if ms_sp > 0 else 0.0
            rooflines[scan_name + "<fp32 codes, single pass>"] = {
                "bound": "hbm", "achieved": g_sp, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": g_sp / pk["hbm_gbs"], "traffic": None,
                "peak_source": hbm_src, "ms_per_launch": ms_sp, "ms_per_step": 0.0, "algorithmic_bytes_per_launch": st_sp["bytes"],
                "vectors_per_launch": st_sp["vectors"],
                "note": "not part of the timed step (the step uses the two-stage scan): 5 single-pass searches of the step's own "
                        "queries after the timed region; algorithmic bytes = probed vectors x (4 x 1024 + 8) (SURVEY 8d)"}
        else:
            scan_ms = ix_prof["scan_ms"] / max(1, ix_prof["scan_launches"])
            scan_gbs = st["bytes"] / max(1, ix_prof["scan_launches"] // args.steps) / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
            rooflines[scan_name + "<fp32 codes, single pass>"] = {
                "bound": "hbm", "achieved": scan_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": scan_gbs / pk["hbm_gbs"],
                "traffic": None, "peak_source": hbm_src,
                "ms_per_launch": scan_ms, "ms_per_step": scan_step_ms,
                "algorithmic_bytes_per_step": st["bytes"], "vectors_per_step": st["vectors"]}
        gemm_tf = enc_prof["gemm_flops"] / (enc_prof["gemm_ms"] * 1e-3) / 1e12 if enc_prof["gemm_ms"] > 0 else 0.0
        n_gemm = 4 * P.STELLA_1_5B.num_layers + 1
        rooflines["gemm_bf16_tc_kernel"] = {
            "bound": "tensor", "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": gemm_tf / pk["tf_sustained"], "traffic": None,
            "peak_source": pk["source"] + " cuBLAS bf16, sustained (kernel timed inside a long step)",
            "ms_per_launch": enc_prof["gemm_ms"] / max(1, enc_prof["forwards"] * n_gemm),
            "ms_per_step": enc_prof["gemm_ms"] / args.steps,
            "algorithmic_flops_per_step": enc_prof["gemm_flops"] / args.steps}
        phases = {"encode_gemm_ms": enc_prof["gemm_ms"] / args.steps, "encode_attention_ms": enc_prof["attention_ms"] / args.steps,
                  "encode_other_ms": enc_prof["other_ms"] / args.steps, "coarse_gemm_ms": ix_prof["coarse_gemm_ms"] / args.steps,
                  "scan_ms": ix_prof["scan_ms"] / args.steps, "search_other_ms": ix_prof["other_ms"] / args.steps}
        dominant = max(rooflines, key=lambda n: rooflines[n]["ms_per_step"])
    else:
        phases, dominant = {}, None
    load_traffic(rooflines)

    # ---- oracle parity on a sample of this run's queries, both scan modes (untimed) -----------------
    fallbacks = ix.two_stage_fallbacks() if args.two_stage else 0  # of the timed + warm-up steps, before the parity leg
    if world > 1:
        t = torch.tensor([fallbacks], dtype=torch.int64, device=device)
        dist.all_reduce(t)
        fallbacks = int(t.item())
    parity = None
    if args.parity_queries != 0:
        step_dev()
        emb_step = last["emb"].clone()
        parity = parity_sample(args, P, torch, dist, rank, world, ix, (lambda qd: sh.search(qd, k)) if world > 1 else (lambda qd: ix.search(qd, k)),
                               emb_step, device)
        if world > 1:
            dist.barrier()

    if rank != 0:
        return
    # ---- CPU baseline (oracle port) on a bounded sample ----------------------------------------
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        n_cpu = args.cpu_sample or auto_cpu_sample(args.rows_per_gpu, args.nlist, args.nprobe)
        cp = CpuPath(args.rows_per_gpu, args.nlist, args.nprobe, k, n_cpu, S, corpus=args.corpus)
        qps, _ = time_cpu(cp, 2, 1)
        cpu = {"value": qps, "unit": UNIT, "cores": cp.cores, "kind": "port", "sample": cp.describe()}

    cfg = workload_config(args, world)
    cfg.update({"two_stage_shortlist": args.two_stage,
                "two_stage_fallback_queries_total": (fallbacks if args.two_stage else None),
                "two_stage_fallback_note": "summed over ranks, warm-up + timed + instrumented + e2e steps, %d queries each" % nq,
                "pipeline": (("QueryPipeline, two streams: encode of batch i+1 overlaps search of batch i; K batches timed end to end "
                              "including fill and drain; " + ("scan as one 8-warp 64 KB CTA per SM beside the GEMM CTAs (161 KB)"
                                                               if args.coresident else "full-size kernels on both streams"))
                             if args.pipeline else "none: encode then search, one stream"),
                "ms_per_step_serial_same_run": ms_serial,
                "exchange": exchange, "index_build_s": build_s, "distinct_probed_lists_per_step": distinct_lists,
                "scan_work_items_per_step": st["items"], "ms_per_step_with_kernel_events": ms_instrumented, "coarse_impl": "tcgen05 split-bf16 (6 bf16 products, fp32-faithful)"
                if args.coarse_impl == 1 else "fp32 FFMA", "kernel_events": "second timed pass of the same K steps" if kernel_events else "off"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (encoder GEMMs, fp32 accumulate) + f32 (IVF scores)", "data": "synthetic", "config": cfg,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
        "roofline": dict(rooflines[dominant], kernel=dominant) if dominant else None,
        "rooflines": rooflines, "phases_ms_per_step": phases, "cpu_baseline": cpu,
        "parity_sample": parity, "secondary": secondary,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# secondary workload: bulk encode (BASELINE configs[1]: b=32, 256-token abstracts) -> embeddings/s
# ------------------------------------------------------------------------------------------------
def cpu_encode_rate(B: int, S: int, steps: int = 2):
    import torch

    from oracle import encoder as oenc

    P = importlib.import_module("abstracts-search_b200.encoder")
    cfg = P.STELLA_1_5B
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    sd = {}
    for name, shape in cfg.param_shapes().items():
        t = torch.empty(shape, dtype=torch.float32)
        t.normal_(1.0 if name.endswith("norm.weight") else 0.0, 0.05 if name.endswith("norm.weight") else 0.02, generator=g)
        sd[name] = t.numpy()
    ids = np.random.default_rng(1).integers(0, cfg.vocab_size, (B, S)).astype(np.int64)
    mask = np.ones_like(ids)
    oenc.forward_plain(cfg, sd, ids, mask, normalize=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        oenc.forward_plain(cfg, sd, ids, mask, normalize=True)
    return B * steps / (time.perf_counter() - t0)


def run_encode(args, rank: int, world: int, local_rank: int):
    metric = f"embeddings/sec (stella_en_1.5B_v5 bulk encode, b={args.encode_batch}, {args.seq_len}-token abstracts)"
    if args.impl == "reference":
        if rank == 0:
            nb = 4
            v = cpu_encode_rate(nb, args.seq_len, max(1, args.steps))
            emit({"impl": "reference", "metric": metric, "value": v, "unit": "embeddings/s", "n_gpus": args.gpus,
                              "steps": args.steps, "warmup": 1, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
                              "data": "synthetic", "config": {"workload": "BASELINE configs[1] bulk encode"},
                              "cpu_baseline": {"value": v, "unit": "embeddings/s", "cores": os.cpu_count(), "kind": "port",
                                               "sample": f"{nb} x {args.seq_len}-token sequences per step, torch-CPU fp32"},
                              "e2e": {"value": v, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0})
        return
    import torch
    import torch.distributed as dist

    P = importlib.import_module("abstracts-search_b200")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    pk = peaks()
    B, S = args.encode_batch, args.seq_len
    if args.gemm_variant:
        importlib.import_module("abstracts-search_b200.encoder").gemm_set_variant(args.gemm_variant)
    if args.gemm_ksplit:
        importlib.import_module("abstracts-search_b200.encoder").gemm_set_ksplit(args.gemm_ksplit)
    enc = P.Encoder(config=P.STELLA_1_5B, device=device, random_init_seed=0)
    nbuf = 4
    g = torch.Generator().manual_seed(99 + rank)
    ids_h = torch.randint(0, P.STELLA_1_5B.vocab_size, (nbuf, B, S), generator=g, dtype=torch.int64).pin_memory()
    mask_h = torch.ones((B, S), dtype=torch.int32).pin_memory()
    ids_np, mask_np = ids_h.numpy(), mask_h.numpy()
    ids_d, mask_d = ids_h.to(device), mask_h.to(device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(steps):
            enc.encode_tokens(ids_d[i % nbuf], mask_d, normalize_embeddings=True)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    for i in range(max(3, args.warmup)):
        out = enc.encode_tokens(ids_d[i % nbuf], mask_d, normalize_embeddings=True)
    es = enc.last_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = timed(args.steps) / args.steps
    enc.set_profile(2)
    ms_instr = timed(args.steps) / args.steps
    prof = enc.get_profile()
    enc.set_profile(0)
    clocks = sampler.stop() if sampler else None
    # e2e: numpy in, numpy out
    ref = out.cpu().numpy()
    for i in range(2):
        e = enc.encode_tokens(ids_np[(max(3, args.warmup) - 1) % nbuf], mask_np, normalize_embeddings=True)
    assert np.array_equal(e, ref), "host-API embeddings differ from the device-API embeddings"
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        enc.encode_tokens(ids_np[i % nbuf], mask_np, normalize_embeddings=True)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    if rank != 0:
        return
    gemm_tf = prof["gemm_flops"] / (prof["gemm_ms"] * 1e-3) / 1e12
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        v = cpu_encode_rate(4, S, 2)
        cpu = {"value": v, "unit": "embeddings/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"4 x {S}-token sequences per step, 2 steps, torch-CPU fp32 oracle"}
    line = {
        "metric": metric, "value": B * world / (ms * 1e-3), "unit": "embeddings/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: stella_en_1.5B_v5 batch encode, b={B}, {S}-token synthetic abstracts, "
                               f"one batch per step per GPU (pure data parallel)", "batch": B, "seq_len": S,
                   "tokens_per_step_per_gpu": B * S, "flops_per_step_per_gpu": es["flops"],
                   "ms_per_step_with_kernel_events": ms_instr, "l2": "3.1 GB of bf16 weights stream per step (L2 = 126 MB)"},
        "clocks": clocks,
        "e2e": {"value": B * world / e2e_s, "unit": "embeddings/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": B * S * 12, "d2h_bytes_per_step": B * 1024 * 4, "api": "Encoder.encode_tokens(numpy)"},
        "gpu_launches": int(es["launches"]) * args.steps,
        "roofline": {"kernel": "gemm_bf16_tc_kernel", "bound": "tensor", "achieved": gemm_tf, "peak": pk["tf_sustained"],
                     "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sustained"], "traffic": None,
                     "ms_per_step": prof["gemm_ms"] / args.steps},
        "phases_ms_per_step": {"encode_gemm_ms": prof["gemm_ms"] / args.steps, "encode_attention_ms": prof["attention_ms"] / args.steps,
                               "encode_other_ms": prof["other_ms"] / args.steps},
        "cpu_baseline": cpu,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# secondary workload: index build (BASELINE configs[4]) -> rows/s of Index.add, k-means iteration rate
# ------------------------------------------------------------------------------------------------
def run_build(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    P = importlib.import_module("abstracts-search_b200")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    device = f"cuda:{dev}"
    pk = peaks()
    d, nlist, n_add = 1024, args.nlist, args.add_rows
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT, device=dev)
    ix.set_tunables(coarse_impl=args.coarse_impl)
    sh = None
    if world > 1:
        ix.set_shard(rank, world)
        sh = P.ShardedIndexIVFFlat(ix)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- Index.train: one Lloyd iteration rate on a per-GPU sample (niter = 2 keeps the run short) ----
    xt = P.synth.corpus(SEED + 1, rank * args.train_rows, args.train_rows, d, nlist, device=dev)
    ix.cp.niter = 2
    ix.cp.max_points_per_centroid = 1 << 30
    barrier()
    t0 = time.perf_counter()
    if sh is not None:
        sh.train_distributed(xt)
    else:
        ix.train(xt)
    barrier()
    train_s = max_over_ranks(time.perf_counter() - t0)
    del xt
    torch.cuda.empty_cache()
    # the bench index uses the generating centres as centroids (every list gets rows)
    ix.set_centroids(P.synth.centroids(SEED, nlist, d, device=dev))

    # ---- Index.add: assign (fused arg-max GEMM) + append; rows spread over ranks, all-to-all to owners ----
    total_steps = max(args.warmup, 3) + args.steps
    xbuf = [P.synth.corpus(SEED, (s * world + rank) * n_add, n_add, d, nlist, device=dev) for s in range(min(4, total_steps))]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def add_step(i):
        x = xbuf[i % len(xbuf)]
        if sh is not None:
            sh.add_distributed(x)
        else:
            ix.add(x)

    for i in range(max(args.warmup, 3)):
        add_step(i)
    sampler = ClockSampler(dev) if rank == 0 else None
    barrier()
    ev0.record()
    for i in range(args.steps):
        add_step(i)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    clocks = sampler.stop() if sampler else None
    # e2e: host rows through the numpy API (H2D of the rows inside the timed region)
    xh = torch.empty((n_add, d), dtype=torch.float32).pin_memory()
    xh.copy_(xbuf[0])
    xnp = xh.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        if sh is not None:
            sh.add_distributed(xnp)
        else:
            ix.add(xnp)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    ntotal = sh.ntotal if sh is not None else ix.ntotal
    if rank != 0:
        return
    rows = n_add * world
    flops_row = 6 * 2.0 * nlist * d  # split-bf16: six bf16 products per fp32-faithful inner product
    assign_tf = n_add * flops_row / (ms * 1e-3) / 1e12
    line = {
        "metric": f"rows/sec (Index.add: coarse assign over {nlist} centroids + list append, d=1024 fp32)",
        "value": rows / (ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 x3 split (fp32-faithful scores, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[4]: Index.add of {n_add} rows per GPU per step into IVF{nlist},Flat "
                               f"(rows spread over {world} GPU(s); assign locally, one all-to-all to the list owners)",
                   "rows_per_step": rows, "ntotal_after": int(ntotal),
                   "train": {"rows_per_gpu": args.train_rows, "niter": 2, "seconds": train_s,
                             "rows_x_iters_per_s": args.train_rows * world * 2 / train_s},
                   "l2": "each step reads 4 GB of fresh rows and 403 MB of split centroids per GPU (L2 = 126 MB)"},
        "clocks": clocks,
        "e2e": {"value": rows / e2e_s, "unit": "rows/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": n_add * d * 4,
                "d2h_bytes_per_step": 0, "api": "IndexIVFFlat.add(numpy) / ShardedIndexIVFFlat.add_distributed(numpy)"},
        "gpu_launches": None,
        "roofline": {"kernel": "gemm_bf16_tc_kernel (EPI_ARGMAX, 6 split-bf16 segments)", "bound": "tensor",
                     "achieved": assign_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": assign_tf / pk["tf_sustained"],
                     "traffic": None, "note": "whole add() step (assign + sort + scatter) charged to the GEMM"},
        "cpu_baseline": None,
    }
    emit(line)


def run_oa_jsonl(args):
    """Host-only secondary workload: OpenAlex works JSONL -> {"id","document"} JSONL.
    A step converts one resident block of synthetic records on all host threads; the baseline arm is
    the REFERENCE PROGRAM ITSELF (oracle/_ref/oa_jsonl, compiled from /root/reference/oa_jsonl.c) fed
    the same block through a pipe, as in the reference's Makefile:60-65 pipeline."""
    from oracle import oa_jsonl as O

    P = importlib.import_module("abstracts-search_b200")
    data = P.oa_jsonl.synth_records(SEED, args.oa_records)
    mb = len(data) / 1e6
    cores = os.cpu_count() or 1

    def ours():
        return P.oa_jsonl.convert(data, threads=0)

    def ref():
        return O.convert_reference(data)

    def rate(fn, steps, warmup):
        for _ in range(warmup):
            out = fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        return (time.perf_counter() - t0) / steps, out

    have_ref = O.reference_available()
    line = {"metric": "MB/s of OpenAlex works JSONL converted to {id, document} JSONL", "unit": "MB/s", "n_gpus": 0,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "oa_jsonl: %d synthetic OpenAlex works records (%.1f MB) per step, resident in host memory"
                                   % (args.oa_records, mb), "records": args.oa_records, "bytes": len(data)},
            "gpu_launches": 0}
    if args.impl == "reference":
        if not have_ref:
            emit({"impl": "reference", "unavailable": "oracle/_ref/oa_jsonl is not built (needs /root/reference)"})
            return
        dt, out = rate(ref, max(1, args.steps), max(1, args.warmup))
        line.update({"impl": "reference", "value": mb / dt, "ms_per_step": dt * 1e3,
                     "cpu_baseline": {"value": mb / dt, "unit": "MB/s", "cores": 1, "kind": "reference",
                                      "sample": "the whole block through a pipe into oracle/_ref/oa_jsonl"}})
    else:
        dt, out = rate(ours, args.steps, args.warmup)
        dt1, _ = rate(lambda: P.oa_jsonl.convert(data, threads=1), max(1, args.steps // 2), 1)
        line.update({"value": mb / dt, "ms_per_step": dt * 1e3, "threads": cores, "single_thread_mb_s": mb / dt1,
                     "records_kept": out.count(b"\n")})
        if have_ref:
            dtr, out_ref = rate(ref, 2, 1)
            line["cpu_baseline"] = {"value": mb / dtr, "unit": "MB/s", "cores": 1, "kind": "reference",
                                    "sample": "the whole block through a pipe into oracle/_ref/oa_jsonl (the reference program)"}
            line["matches_reference_bytes"] = bool(out_ref == out)
    emit(line)


def load_traffic(rooflines: dict):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return
    try:
        t = json.load(open(p))
    except Exception:
        return
    for name, r in rooflines.items():
        if name in t:
            r["traffic"] = t[name].get("dram_bytes_per_launch")
            r["traffic_source"] = t[name].get("source")


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries the JSON line and nothing else: native libraries (NCCL prints its version banner
    # on stdout at NCCL_DEBUG=VERSION/WARN/INFO) and stray prints are sent to stderr by pointing fd 1
    # at fd 2 for the whole run; emit() writes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "oa_jsonl":
        if rank == 0:
            run_oa_jsonl(args)
        return
    if args.impl == "reference":
        if args.workload == "encode":
            run_encode(args, rank, world, local_rank)
        else:
            run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        if args.workload == "encode":
            run_encode(args, rank, world, local_rank)
        elif args.workload == "build":
            run_build(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
