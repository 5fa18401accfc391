#!/usr/bin/env python
"""Which kernels of the two QueryPipeline streams run at the same time?  Runs a few pipelined batches with the
library's per-kernel CUDA events on, dumps every span relative to one base event and prints, per batch, the
window of the fine scan and the encoder kernels that ran inside it.

    python tools/overlap_timeline.py [--query-tokens 4] [--no-coresident] [--batches 6]
"""
import argparse
import ctypes
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_875_000)
    ap.add_argument("--query-tokens", type=int, default=4)
    ap.add_argument("--batches", type=int, default=6)
    ap.add_argument("--no-coresident", action="store_true")
    ap.add_argument("--budget-kb", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench

    P = importlib.import_module("abstracts-search_b200")
    L = P.lib()
    bargs = argparse.Namespace(nlist=65536, rows_per_gpu=args.rows, scan_chunk=512, coarse_impl=1, scan_ctas=-1, scan_order=1,
                               two_stage=64, corpus="unit", no_compact=False)
    ix, _ = bench.build_shard(P, torch, bargs, 0, 1, 0)
    ix.nprobe = 32
    enc = P.Encoder(config=P.STELLA_1_5B, device="cuda:0", random_init_seed=0)
    nq, S = 512, args.query_tokens
    ids = torch.randint(0, P.STELLA_1_5B.vocab_size, (nq, S), dtype=torch.int64, device="cuda")
    mask = torch.ones((nq, S), dtype=torch.int32, device="cuda")
    pipe = P.QueryPipeline(enc, ix, k=10, nprobe=32, batch=nq, tokens=S, coresident=not args.no_coresident)
    if args.budget_kb:
        assert L.absb_gemm_set_smem_budget(args.budget_kb * 1024) == 0
    for _ in pipe.run((ids, mask) for _ in range(4)):
        pass
    pipe.join()
    torch.cuda.synchronize()
    enc.set_profile(2)
    ix.set_profile(2)
    base = torch.cuda.Event(enable_timing=True)
    end = torch.cuda.Event(enable_timing=True)
    base.record()
    for _ in pipe.run((ids, mask) for _ in range(args.batches)):
        pass
    pipe.join()
    end.record()
    torch.cuda.synchronize()
    total = base.elapsed_time(end)
    cap = 100000
    buf = np.zeros((cap, 3), dtype=np.float32)
    n = ctypes.c_int64()
    assert L.absb_enc_profile_spans(enc._h, ctypes.c_void_p(base.cuda_event), ctypes.c_void_p(buf.ctypes.data), cap, ctypes.byref(n)) == 0
    es = buf[: n.value].copy()
    assert L.absb_ivf_profile_spans(ix._h, ctypes.c_void_p(base.cuda_event), ctypes.c_void_p(buf.ctypes.data), cap, ctypes.byref(n)) == 0
    xs = buf[: n.value].copy()
    enc.set_profile(0)
    ix.set_profile(0)
    ename = {0: "gemm", 1: "attention", 2: "enc_other"}
    xname = {0: "scan", 1: "coarse_gemm", 2: "search_other"}
    print(f"{args.batches} batches in {total:.3f} ms = {total / args.batches:.3f} ms per batch (events on)")
    scans = [r for r in xs if int(r[0]) == 0 and r[2] - r[1] > 0.3]
    for r in scans:
        inside = [e for e in es if e[1] < r[2] and e[2] > r[1]]
        g = sum(min(e[2], r[2]) - max(e[1], r[1]) for e in inside if int(e[0]) == 0)
        a = sum(min(e[2], r[2]) - max(e[1], r[1]) for e in inside if int(e[0]) == 1)
        o = sum(min(e[2], r[2]) - max(e[1], r[1]) for e in inside if int(e[0]) == 2)
        print(f"  scan [{r[1]:8.3f}, {r[2]:8.3f}] {r[2]-r[1]:6.3f} ms: encoder spans overlapping it: {len(inside):3d} "
              f"(gemm {g:.3f} ms, attention {a:.3f} ms, other {o:.3f} ms of its window)")
    # encoder kernel durations: inside vs outside scan windows
    def in_scan(e):
        return any(e[1] < r[2] and e[2] > r[1] for r in scans)
    for kind in (0, 1, 2):
        a = [e[2] - e[1] for e in es if int(e[0]) == kind and in_scan(e)]
        b = [e[2] - e[1] for e in es if int(e[0]) == kind and not in_scan(e)]
        if a and b:
            print(f"  {ename[kind]:10s}: mean span {np.mean(a)*1e3:7.1f} us while a scan is running ({len(a)} spans) vs {np.mean(b)*1e3:7.1f} us otherwise ({len(b)})")
    if args.out:
        json.dump({"ms_per_batch": total / args.batches, "encoder_spans": es.tolist(), "index_spans": xs.tolist(),
                   "kinds": {"encoder": ename, "index": xname}}, open(args.out, "w"))


if __name__ == "__main__":
    main()
