#!/usr/bin/env python
"""One index build, many scan configurations: A/B of the register scan (ivf_scan.cu / ivf_scan16.cu) against the
shared-memory ring scan (ivf_scan_ring.cu) over ring geometries, work-item lengths and resident CTAs per SM.
Search only (queries resident), CUDA events around K searches + the library's per-kernel events.  Every
configuration's (D, I) is compared bit for bit with the first one.  Prints one JSON line per configuration.

    python tools/scan_sweep.py --rows 25875000 --two-stage 64 --configs "0,0,0,0,128,0;1,4,4,2,128,0;..."
    config = impl,warps,depth,stage_vecs,scan_chunk,ctas_per_sm
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=25_875_000)
    ap.add_argument("--nlist", type=int, default=65536)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--two-stage", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--shard-world", type=int, default=1, help="emulate rank 0 of an N-way list-sharded index (rows = this rank's share)")
    ap.add_argument("--configs", default="0,0,0,0,128,0;1,4,4,2,128,0")
    ap.add_argument("--random-queries", action="store_true", help="unit gaussian queries (like encoder outputs) instead of perturbed rows")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench

    P = importlib.import_module("abstracts-search_b200")
    bargs = argparse.Namespace(nlist=args.nlist, rows_per_gpu=args.rows, scan_chunk=-1, coarse_impl=1, scan_ctas=-1, scan_order=1,
                               two_stage=args.two_stage, corpus="unit", no_compact=False)
    world = args.shard_world
    ix, build_s = bench.build_shard(P, torch, bargs, 0, world, 0)
    ix.nprobe = args.nprobe
    total = args.rows * world
    if args.random_queries:
        g = torch.Generator(device="cuda").manual_seed(1)
        q = torch.randn((args.batch, 1024), generator=g, device="cuda")
        q = (q / q.norm(dim=1, keepdim=True)).contiguous()
    else:
        q = P.synth.queries(bench.SEED, 0, args.batch, 1024, args.nlist, total, unit=True)
    ref = None
    pk = bench.peaks()
    for cfg in args.configs.split(";"):
        impl, warps, depth, sv, chunk, ctas = [int(v) for v in cfg.split(",")]
        ix.set_scan_impl(impl, warps, depth, sv)
        ix.set_tunables(scan_chunk=chunk, scan_ctas_per_sm=ctas if ctas > 0 else 0)
        for ts in ([args.two_stage, 0] if args.two_stage else [0]):
            ix.set_two_stage(ts)
            try:
                for _ in range(3):
                    D, I = ix.search(q, args.k)
            except P.AbsbError as e:
                print(json.dumps({"config": cfg, "two_stage": ts, "error": str(e)}), flush=True)
                continue
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(args.steps):
                D, I = ix.search(q, args.k)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / args.steps
            ix.set_profile(2)
            for _ in range(args.steps):
                ix.search(q, args.k)
            prof = ix.get_profile()
            ix.set_profile(0)
            st = ix.last_stats()
            moved = st["vectors"] * 2048 + args.batch * ts * 4104 if ts else st["bytes"]
            scan_ms = prof["scan_ms"] / args.steps
            res = (D.cpu().numpy(), I.cpu().numpy())
            if ref is None:
                ref = res
            same = bool(np.array_equal(ref[0], res[0]) and np.array_equal(ref[1], res[1]))
            print(json.dumps({"impl": impl, "warps": warps, "depth": depth, "stage_vecs": sv, "chunk": chunk, "ctas_per_sm": ctas,
                              "two_stage": ts, "search_ms": round(ms, 4), "scan_ms": round(scan_ms, 4),
                              "other_ms": round(prof["other_ms"] / args.steps, 4), "coarse_ms": round(prof["coarse_gemm_ms"] / args.steps, 4),
                              "scan_gbs": round(moved / (scan_ms * 1e-3) / 1e9, 1), "frac_hbm": round(moved / (scan_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                              "items": st["items"], "vectors": st["vectors"], "fallbacks": ix.two_stage_fallbacks() if ts else 0,
                              "identical_to_first": same}), flush=True)


if __name__ == "__main__":
    main()
