"""`sidecar-search index tune` (/root/reference/Makefile:27-32): sweep the
search-time operating point (nprobe) of a filled IVF index, measuring recall@k against exact search
and queries/sec, keep the Pareto-optimal points, and write `params.json` — the file app.py reads to
set `nprobe` (Makefile:12).

Ground truth is exact inner-product search over the same vectors.  It is computed from the IVF index
itself by probing every list (nprobe = nlist scans all inverted lists = IndexFlatIP over the same
rows), so no second copy of the corpus is needed; a separate IndexFlatIP can be passed instead.
"""
from __future__ import annotations

import json
import time

import numpy as np


def recall_at_k(I: np.ndarray, I_true: np.ndarray) -> float:
    """Mean fraction of the true top-k ids that the approximate search returned (faiss's
    intersection measure used by its autotune criterion)."""
    hits = 0
    for a, b in zip(np.asarray(I), np.asarray(I_true)):
        hits += len(np.intersect1d(a[a >= 0], b[b >= 0], assume_unique=False))
    return hits / max(1, int((np.asarray(I_true) >= 0).sum()))


def _to_numpy(a):
    return a.cpu().numpy() if hasattr(a, "is_cuda") else np.asarray(a)


def sweep(index, queries, k: int = 10, nprobes=None, ground_truth=None, repeats: int = 3) -> list[dict]:
    """One record per nprobe: {"nprobe", "recall", "qps", "ms_per_query"}."""
    nq = queries.shape[0]
    if nprobes is None:
        nprobes = [p for p in (1, 2, 4, 8, 16, 32, 64, 128, 256) if p <= min(index.nlist, 256)]
    saved = index.nprobe
    try:
        if ground_truth is None:
            if index.nlist <= 256:
                index.nprobe = index.nlist
                _, I_true = index.search(queries, k)
            else:
                raise ValueError("pass ground_truth=(exact ids [nq,k]) or an IndexFlatIP for indexes with nlist > 256")
        elif hasattr(ground_truth, "search"):
            _, I_true = ground_truth.search(queries, k)
        else:
            I_true = ground_truth
        I_true = _to_numpy(I_true)
        out = []
        for p in nprobes:
            index.nprobe = int(p)
            _, I = index.search(queries, k)  # warm-up + the result that is scored
            best = float("inf")
            for _ in range(repeats):
                t0 = time.perf_counter()
                _, I2 = index.search(queries, k)
                if hasattr(I2, "is_cuda"):
                    I2 = I2.cpu()
                best = min(best, time.perf_counter() - t0)
            out.append({"nprobe": int(p), "recall": recall_at_k(_to_numpy(I), I_true), "qps": nq / best,
                        "ms_per_query": best * 1e3 / nq})
        return out
    finally:
        index.nprobe = saved


def pareto(points: list[dict]) -> list[dict]:
    """Operating points not dominated in (recall up, qps up), by increasing recall."""
    keep = []
    for p in sorted(points, key=lambda r: (-r["recall"], -r["qps"])):
        if not keep or p["qps"] > max(q["qps"] for q in keep):
            keep.append(p)
    return sorted(keep, key=lambda r: r["recall"])


def tune(index, queries, k: int = 10, min_recall: float = 0.95, nprobes=None, ground_truth=None,
         params_path: str | None = None, untuned_path: str | None = None) -> dict:
    """Pick the fastest Pareto point with recall >= min_recall (else the highest-recall one), set it
    on the index and write params.json ({"nprobe": ..}) / untuned.json (the whole sweep)."""
    pts = sweep(index, queries, k, nprobes, ground_truth)
    front = pareto(pts)
    ok = [p for p in front if p["recall"] >= min_recall]
    choice = max(ok, key=lambda r: r["qps"]) if ok else front[-1]
    index.nprobe = choice["nprobe"]
    if untuned_path:
        json.dump({"k": k, "sweep": pts, "pareto": front}, open(untuned_path, "w"), indent=1)
    if params_path:
        json.dump({"nprobe": choice["nprobe"], "recall": choice["recall"], "qps": choice["qps"], "k": k},
                  open(params_path, "w"), indent=1)
    return choice
