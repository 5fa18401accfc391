"""Device side of the counter-based synthetic corpus (absb_synth_*_dev): fills CUDA tensors with
the same integer-lattice rows that the CPU oracle regenerates for parity checks.
The reference ships no data; bench.py builds its index from these."""
from __future__ import annotations

from ._lib import check, current_stream_ptr, lib, ptr

KIND_CORPUS, KIND_CENTROIDS, KIND_QUERIES = 0, 1, 2
KIND_UNIT = 4  # flag: rows divided by their L2 norm (real-valued fp32 unit vectors, not fp16-representable)


def fill(kind: int, seed: int, row0: int, n: int, d: int, nlist: int, corpus_rows: int = 0, out=None, device=0):
    import torch

    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=f"cuda:{device}")
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= n * d
    with torch.cuda.device(out.device):
        check(lib().absb_synth_fill_dev(kind, seed, row0, n, d, nlist, corpus_rows, ptr(out), current_stream_ptr()))
    return out


def corpus(seed, row0, n, d, nlist, out=None, device=0, unit=False):
    return fill(KIND_CORPUS | (KIND_UNIT if unit else 0), seed, row0, n, d, nlist, 0, out, device)


def centroids(seed, nlist, d, out=None, device=0):
    return fill(KIND_CENTROIDS, seed, 0, nlist, d, nlist, 0, out, device)


def queries(seed, q0, n, d, nlist, corpus_rows, out=None, device=0, unit=False):
    return fill(KIND_QUERIES | (KIND_UNIT if unit else 0), seed, q0, n, d, nlist, corpus_rows, out, device)


def corpus_rows(seed, rows, d, nlist, out=None, unit=False):
    """Corpus rows for an explicit CUDA int64 tensor of row numbers."""
    import torch

    assert rows.is_cuda and rows.dtype == torch.int64 and rows.is_contiguous()
    n = rows.numel()
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=rows.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= n * d
    with torch.cuda.device(rows.device):
        check(lib().absb_synth_fill_rows_dev(KIND_CORPUS | (KIND_UNIT if unit else 0), seed, ptr(rows), n, d, nlist, 0, ptr(out),
                                             current_stream_ptr()))
    return out[:n] if out.shape[0] != n else out


def cluster_of(seed, row0, n, nlist, out=None, device=0):
    import torch

    if out is None:
        out = torch.empty((n,), dtype=torch.int64, device=f"cuda:{device}")
    assert out.is_cuda and out.dtype == torch.int64
    with torch.cuda.device(out.device):
        check(lib().absb_synth_cluster_dev(seed, row0, n, nlist, ptr(out), current_stream_ptr()))
    return out
