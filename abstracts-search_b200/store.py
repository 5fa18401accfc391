"""Embedding store formats either side of the hot path: the parquet shards that
`sidecar-search dump` writes (/root/reference/Makefile:46-49; `--shard-size 2097152
--row-group-size 65536`, README.md:60) and that `index train|fill` read back.

A shard is a parquet file with an `id` column (document id, string) and an `embedding` column
(fixed-size or variable list of float16/float32, length d).  Rows are streamed one row group at a
time (65,536 rows = 256 MB of fp32 at d = 1024) so that a 2,097,152-row shard never has to sit in
host memory, converted to float32 and handed to the index.
"""
from __future__ import annotations

import glob
import os

import numpy as np

SHARD_SIZE = 2_097_152
ROW_GROUP_SIZE = 65_536


def _embedding_matrix(col, d: int | None) -> np.ndarray:
    import pyarrow as pa

    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    flat = arr.flatten()
    x = flat.to_numpy(zero_copy_only=False)
    n = len(arr)
    width = len(flat) // max(n, 1)
    if n and (len(flat) != n * width or (d is not None and width != d)):
        raise RuntimeError(f"embedding column is not a [n, {d}] matrix (n={n}, values={len(flat)})")
    return np.ascontiguousarray(x.reshape(n, width if n else (d or 0)), dtype=np.float32)


def write_shards(dir_path: str, ids, embeddings: np.ndarray, shard_size: int = SHARD_SIZE,
                 row_group_size: int = ROW_GROUP_SIZE, dtype=np.float16) -> list[str]:
    """`sidecar-search dump` (sqlite -> parquet) for an in-memory batch: data-%05d.parquet shards."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    os.makedirs(dir_path, exist_ok=True)
    n, d = embeddings.shape
    ids = list(ids)
    assert len(ids) == n
    paths = []
    for s, r0 in enumerate(range(0, n, shard_size)):
        r1 = min(n, r0 + shard_size)
        emb = np.ascontiguousarray(embeddings[r0:r1], dtype=dtype)
        col = pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1)), d)
        path = os.path.join(dir_path, f"data-{s:05d}.parquet")
        pq.write_table(pa.table({"id": pa.array(ids[r0:r1]), "embedding": col}), path, row_group_size=row_group_size)
        paths.append(path)
    return paths


def iter_row_groups(dir_path: str, d: int | None = None, columns=("id", "embedding")):
    """Yield (ids list, embeddings float32 [n, d]) per parquet row group, shards in name order."""
    import pyarrow.parquet as pq

    files = sorted(glob.glob(os.path.join(dir_path, "*.parquet")))
    if not files:
        raise RuntimeError(f"no parquet shards under {dir_path}")
    for path in files:
        pf = pq.ParquetFile(path)
        for g in range(pf.num_row_groups):
            t = pf.read_row_group(g, columns=list(columns))
            ids = t.column("id").to_pylist() if "id" in t.column_names else None
            yield ids, _embedding_matrix(t.column("embedding"), d)


def count_rows(dir_path: str) -> int:
    import pyarrow.parquet as pq

    return sum(pq.ParquetFile(p).metadata.num_rows for p in sorted(glob.glob(os.path.join(dir_path, "*.parquet"))))


def fill_index(index, dir_path: str, ids_parquet: str | None = None) -> int:
    """`sidecar-search index fill DATA_DIR` (/root/reference/Makefile:24-25): stream every shard into
    index.add(); faiss ids are the running row numbers, and `ids.parquet` (row -> document id) is
    written next to the index when a path is given."""
    from .faiss_io import write_ids_parquet

    all_ids, n = [], 0
    for ids, x in iter_row_groups(dir_path, index.d):
        index.add(x)
        n += x.shape[0]
        if ids_parquet is not None:
            all_ids.extend(ids)
    if hasattr(index, "compact"):
        index.compact()  # row-group-sized add() calls leave every list scattered over the page pool
    if ids_parquet is not None:
        write_ids_parquet(ids_parquet, all_ids)
    return n


def train_index(index, dir_path: str, max_rows: int | None = None, seed: int = 1234) -> int:
    """`sidecar-search index train DATA_DIR` (/root/reference/Makefile:38-39): train on a sample of
    the store.  faiss itself subsamples to 256 x nlist rows; reading more than that from disk is
    wasted I/O, so whole row groups are drawn (seeded) until that many rows are gathered."""
    cap = max_rows or index.nlist * index.cp.max_points_per_centroid
    groups = list(iter_row_groups(dir_path, index.d, columns=("embedding",)))
    order = np.random.RandomState(seed).permutation(len(groups))
    take, rows = [], 0
    for g in order:
        take.append(groups[g][1])
        rows += take[-1].shape[0]
        if rows >= cap:
            break
    x = np.concatenate(take, axis=0)
    index.train(x)
    return x.shape[0]
