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


def list_row_groups(dir_path: str) -> list[tuple[str, int, int]]:
    """(file, row group, rows) of every row group of the store, from the parquet FOOTERS only — no
    column data is read."""
    import pyarrow.parquet as pq

    files = sorted(glob.glob(os.path.join(dir_path, "*.parquet")))
    if not files:
        raise RuntimeError(f"no parquet shards under {dir_path}")
    out = []
    for path in files:
        md = pq.ParquetFile(path).metadata
        out.extend((path, g, md.row_group(g).num_rows) for g in range(md.num_row_groups))
    return out


def _read_group(path: str, g: int, d: int | None, columns):
    import pyarrow.parquet as pq

    t = pq.ParquetFile(path).read_row_group(g, columns=list(columns))
    ids = t.column("id") if "id" in t.column_names else None
    return ids, _embedding_matrix(t.column("embedding"), d)


def _prefetched(groups, d, columns, depth: int = 2):
    """Row groups decoded by a background thread `depth` ahead of the consumer: parquet decode + fp16 ->
    fp32 conversion of group i+1 (pyarrow and numpy release the GIL) overlap the H2D copy and the GPU
    add() of group i (ctypes releases the GIL for the C call)."""
    import queue
    import threading

    q: queue.Queue = queue.Queue(maxsize=depth)

    def work():
        try:
            for path, g, _ in groups:
                q.put(_read_group(path, g, d, columns))
            q.put(None)
        except BaseException as e:  # noqa: BLE001 - re-raised in the consumer
            q.put(e)

    th = threading.Thread(target=work, daemon=True)
    th.start()
    while True:
        item = q.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        yield item
    th.join()


def fill_index(index, dir_path: str, ids_parquet: str | None = None) -> int:
    """`sidecar-search index fill DATA_DIR` (/root/reference/Makefile:24-25): stream every shard into
    index.add(); faiss ids are the running row numbers, and `ids.parquet` (row -> document id) is
    written next to the index when a path is given.  Memory stays bounded: two decoded row groups in
    flight, and the id column goes to the ids.parquet writer one row group at a time (the 207M-row
    corpus would otherwise leave 207M Python strings in a list)."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    groups = list_row_groups(dir_path)
    writer, n = None, 0
    try:
        for ids, x in _prefetched(groups, index.d, ("id", "embedding")):
            index.add(x)
            n += x.shape[0]
            if ids_parquet is not None:
                t = pa.table({"id": ids})
                if writer is None:
                    writer = pq.ParquetWriter(ids_parquet, t.schema)
                writer.write_table(t)
    finally:
        if writer is not None:
            writer.close()
    if ids_parquet is not None and writer is None:
        from .faiss_io import write_ids_parquet

        write_ids_parquet(ids_parquet, [])
    if hasattr(index, "compact"):
        index.compact()  # row-group-sized add() calls leave every list scattered over the page pool
    return n


def train_index(index, dir_path: str, max_rows: int | None = None, seed: int = 1234) -> int:
    """`sidecar-search index train DATA_DIR` (/root/reference/Makefile:38-39): train on a sample of
    the store.  faiss itself subsamples to 256 x nlist rows; reading more than that from disk is
    wasted I/O, so whole row groups are drawn (seeded permutation of the (file, row group) pairs listed
    from the parquet footers) and ONLY the drawn groups are read, until that many rows are gathered —
    68 GB of fp32 for IVF65536 however large the store is (the reference trains on a 16 GB box with
    faiss's own on-disk sampling, Makefile:34-39)."""
    cap = max_rows or index.nlist * index.cp.max_points_per_centroid
    groups = list_row_groups(dir_path)
    order = np.random.RandomState(seed).permutation(len(groups))
    chosen, rows = [], 0
    for g in order:
        chosen.append(groups[g])
        rows += groups[g][2]
        if rows >= cap:
            break
    d = index.d
    x = np.empty((rows, d), dtype=np.float32)
    r0 = 0
    for _, xg in _prefetched(chosen, d, ("embedding",)):
        x[r0:r0 + xg.shape[0]] = xg
        r0 += xg.shape[0]
    assert r0 == rows
    index.train(x)
    return rows
