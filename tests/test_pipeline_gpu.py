"""The whole reference pipeline on one GPU, each stage through the product's drop-in surface and
checked against the oracle chain on the same input:

    OpenAlex works JSONL --oa_jsonl--> {"id","document"} --SentenceTransformer.encode(b=32)-->
    embeddings --parquet shards--> index train / fill (ids.parquet) --> query encode (s2p_query) + search

Reference call sites: Makefile:64 (oa_jsonl), :65 (build -b 32), :48 (dump to parquet shards), :38-39
(train), :24-25 (fill), README.md:16,28 (app.py query).  Tolerances: text stage byte-exact, embedding
cosine within 1e-3 of the fp32 oracle (north star), returned document ids identical to the oracle's
wherever its fp64 margin is clear of bf16 encoder noise (and the self-match always first)."""
import io

import numpy as np
import pytest

from oracle import encoder as oenc
from oracle import ivf as oivf
from oracle import oa_jsonl as ooa
from tiny_cfg import TINY

pytestmark = pytest.mark.gpu


def test_openalex_to_search_pipeline_matches_oracle_chain(gpu_pkg, tmp_path):
    P = gpu_pkg
    raw = P.oa_jsonl.synth_records(77, 900, mean_words=40, filler=2)
    # stage 1: front end, byte-exact against the restatement of the reference program
    docs_bytes = P.oa_jsonl.convert(raw)
    assert docs_bytes == ooa.convert(raw)
    pairs = list(P.oa_jsonl.iter_documents(io.BytesIO(raw), block_bytes=1 << 16))
    ids, docs = [p[0] for p in pairs], [p[1] for p in pairs]
    n = len(docs)
    assert n > 300 and len(set(ids)) == n

    # stage 2: bulk encode, batch 32 (tiny config of the true architecture, seeded weights)
    sd = oenc.random_state_dict(TINY, seed=3, std=0.05)
    cfg = P.EncoderConfig(**{f: getattr(TINY, f) for f in TINY.__dataclass_fields__})
    enc = P.SentenceTransformer(config=cfg)
    for name, arr in sd.items():
        enc.load_weight(name, arr)
    emb = enc.encode(docs, batch_size=32, normalize_embeddings=True)
    assert emb.shape == (n, TINY.embed_dim) and emb.dtype == np.float32
    ref = np.empty_like(emb)
    for b0 in range(0, n, 64):
        feats = enc.tokenize(docs[b0:b0 + 64])
        ref[b0:b0 + 64] = oenc.forward_plain(TINY, sd, feats["input_ids"], feats["attention_mask"], normalize=True)
    cos = oenc.cosine_rows(emb, ref)
    assert (1 - cos).max() < 1e-3, f"embedding cosine vs fp32 oracle: min {cos.min()}"

    # stage 3: embedding store -> train -> fill (+ ids.parquet), then compaction
    P.store.write_shards(str(tmp_path / "data"), ids, emb, shard_size=400, row_group_size=64)
    nlist, k, nprobe = 16, 5, 4
    ix = P.index_factory(TINY.embed_dim, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    P.store.train_index(ix, str(tmp_path / "data"))
    assert P.store.fill_index(ix, str(tmp_path / "data"), ids_parquet=str(tmp_path / "ids.parquet")) == n
    assert ix.ntotal == n
    doc_ids = P.faiss_io.read_ids_parquet(str(tmp_path / "ids.parquet"))
    assert doc_ids == ids

    # stage 4: app.py's query loop: encode(query, prompt_name="s2p_query") -> search -> document ids
    pick = list(range(0, n, 37))
    queries = [docs[i] for i in pick]
    qe = enc.encode(queries, normalize_embeddings=True)  # same text, no prompt: must find itself
    ix.nprobe = nprobe
    D, I = ix.search(qe, k)
    assert [doc_ids[i] for i in I[:, 0]] == [ids[i] for i in pick], "a document must be its own nearest neighbour"
    assert np.all(D[:, 0] > 0.999)
    qp = enc.encode(queries[0], prompt_name="s2p_query", normalize_embeddings=True)
    assert qp.shape == (TINY.embed_dim,) and abs(float(np.linalg.norm(qp)) - 1.0) < 1e-5

    # oracle chain on the PRODUCT's embeddings and centroids: ids and scores of the index half agree
    # exactly where the oracle's fp64 margin allows, scores within 1e-5
    o = oivf.IVFFlat(TINY.embed_dim, nlist)
    o.set_centroids(ix.get_centroids())
    emb16 = np.concatenate([e for _, e in P.store.iter_row_groups(str(tmp_path / "data"), TINY.embed_dim)])
    o.add(emb16)
    Do, Io = o.search(qe, k, nprobe=nprobe)
    cm, fm = o.ambiguity(qe, k, nprobe)
    clear = (cm > 1e-5) & (fm > 1e-5)
    assert clear.sum() >= len(pick) // 2, "too few unambiguous queries to compare"
    assert np.array_equal(I[clear], Io[clear]) and np.abs(D[clear] - Do[clear]).max() < 1e-5


@pytest.mark.parametrize("coresident", [True, False])
def test_query_pipeline_equals_serial_query_loop(gpu_pkg, coresident):
    """QueryPipeline (encode of batch i+1 on one stream while batch i is searched on another; with
    `coresident` the scan runs as one small register-capped ring CTA per SM next to GEMM CTAs that use a
    smaller operand ring) returns, batch for batch, the bits the serial app.py-style loop returns:
    device tensors in / out, and pinned host ids in / numpy out with the copies inside the pipeline."""
    import torch

    from oracle import synth as osynth

    P = gpu_pkg
    d, nlist, n, nq, S, k, nprobe, nb = 1024, 256, 40000, 64, 16, 10, 8, 5
    enc = P.Encoder(config=P.STELLA_1_5B, random_init_seed=0)
    ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
    ix.set_two_stage(64)
    ix.set_centroids(osynth.centroids(5, nlist, d))
    ix.add(osynth.corpus_unit(5, 0, n, d, nlist))
    ix.nprobe = nprobe
    g = torch.Generator().manual_seed(3)
    ids_h = torch.randint(0, P.STELLA_1_5B.vocab_size, (nb, nq, S), generator=g, dtype=torch.int64).pin_memory()
    mask_h = torch.ones((nq, S), dtype=torch.int32)
    mask_h[::3, S - 5:] = 0
    mask_h = mask_h.pin_memory()
    ids_d, mask_d = ids_h.cuda(), mask_h.cuda()
    want = []
    for b in range(nb):
        e = enc.encode_tokens(ids_d[b], mask_d, normalize_embeddings=True)
        D, I = ix.search(e, k)
        want.append((D.cpu().numpy(), I.cpu().numpy()))
    assert len({w[1].tobytes() for w in want}) == nb  # the batches really differ
    try:
        pipe = P.QueryPipeline(enc, ix, k=k, nprobe=nprobe, batch=nq, tokens=S, coresident=coresident)
        got = [(D.cpu().numpy(), I.cpu().numpy()) for D, I in pipe.run((ids_d[b], mask_d) for b in range(nb))]
        pipe.join()
        assert len(got) == nb
        for b in range(nb):
            assert np.array_equal(got[b][1], want[b][1]) and np.array_equal(got[b][0], want[b][0]), b
        got_h = list(pipe.run((ids_h[b], mask_h) for b in range(nb)))
        for b in range(nb):
            assert isinstance(got_h[b][1], np.ndarray)
            assert np.array_equal(got_h[b][1], want[b][1]) and np.array_equal(got_h[b][0], want[b][0]), b
    finally:
        assert P.lib().absb_gemm_set_smem_budget(0) == 0  # process-wide setting: leave it as the other tests expect it
    # the reduced GEMM operand ring alone changes no bit of the encoder output either
    ref = enc.encode_tokens(ids_d[0], mask_d, normalize_embeddings=True).clone()
    assert P.lib().absb_gemm_set_smem_budget(128 * 1024) == 0
    try:
        small = enc.encode_tokens(ids_d[0], mask_d, normalize_embeddings=True)
        assert torch.equal(small, ref)
    finally:
        assert P.lib().absb_gemm_set_smem_budget(0) == 0
