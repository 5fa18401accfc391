"""CPU tests of the C-ABI boundary and the Python host logic (no compute: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_pkg


@pytest.fixture(scope="module")
def P():
    p = load_pkg()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


def test_library_exports_every_declared_symbol(P):
    names = P.header_functions()
    assert len(names) >= 50
    handle = ctypes.CDLL(P.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in include/absb200.h but not exported: {missing}"
    # and the ctypes table covers the header exactly
    from importlib import import_module

    sigs = import_module("abstracts-search_b200._lib")._SIGS
    assert sorted(sigs) == names


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, "include", "absb200.h")).read()
    assert "extern \"C\"" in src
    assert len(re.findall(r"/root/reference/(Makefile|README\.md):\d+", src)) >= 4
    assert "torch" not in src.lower() and "at::" not in src


def test_version_and_error_string(P):
    L = P.lib()
    assert L.absb_version() == 100
    n = ctypes.c_int(-7)
    rc = L.absb_device_count(ctypes.byref(n))
    import torch

    if not torch.cuda.is_available():
        # no device: the library reports a CUDA error instead of pretending
        assert rc != 0 and len(L.absb_last_error()) > 0


def test_product_path_fails_loudly_without_gpu(P):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        P.index_factory(64, "IVF16,Flat", P.METRIC_INNER_PRODUCT)
    with pytest.raises(RuntimeError):
        P.IndexFlatIP(64)
    with pytest.raises(RuntimeError):
        P.Encoder(config=P.EncoderConfig(vocab_size=100, hidden_size=256, num_layers=1, num_heads=2, num_kv_heads=1,
                                         intermediate_size=256, embed_dim=64), random_init_seed=0)


def test_no_oracle_import_in_product():
    pk = os.path.join(ROOT, "abstracts-search_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", s, flags=re.M), f
                assert "liboracle" not in s and "oracle/" not in s.replace("oracle/synth.py", "").replace(
                    "oracle/ivf_oracle.c", ""), f


def test_index_factory_parsing(P):
    with pytest.raises(RuntimeError):
        P.index_factory(64, "IVF16,Flat", P.METRIC_L2)
    with pytest.raises(RuntimeError):
        P.index_factory(64, "IVF16,PQ8", P.METRIC_INNER_PRODUCT)
    with pytest.raises(RuntimeError):
        P.index_factory(64, "HNSW32", P.METRIC_INNER_PRODUCT)


def test_numpy_contract_helpers(P):
    from importlib import import_module

    ix = import_module("abstracts-search_b200.index")
    x = ix._as_f32_matrix(np.arange(12, dtype=np.float64).reshape(3, 4), 4)
    assert x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]
    x = ix._as_f32_matrix(np.asfortranarray(np.ones((3, 4), dtype=np.float32)), 4)
    assert x.flags["C_CONTIGUOUS"]
    with pytest.raises(AssertionError):
        ix._as_f32_matrix(np.ones((3, 5), dtype=np.float32), 4)
    with pytest.raises(ValueError):
        ix._as_f32_matrix(np.ones(4, dtype=np.float32), 4)
    with pytest.raises(AssertionError):
        ix._as_i64_vector(np.arange(4), 3, x)


def test_merge_partials_host(P):
    rng = np.random.default_rng(0)
    world, n, k = 3, 5, 4
    D = -np.sort(-rng.standard_normal((world, n, k)).astype(np.float32), axis=2)
    I = rng.permutation(world * n * k).reshape(world, n, k).astype(np.int64)
    D[1, 2, 2:] = -3.4028234663852886e38
    I[1, 2, 2:] = -1
    Dm, Im = P.merge_partials_host(D, I, k)
    for q in range(n):
        cand = [(-(D[w, q, j]), I[w, q, j]) for w in range(world) for j in range(k) if I[w, q, j] >= 0]
        cand.sort()
        assert [c[1] for c in cand[:k]] == Im[q].tolist()
    # exact ties resolve by id
    D2 = np.zeros((2, 1, 2), dtype=np.float32)
    I2 = np.array([[[9, 4]], [[7, 1]]], dtype=np.int64)
    assert P.merge_partials_host(D2, I2, 3)[1][0].tolist() == [1, 4, 7]


def test_encoder_host_logic(P):
    cfg = P.EncoderConfig()
    shapes = cfg.param_shapes()
    n_params = sum(int(np.prod(s)) for n, s in shapes.items() if not n.startswith("dense."))
    assert n_params == 1_543_268_864  # Qwen2-1.5B backbone incl. embeddings
    assert abs(cfg.flops_per_token_linear() - 2.6204e9) / 2.6204e9 < 1e-3
    tok = import_tok(P)
    out = tok(["a b c", "hello, world"], max_length=3)
    assert out["input_ids"].shape == (2, 3) and out["attention_mask"].sum() == 6
    out = tok(["a", "b c d"])
    assert out["attention_mask"].tolist() == [[1, 0, 0], [1, 1, 1]]


def import_tok(P):
    from importlib import import_module

    return import_module("abstracts-search_b200.encoder").HashTokenizer(1000)


def test_safetensors_reader(tmp_path, P):
    import json
    import struct
    from importlib import import_module

    enc = import_module("abstracts-search_b200.encoder")
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([0x3F80, 0x4000], dtype=np.uint16)  # bf16 1.0, 2.0
    header = {"w": {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]},
              "v": {"dtype": "BF16", "shape": [2], "data_offsets": [24, 28]}, "__metadata__": {"format": "pt"}}
    hj = json.dumps(header).encode()
    path = tmp_path / "m.safetensors"
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + a.tobytes() + b.tobytes())
    t = enc.read_safetensors(str(path))
    assert np.array_equal(t["w"][0], a) and t["w"][1] == "F32"
    assert np.array_equal(t["v"][0], b) and t["v"][1] == "BF16"


# ---------------------------------------------------------------------------------------------
# in-place page compaction schedule (absb_ivf_compact): host-side planner, simulated here
# ---------------------------------------------------------------------------------------------
def _plan(src, scratch):
    import ctypes

    import numpy as np

    L = load_pkg().lib()
    n = len(src)
    src = np.ascontiguousarray(src, dtype=np.int32)
    moves = np.empty((3 * n + 1, 2), dtype=np.int32)
    phases = np.empty(3 * n + 3, dtype=np.int64)
    nm, nph = ctypes.c_int64(), ctypes.c_int64()
    rc = L.absb_plan_page_compaction(n, src.ctypes.data, scratch, moves.ctypes.data, len(moves), phases.ctypes.data,
                                     len(phases), ctypes.byref(nm), ctypes.byref(nph))
    assert rc == 0, L.absb_last_error()
    return moves[: nm.value], phases[: nph.value]


@pytest.mark.parametrize("n,scratch,seed", [(1, 1, 0), (2, 1, 1), (17, 1, 2), (64, 5, 3), (1000, 64, 4), (1000, 1000, 5),
                                            (4096, 100, 6), (333, 7, 7)])
def test_compaction_plan_permutes_pages_in_place(n, scratch, seed):
    import numpy as np

    rng = np.random.default_rng(seed)
    src = rng.permutation(n).astype(np.int32)
    if seed % 2:  # a partly ordered table, as left by a few large add() calls
        src[: n // 2] = np.sort(src[: n // 2])
    _check_plan(src, scratch)


def _check_plan(src, scratch):
    """Replays a plan on a pool whose page p holds 1000 + p: phases must be internally race-free and the pool must end
    up as content_new[t] == content_old[src[t]]."""
    import numpy as np

    n = len(src)
    moves, phases = _plan(src, scratch)
    pool = np.arange(n, dtype=np.int64) + 1000  # page p holds content 1000 + p
    sc = np.full(scratch, -1, dtype=np.int64)
    begin = 0
    for end in phases:
        ph = moves[begin:end]
        begin = end
        reads, writes = set(ph[:, 0].tolist()), ph[:, 1].tolist()
        assert len(set(writes)) == len(writes), "two copies of one phase write the same page"
        assert not reads & set(writes), "a phase reads a page it also writes"
        vals = [pool[f] if f >= 0 else sc[-1 - f] for f in ph[:, 0]]
        for (f, t), v in zip(ph, vals):
            assert -scratch <= min(f, t) and max(f, t) < n
            if t >= 0:
                pool[t] = v
            else:
                sc[-1 - t] = v
    assert begin == len(moves)
    assert np.array_equal(pool, src.astype(np.int64) + 1000), "content_new[t] must equal content_old[src[t]]"
    assert len(moves) <= 3 * n


def test_compaction_plan_fuzz():
    """hypothesis: permutations built from the shapes real page tables have — long ordered runs (large add() calls),
    rotations, swapped blocks, fixed points, plus uniformly random ones — under scratch sizes from 1 page to > n."""
    import numpy as np
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(n=st.integers(1, 400), scratch=st.integers(1, 450), kind=st.integers(0, 4), seed=st.integers(0, 2**31))
    def run(n, scratch, kind, seed):
        rng = np.random.default_rng(seed)
        if kind == 0:
            src = rng.permutation(n)
        elif kind == 1:  # rotation: one cycle through every page
            src = np.roll(np.arange(n), int(rng.integers(0, n)))
        elif kind == 2:  # interleave of a few ordered runs (lists appended to in turn)
            runs = int(rng.integers(1, 6))
            owner = rng.integers(0, runs, n)
            src = np.argsort(owner, kind="stable")
        elif kind == 3:  # mostly fixed points, a few transpositions
            src = np.arange(n)
            for _ in range(int(rng.integers(0, 5))):
                i, j = rng.integers(0, n, 2)
                src[[i, j]] = src[[j, i]]
        else:  # two swapped blocks
            cut = int(rng.integers(0, n + 1))
            src = np.concatenate([np.arange(cut, n), np.arange(0, cut)])
        _check_plan(src.astype(np.int32), scratch)

    run()


def test_compaction_plan_skips_an_ordered_table_and_rejects_garbage():
    import numpy as np

    moves, phases = _plan(np.arange(100), 8)
    assert len(moves) == 0 and len(phases) == 0
    L = load_pkg().lib()
    import ctypes

    bad = np.array([0, 0, 1], dtype=np.int32)
    out = np.empty((16, 2), dtype=np.int32)
    ph = np.empty(16, dtype=np.int64)
    a, b = ctypes.c_int64(), ctypes.c_int64()
    assert L.absb_plan_page_compaction(3, bad.ctypes.data, 2, out.ctypes.data, 16, ph.ctypes.data, 16,
                                       ctypes.byref(a), ctypes.byref(b)) != 0


def test_encode_host_logic_fuzz(P):
    """hypothesis: `SentenceTransformer.encode` host logic — prompt prefix, sort by text length (descending, stable),
    batches of `batch_size`, results restored to input order, str in -> vector out — with the device forward replaced
    by a function of each row's own (unpadded) token ids."""
    from importlib import import_module

    from hypothesis import given, settings, strategies as st

    E = import_module("abstracts-search_b200.encoder")
    enc = E.Encoder.__new__(E.Encoder)  # no device handle: only the host side is exercised
    enc.config = P.EncoderConfig()
    enc.prompts = {"s2p_query": "Instruct: find. Query: ", "empty": ""}
    enc.default_prompt_name = None
    enc.max_seq_length = 12
    enc.tokenizer = E.HashTokenizer(1000)
    enc.device = "cuda:0"
    enc._h = None
    seen_batches = []

    def row_embedding(ids):
        v = np.zeros(enc.config.embed_dim, dtype=np.float32)
        for j, t in enumerate(ids):
            v[(int(t) * 31 + j) % v.size] += 1.0 + j
        return v

    def fake_forward(input_ids, attention_mask=None, normalize_embeddings=False):
        seen_batches.append(input_ids.shape)
        assert input_ids.shape[1] <= enc.max_seq_length and attention_mask.shape == input_ids.shape
        # right padding only, never a fully padded column
        assert (np.diff(attention_mask, axis=1) <= 0).all() and attention_mask[:, 0].all() and attention_mask.any(axis=0).all()
        return np.stack([row_embedding(r[m == 1]) for r, m in zip(input_ids, attention_mask)])

    enc.encode_tokens = fake_forward
    words = st.text(alphabet="abc xyz,.", min_size=0, max_size=40)

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(sentences=st.lists(words, min_size=0, max_size=20), batch_size=st.integers(1, 7),
           prompt_name=st.sampled_from([None, "s2p_query", "empty"]))
    def run(sentences, batch_size, prompt_name):
        seen_batches.clear()
        out = enc.encode(sentences, prompt_name=prompt_name, batch_size=batch_size)
        assert out.shape == (len(sentences), enc.config.embed_dim)
        prefix = enc.prompts[prompt_name] if prompt_name else ""
        for s, got in zip(sentences, out):
            want = row_embedding(enc.tokenizer.encode(prefix + s)[: enc.max_seq_length])
            assert np.array_equal(got, want)
        assert len(seen_batches) == -(-len(sentences) // batch_size)
        assert all(b[0] <= batch_size for b in seen_batches)
        if sentences:
            # longest texts first: sequence lengths of the batches never increase by more than truncation allows
            one = enc.encode(sentences[0], prompt_name=prompt_name)
            assert one.shape == (enc.config.embed_dim,) and np.array_equal(one, out[0])

    run()
    import pytest

    with pytest.raises(ValueError):
        enc.encode(["a"], prompt_name="no_such_prompt")
    with pytest.raises(ValueError):
        enc.encode(["a"], precision="int8")


def test_safetensors_reader_fuzz(tmp_path, P):
    """hypothesis: files in the safetensors layout (u64 header length, JSON header padded with spaces as the real
    writer does, tensors at arbitrary — also unaligned — offsets, F32 / BF16 / F16, scalars, empty tensors, unknown
    dtypes skipped) read back exactly."""
    import json
    import struct
    from importlib import import_module

    from hypothesis import HealthCheck, given, settings, strategies as st

    enc = import_module("abstracts-search_b200.encoder")
    case = [0]
    shape = st.lists(st.integers(0, 5), min_size=0, max_size=3)
    tensor = st.tuples(st.sampled_from(["F32", "BF16", "F16", "I64"]), shape)

    @settings(max_examples=80, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(tensors=st.lists(tensor, min_size=0, max_size=6), pad=st.integers(0, 9), gap=st.integers(0, 3), seed=st.integers(0, 2**31))
    def run(tensors, pad, gap, seed):
        rng = np.random.default_rng(seed)
        header, blobs, want, off = {"__metadata__": {"format": "pt"}}, [], {}, 0
        for i, (dt, shp) in enumerate(tensors):
            n = int(np.prod(shp)) if shp else 1
            if dt == "F32":
                a = rng.standard_normal(n).astype(np.float32)
                want[f"t{i}"] = (a.reshape(shp), "F32")
            elif dt == "BF16":
                a = rng.integers(0, 1 << 16, n).astype(np.uint16)
                want[f"t{i}"] = (a.reshape(shp), "BF16")
            elif dt == "F16":
                a = rng.standard_normal(n).astype(np.float16)
                want[f"t{i}"] = (a.astype(np.float32).reshape(shp), "F32")
            else:
                a = rng.integers(0, 100, n).astype(np.int64)  # not a weight dtype: ignored by the reader
            raw = a.tobytes() + b"\0" * gap  # gaps make the next tensor's offset unaligned
            header[f"t{i}"] = {"dtype": dt, "shape": shp, "data_offsets": [off, off + a.nbytes]}
            blobs.append(raw)
            off += len(raw)
        hj = json.dumps(header).encode() + b" " * pad
        case[0] += 1
        path = tmp_path / f"m{case[0]}.safetensors"
        path.write_bytes(struct.pack("<Q", len(hj)) + hj + b"".join(blobs))
        got = enc.read_safetensors(str(path))
        assert set(got) == set(want)
        for name, (arr, dt) in want.items():
            assert got[name][1] == dt and got[name][0].shape == arr.shape
            assert np.array_equal(np.asarray(got[name][0]), arr)

    run()
