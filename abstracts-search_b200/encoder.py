"""SentenceTransformer-shaped surface over the sm_100a stella_en_1.5B_v5 encoder.

Mirrors what `sidecar-search build -b 32` (/root/reference/Makefile:65, README.md:60) and app.py
(/root/reference/README.md:28: MODEL_NAME, PROMPT_NAME=s2p_query, TRUST_REMOTE_CODE) call:

    model = SentenceTransformer("NovaSearch/stella_en_1.5B_v5", trust_remote_code=True)
    emb = model.encode(docs, batch_size=32)                       # np.float32 [n, 1024]
    q = model.encode(query, prompt_name="s2p_query")              # np.float32 [1024]

encode() = prepend prompt -> sort by length (desc) -> per batch: tokenize (pad to longest,
truncate to max_seq_length) -> Qwen2 backbone (bidirectional) -> mean pool -> Dense 1536->1024
-> optional L2 normalise -> restore order.  All arithmetic after tokenisation runs in
libabsb200.so on the GPU; there is no PyTorch/CPU model fallback.
"""
from __future__ import annotations

import json
import os
import struct
from ctypes import byref, c_double, c_int64, c_void_p
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import EncConfig, check, lib, ptr


@dataclass
class EncoderConfig:
    vocab_size: int = 151646
    hidden_size: int = 1536
    num_layers: int = 28
    num_heads: int = 12
    num_kv_heads: int = 2
    head_dim: int = 128
    intermediate_size: int = 8960
    embed_dim: int = 1024
    max_seq_len: int = 512
    causal: bool = False  # stella runs the Qwen2 stack with a padding-only (bidirectional) mask
    rms_eps: float = 1e-6
    rope_theta: float = 1e6

    def to_c(self) -> EncConfig:
        return EncConfig(self.vocab_size, self.hidden_size, self.num_layers, self.num_heads, self.num_kv_heads,
                         self.head_dim, self.intermediate_size, self.embed_dim, self.max_seq_len, int(self.causal),
                         self.rms_eps, self.rope_theta)

    def param_shapes(self) -> dict:
        """Hugging Face parameter names -> shapes."""
        H, I, hd = self.hidden_size, self.intermediate_size, self.head_dim
        s = {"embed_tokens.weight": (self.vocab_size, H), "norm.weight": (H,),
             "dense.weight": (self.embed_dim, H), "dense.bias": (self.embed_dim,)}
        for l in range(self.num_layers):
            p = f"layers.{l}."
            s[p + "input_layernorm.weight"] = (H,)
            s[p + "post_attention_layernorm.weight"] = (H,)
            s[p + "self_attn.q_proj.weight"] = (self.num_heads * hd, H)
            s[p + "self_attn.q_proj.bias"] = (self.num_heads * hd,)
            s[p + "self_attn.k_proj.weight"] = (self.num_kv_heads * hd, H)
            s[p + "self_attn.k_proj.bias"] = (self.num_kv_heads * hd,)
            s[p + "self_attn.v_proj.weight"] = (self.num_kv_heads * hd, H)
            s[p + "self_attn.v_proj.bias"] = (self.num_kv_heads * hd,)
            s[p + "self_attn.o_proj.weight"] = (H, self.num_heads * hd)
            s[p + "mlp.gate_proj.weight"] = (I, H)
            s[p + "mlp.up_proj.weight"] = (I, H)
            s[p + "mlp.down_proj.weight"] = (H, I)
        return s

    def flops_per_token_linear(self) -> float:
        H, I, hd = self.hidden_size, self.intermediate_size, self.head_dim
        per_layer = H * (self.num_heads + 2 * self.num_kv_heads) * hd + self.num_heads * hd * H + 3 * H * I
        return 2.0 * per_layer * self.num_layers


STELLA_1_5B = EncoderConfig()

STELLA_PROMPTS = {
    "s2p_query": "Instruct: Given a web search query, retrieve relevant passages that answer the query.\nQuery: ",
    "s2s_query": "Instruct: Retrieve semantically similar text.\nQuery: ",
}


class HashTokenizer:
    """Offline stand-in for the Qwen2 BPE tokenizer (its files are fetched from the HF Hub in the
    reference and are not available without network).  Whitespace/punctuation pieces hashed into
    the vocabulary; deterministic.  Pass a real tokenizer to Encoder(tokenizer=...) in deployment."""

    def __init__(self, vocab_size: int, pad_token_id: int = 0):
        self.vocab_size, self.pad_token_id = vocab_size, pad_token_id
        self.padding_side = "right"

    def encode(self, text: str) -> list[int]:
        import re
        import zlib

        pieces = re.findall(r"\w+|[^\w\s]", text)
        return [1 + zlib.crc32(p.encode("utf-8")) % (self.vocab_size - 1) for p in pieces] or [1]

    def __call__(self, texts, padding=True, truncation=True, max_length=512, return_tensors="np"):
        toks = [self.encode(t)[:max_length] for t in texts]
        L = max(len(t) for t in toks)
        ids = np.full((len(toks), L), self.pad_token_id, dtype=np.int64)
        mask = np.zeros((len(toks), L), dtype=np.int64)
        for i, t in enumerate(toks):
            ids[i, : len(t)] = t
            mask[i, : len(t)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def read_safetensors(path: str) -> dict:
    """Minimal safetensors reader: name -> (np.ndarray view or raw bf16 uint16 array, dtype str)."""
    out = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = 8 + hlen
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            b, e = meta["data_offsets"]
            raw = mm[base + b: base + e]
            dt = meta["dtype"]
            if dt == "F32":
                arr = raw.view(np.float32)
            elif dt == "BF16":
                arr = raw.view(np.uint16)
            elif dt == "F16":
                arr = raw.view(np.float16).astype(np.float32)
                dt = "F32"
            else:
                continue
            out[name] = (arr.reshape(meta["shape"]), dt)
    return out


class Encoder:
    """The B200 encoder with a SentenceTransformer-compatible surface."""

    def __init__(self, model_name_or_path: str | None = None, device: str | int = "cuda:0",
                 trust_remote_code: bool = False, config: EncoderConfig | None = None, tokenizer=None,
                 prompts: dict | None = None, random_init_seed: int | None = None, random_init_std: float = 0.02):
        self.config = config or STELLA_1_5B
        self.device_index = int(str(device).split(":")[1]) if isinstance(device, str) and ":" in device else (
            device if isinstance(device, int) else 0)
        self.device = f"cuda:{self.device_index}"
        self.prompts = dict(STELLA_PROMPTS if prompts is None else prompts)
        self.default_prompt_name = None
        self.max_seq_length = self.config.max_seq_len
        self.trust_remote_code = trust_remote_code
        self.model_name_or_path = model_name_or_path
        self._h = c_void_p()
        cfg = self.config.to_c()
        check(lib().absb_enc_create(byref(cfg), self.device_index, byref(self._h)))
        self.tokenizer = tokenizer
        if model_name_or_path is not None and os.path.isdir(model_name_or_path):
            self._load_dir(model_name_or_path)
        elif random_init_seed is not None:
            self.init_random(random_init_seed, random_init_std)
        elif model_name_or_path is not None:
            raise RuntimeError(
                f"{model_name_or_path!r} is not a local directory and there is no network access to the HF Hub; "
                "pass a local snapshot directory, or random_init_seed=... for the synthetic-weights model")
        if self.tokenizer is None:
            self.tokenizer = HashTokenizer(self.config.vocab_size)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().absb_enc_destroy(h)

    # ---- weights -------------------------------------------------------------------------
    def init_random(self, seed: int = 0, std: float = 0.02):
        check(lib().absb_enc_init_random(self._h, seed, std))

    def load_weight(self, name: str, array, bf16_raw: bool = False):
        a = np.ascontiguousarray(array)
        if bf16_raw:
            assert a.dtype == np.uint16
            dtype = 1
        else:
            a = np.ascontiguousarray(a, dtype=np.float32)
            dtype = 0
        shape = (c_int64 * a.ndim)(*a.shape)
        check(lib().absb_enc_load_weight(self._h, name.encode(), ptr(a), dtype, shape, a.ndim))

    def get_weight(self, name: str) -> np.ndarray:
        shape = self.config.param_shapes()[name]
        out = np.empty(shape, dtype=np.float32)
        check(lib().absb_enc_get_weight(self._h, name.encode(), ptr(out), out.size))
        return out

    def state_dict(self) -> dict:
        return {n: self.get_weight(n) for n in self.config.param_shapes()}

    def _load_dir(self, path: str):
        """A sentence-transformers snapshot: model.safetensors (+ shards) and 2_Dense_*/model.safetensors."""
        files = [os.path.join(path, f) for f in sorted(os.listdir(path)) if f.endswith(".safetensors")]
        for sub in sorted(os.listdir(path)):
            d = os.path.join(path, sub)
            if os.path.isdir(d) and "Dense" in sub:
                files += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".safetensors")]
        if not files:
            raise RuntimeError(f"no .safetensors files under {path}")
        known = set(self.config.param_shapes())
        loaded = set()
        for fpath in files:
            for name, (arr, dt) in read_safetensors(fpath).items():
                short = name[6:] if name.startswith("model.") else name
                short = {"linear.weight": "dense.weight", "linear.bias": "dense.bias"}.get(short, short)
                if short in known:
                    self.load_weight(short, arr, bf16_raw=(dt == "BF16"))
                    loaded.add(short)
        missing = sorted(known - loaded)
        if missing:
            # e.g. a snapshot without its 2_Dense_1024 folder: embeddings from uninitialised weights
            # would be meaningless and nothing downstream could tell
            raise RuntimeError(f"{path}: {len(missing)} parameters of the model are in none of the safetensors files "
                               f"(first: {missing[:4]})")
        if self.tokenizer is None:
            # real weights with the hash stand-in would produce garbage embeddings silently: a model
            # directory must bring its tokenizer (or the caller passes tokenizer=...)
            try:
                from transformers import AutoTokenizer

                self.tokenizer = AutoTokenizer.from_pretrained(path, trust_remote_code=self.trust_remote_code)
            except Exception as e:
                raise RuntimeError(f"{path}: the tokenizer could not be loaded ({e!r}); pass tokenizer=... explicitly") from e

    # ---- SentenceTransformer surface -------------------------------------------------------
    def get_sentence_embedding_dimension(self) -> int:
        return self.config.embed_dim

    def get_max_seq_length(self) -> int:
        return self.max_seq_length

    def to(self, *_a, **_k):
        return self

    def half(self):
        return self

    def bfloat16(self):
        return self

    def eval(self):
        return self

    def tokenize(self, texts):
        out = self.tokenizer(list(texts), padding=True, truncation=True, max_length=self.max_seq_length,
                             return_tensors="np")
        return {"input_ids": np.asarray(out["input_ids"], dtype=np.int64),
                "attention_mask": np.asarray(out["attention_mask"], dtype=np.int64)}

    def encode_tokens(self, input_ids, attention_mask=None, normalize_embeddings: bool = False):
        """Tokenizer-free entry: input_ids [B, S] (+ attention_mask [B, S], 1 = token) -> [B, dim].
        numpy in -> numpy out (copies inside the C call); CUDA tensors in -> CUDA tensor out."""
        if hasattr(input_ids, "is_cuda") and input_ids.is_cuda:
            import torch

            ids = input_ids.to(torch.int64).contiguous()
            B, S = ids.shape
            mask = (torch.ones((B, S), dtype=torch.int32, device=ids.device) if attention_mask is None
                    else attention_mask.to(torch.int32).contiguous())
            out = torch.empty((B, self.config.embed_dim), dtype=torch.float32, device=ids.device)
            check(lib().absb_enc_forward_dev(self._h, B, S, ptr(ids), ptr(mask), int(normalize_embeddings), ptr(out),
                                             _lib.current_stream_ptr()))
            return out
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        assert ids.ndim == 2, "input_ids must be [B, S]"
        B, S = ids.shape
        mask = (np.ones((B, S), dtype=np.int32) if attention_mask is None
                else np.ascontiguousarray(attention_mask, dtype=np.int32))
        assert mask.shape == ids.shape
        out = np.empty((B, self.config.embed_dim), dtype=np.float32)
        check(lib().absb_enc_forward(self._h, B, S, ptr(ids), ptr(mask), int(normalize_embeddings), ptr(out)))
        return out

    def encode(self, sentences, prompt_name: str | None = None, prompt: str | None = None, batch_size: int = 32,
               show_progress_bar: bool = False, output_value: str = "sentence_embedding",
               precision: str = "float32", convert_to_numpy: bool = True, convert_to_tensor: bool = False,
               device=None, normalize_embeddings: bool = False, **kwargs):
        if output_value != "sentence_embedding":
            raise ValueError("only output_value='sentence_embedding' is on the abstracts-search path")
        if precision != "float32":
            raise ValueError("only precision='float32' is on the abstracts-search path")
        single = isinstance(sentences, str)
        if single:
            sentences = [sentences]
        sentences = list(sentences)
        if prompt is None:
            name = prompt_name if prompt_name is not None else self.default_prompt_name
            if name is not None:
                if name not in self.prompts:
                    raise ValueError(f"Prompt name '{name}' not found in the configured prompts dictionary with keys "
                                     f"{list(self.prompts.keys())!r}.")
                prompt = self.prompts[name]
        if prompt:
            sentences = [prompt + s for s in sentences]
        n = len(sentences)
        out = np.empty((n, self.config.embed_dim), dtype=np.float32)
        order = np.argsort([-len(s) for s in sentences], kind="stable")
        for b0 in range(0, n, batch_size):
            idx = order[b0:b0 + batch_size]
            feats = self.tokenize([sentences[i] for i in idx])
            out[idx] = self.encode_tokens(feats["input_ids"], feats["attention_mask"], normalize_embeddings)
        if convert_to_tensor:
            import torch

            res = torch.from_numpy(out).to(self.device)
            return res[0] if single else res
        return out[0] if single else out

    # ---- measurement -----------------------------------------------------------------------
    def last_hidden_state(self, B: int, S: int) -> np.ndarray:
        out = np.empty((B, S, self.config.hidden_size), dtype=np.float32)
        check(lib().absb_enc_last_hidden(self._h, ptr(out), out.size))
        return out

    def last_stats(self) -> dict:
        f, l = c_double(), c_int64()
        check(lib().absb_enc_last_stats(self._h, byref(f), byref(l)))
        return {"flops": f.value, "launches": l.value}

    def set_attention_impl(self, impl: int):
        """1 = tcgen05 attention (default; every S <= 512), 0 = the mma.sync kernel (test hook)."""
        check(lib().absb_enc_set_attention_impl(self._h, int(impl)))

    def set_profile(self, on: int):
        check(lib().absb_enc_set_profile(self._h, int(on)))

    def get_profile(self) -> dict:
        g, gf, a, o, n = c_double(), c_double(), c_double(), c_double(), c_int64()
        check(lib().absb_enc_get_profile(self._h, byref(g), byref(gf), byref(a), byref(o), byref(n)))
        return {"gemm_ms": g.value, "gemm_flops": gf.value, "attention_ms": a.value, "other_ms": o.value,
                "forwards": n.value}


SentenceTransformer = Encoder


def gemm_set_variant(variant: int = 0):
    """Test hook: 0 auto, 1 = one CTA 128x256 tiles, 2 = CTA pair 256x256, 3 = CTA pair 256x192, 4 / 5 = quad (two
    pairs sharing the B tile by TMA multicast) with 256- / 192-column tiles."""
    check(lib().absb_gemm_set_variant(int(variant)))


def gemm_set_ksplit(slices: int = 0):
    """Split K of the residual-add GEMMs: 0 / 1 = never (default; measured slower on B200), n = n slices (test hook)."""
    check(lib().absb_gemm_set_ksplit(int(slices)))


def gemm_bf16_epi(A, B, epi: int, out=None, bias=None):
    """The tcgen05 GEMM with a fused epilogue (0 bf16, 1 f32, 2 f32 +=, 3 SwiGLU bf16) — test hook."""
    import torch

    A, B = A.contiguous(), B.contiguous()
    M, K = A.shape
    N = B.shape[0]
    if out is None:
        out = torch.empty((M, N // 2 if epi == 3 else N), dtype=torch.bfloat16 if epi in (0, 3) else torch.float32,
                          device=A.device)
    check(lib().absb_gemm_bf16_epi_dev(A.device.index or 0, epi, M, N, K, ptr(A), ptr(B), ptr(out), out.shape[1],
                                       ptr(bias), _lib.current_stream_ptr()))
    return out


def gemm_bf16(A, B):
    """C[M,N] fp32 = A[M,K] @ B[N,K]^T on tcgen05 (CUDA bf16 tensors) — test / micro-benchmark hook."""
    import torch

    assert A.is_cuda and B.is_cuda and A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    A, B = A.contiguous(), B.contiguous()
    M, K = A.shape
    N, K2 = B.shape
    assert K == K2
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(lib().absb_gemm_bf16_dev(A.device.index or 0, M, N, K, ptr(A), ptr(B), ptr(C), _lib.current_stream_ptr()))
    return C
