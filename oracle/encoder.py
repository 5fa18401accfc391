"""CPU oracle for the encoder half of the path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the arithmetic lives in sentence-transformers + stella's remote modeling_qwen.py
under transformers<=4.49.0 (/root/reference/requirements.txt:1-3), none of which is vendored or
installable offline, and the reference holds no tests or golden vectors; the
real weights cannot be fetched either.  This module restates the computation at the reference's call
sites — `SentenceTransformer.encode` in `sidecar-search build` (/root/reference/Makefile:65) and
app.py (/root/reference/README.md:28):

    Qwen2 backbone (installed transformers' Qwen2Model, fp32, eager attention) run with a
    bidirectional padding-only mask  ->  sentence-transformers Pooling(mean)  ->
    Dense(1536 -> 1024, bias, identity)  ->  optional F.normalize(p=2, dim=1)

Two statements check each other: the transformers module (`TransformersOracle`) and a plain
torch restatement written from the model definition (`forward_plain`; modeling_qwen2.py: MLP :35-48,
RoPE :51-146, eager attention :161-183, attention block :187-245, RMSNorm :249-263, layer :269-309).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def rope_tables(S: int, head_dim: int, theta: float):
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    pos = torch.arange(S, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def forward_plain(cfg, sd: dict, input_ids: np.ndarray, attention_mask: np.ndarray, normalize: bool = False,
                  return_hidden: bool = False):
    """Plain fp32 torch restatement.  cfg: EncoderConfig-like; sd: name -> float32 ndarray."""
    W = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in sd.items()}
    ids = torch.from_numpy(np.asarray(input_ids, dtype=np.int64))
    mask = torch.from_numpy(np.asarray(attention_mask)).to(torch.float32)
    B, S = ids.shape
    nh, nkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    h = W["embed_tokens.weight"][ids]
    cos, sin = rope_tables(S, hd, cfg.rope_theta)
    bias = torch.zeros(B, 1, S, S)
    bias = bias.masked_fill(mask[:, None, None, :] == 0, torch.finfo(torch.float32).min)
    if cfg.causal:
        bias = bias + torch.full((S, S), torch.finfo(torch.float32).min).triu(1)[None, None]

    def rms(x, w):
        v = x.pow(2).mean(-1, keepdim=True)
        return w * (x * torch.rsqrt(v + cfg.rms_eps))

    for l in range(cfg.num_layers):
        p = f"layers.{l}."
        x = rms(h, W[p + "input_layernorm.weight"])
        q = x @ W[p + "self_attn.q_proj.weight"].T + W[p + "self_attn.q_proj.bias"]
        k = x @ W[p + "self_attn.k_proj.weight"].T + W[p + "self_attn.k_proj.bias"]
        v = x @ W[p + "self_attn.v_proj.weight"].T + W[p + "self_attn.v_proj.bias"]
        q = q.view(B, S, nh, hd).transpose(1, 2)
        k = k.view(B, S, nkv, hd).transpose(1, 2)
        v = v.view(B, S, nkv, hd).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        k = k.repeat_interleave(nh // nkv, dim=1)
        v = v.repeat_interleave(nh // nkv, dim=1)
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd) + bias, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, nh * hd)
        h = h + a @ W[p + "self_attn.o_proj.weight"].T
        x = rms(h, W[p + "post_attention_layernorm.weight"])
        g = x @ W[p + "mlp.gate_proj.weight"].T
        u = x @ W[p + "mlp.up_proj.weight"].T
        h = h + (torch.nn.functional.silu(g) * u) @ W[p + "mlp.down_proj.weight"].T
    hidden = rms(h, W["norm.weight"])
    pooled = (hidden * mask[:, :, None]).sum(1) / mask.sum(1, keepdim=True).clamp(min=1e-9)
    emb = pooled @ W["dense.weight"].T + W["dense.bias"]
    if normalize:
        emb = torch.nn.functional.normalize(emb, p=2, dim=1)
    if return_hidden:
        return emb.numpy(), hidden.numpy()
    return emb.numpy()


class TransformersOracle:
    """transformers' own Qwen2Model (fp32, eager) with the ST pooling/dense/normalise restated."""

    def __init__(self, cfg, sd: dict):
        from transformers import Qwen2Config, Qwen2Model

        self.cfg = cfg
        hf = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                         intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers,
                         num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_kv_heads,
                         max_position_embeddings=max(cfg.max_seq_len, 512), rms_norm_eps=cfg.rms_eps,
                         rope_theta=cfg.rope_theta, use_sliding_window=False, attention_dropout=0.0,
                         tie_word_embeddings=False)
        hf.head_dim = cfg.head_dim
        if hasattr(hf, "rope_parameters"):
            try:
                hf.rope_parameters = {"rope_type": "default", "rope_theta": float(cfg.rope_theta)}
            except Exception:
                pass
        hf._attn_implementation = "eager"
        with torch.no_grad():
            self.model = Qwen2Model(hf).to(torch.float32).eval()
            msd = self.model.state_dict()
            for name, arr in sd.items():
                if name.startswith("dense."):
                    continue
                assert name in msd, name
                msd[name].copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)))
        self.dense_w = torch.from_numpy(np.ascontiguousarray(sd["dense.weight"], dtype=np.float32))
        self.dense_b = torch.from_numpy(np.ascontiguousarray(sd["dense.bias"], dtype=np.float32))

    @torch.no_grad()
    def forward(self, input_ids, attention_mask, normalize: bool = False, return_hidden: bool = False):
        ids = torch.from_numpy(np.asarray(input_ids, dtype=np.int64))
        mask = torch.from_numpy(np.asarray(attention_mask)).to(torch.float32)
        B, S = ids.shape
        if self.cfg.causal:
            out = self.model(input_ids=ids, attention_mask=mask.to(torch.int64))
        else:
            add = torch.zeros(B, 1, S, S).masked_fill(mask[:, None, None, :] == 0, torch.finfo(torch.float32).min)
            # a dict mask makes Qwen2Model skip its own causal-mask construction (modeling_qwen2.py:378-393)
            out = self.model(input_ids=ids, attention_mask={"full_attention": add},
                             position_ids=torch.arange(S)[None, :].expand(B, S))
        hidden = out.last_hidden_state
        pooled = (hidden * mask[:, :, None]).sum(1) / mask.sum(1, keepdim=True).clamp(min=1e-9)
        emb = pooled @ self.dense_w.T + self.dense_b
        if normalize:
            emb = torch.nn.functional.normalize(emb, p=2, dim=1)
        if return_hidden:
            return emb.numpy(), hidden.numpy()
        return emb.numpy()


def random_state_dict(cfg, seed: int = 0, std: float = 0.02) -> dict:
    """Seeded fp32 weights of the true shapes (bf16-representable so both sides hold identical values)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in cfg.param_shapes().items():
        if name.endswith("layernorm.weight") or name == "norm.weight":
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        sd[name] = t.to(torch.bfloat16).to(torch.float32).numpy()
    return sd


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return (a64 * b64).sum(1) / (np.linalg.norm(a64, axis=1) * np.linalg.norm(b64, axis=1))
