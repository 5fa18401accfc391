"""Small encoder configuration shared by the CPU and GPU tests and the golden generator.  It keeps
the true architecture's structure (GQA, head_dim 128, SwiGLU, bias on q/k/v only) at a size whose
fp32 oracle runs in milliseconds.  Defined without importing the product package so that the
oracle-only tests never need libabsb200.so."""
from dataclasses import dataclass


@dataclass
class TinyCfg:
    vocab_size: int = 1000
    hidden_size: int = 256
    num_layers: int = 2
    num_heads: int = 2
    num_kv_heads: int = 1
    head_dim: int = 128
    intermediate_size: int = 512
    embed_dim: int = 128
    max_seq_len: int = 512
    causal: bool = False
    rms_eps: float = 1e-6
    rope_theta: float = 1e6

    def param_shapes(self) -> dict:
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


TINY = TinyCfg()
