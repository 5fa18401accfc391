"""Context number (SURVEY §8d "reference GPU path"): what the reference's own software stack would
do for the encoder on this B200 — stock `transformers.Qwen2Model` of the stella_en_1.5B_v5 shape in
bf16 with PyTorch SDPA (cuBLAS GEMMs + torch's fused attention), bidirectional attention, mean pool,
Dense 1536->1024, L2 normalise — timed with CUDA events on the same two shapes bench.py uses.
sentence-transformers itself is not installable offline; this is its hot loop without the tokenizer.
Independent of oracle/ (tools may not import it).  Run under gpurun:
    python tools/hf_gpu_encoder_bench.py            # prints one JSON line per shape
    python tools/hf_gpu_encoder_bench.py --tiny-cpu # CPU self-check of the non-causal switch
"""
import json
import sys

import torch


def build(device, dtype, tiny=False):
    from transformers import Qwen2Config, Qwen2Model

    if tiny:
        kw = dict(vocab_size=1000, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                  num_key_value_heads=1)
    else:
        kw = dict(vocab_size=151646, hidden_size=1536, intermediate_size=8960, num_hidden_layers=28,
                  num_attention_heads=12, num_key_value_heads=2)
    cfg = Qwen2Config(max_position_embeddings=512, rms_norm_eps=1e-6, rope_theta=1e6, use_sliding_window=False,
                      attention_dropout=0.0, tie_word_embeddings=False, **kw)
    cfg.head_dim = 128
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    with torch.device(device):
        model = Qwen2Model(cfg).to(dtype).eval()
        dense = torch.nn.Linear(kw["hidden_size"], 1024 if not tiny else 128, bias=True).to(dtype)
    return model, dense


def set_causal(model, causal: bool):
    for layer in model.layers:
        layer.self_attn.is_causal = causal


@torch.no_grad()
def encode(model, dense, ids):
    B, S = ids.shape
    # a dict mask makes Qwen2Model skip its own causal-mask construction; with no padding the
    # bidirectional mask is "no mask", and is_causal=False on the attention modules keeps SDPA non-causal
    out = model(input_ids=ids, attention_mask={"full_attention": None},
                position_ids=torch.arange(S, device=ids.device)[None, :].expand(B, S))
    pooled = out.last_hidden_state.mean(1)
    return torch.nn.functional.normalize(dense(pooled).float(), p=2, dim=1)


def main():
    if "--tiny-cpu" in sys.argv:
        model, dense = build("cpu", torch.float32, tiny=True)
        ids = torch.randint(0, 1000, (2, 16))
        set_causal(model, False)
        a = encode(model, dense, ids)
        set_causal(model, True)
        b = encode(model, dense, ids)
        print("non-causal vs causal max diff:", float((a - b).abs().max()))
        assert (a - b).abs().max() > 1e-4, "the is_causal switch had no effect"
        return
    model, dense = build("cuda", torch.bfloat16)
    set_causal(model, False)
    for B, S, what in ((32, 256, "bulk encode b=32 x 256 tokens (BASELINE configs[1])"),
                       (512, 32, "query encode 512 x 32 tokens (one bench step at N=1)"),
                       (64, 32, "query encode 64 x 32 tokens (one bench step per GPU at N=8)")):
        ids = torch.randint(0, 151646, (B, S), device="cuda")
        for _ in range(3):
            encode(model, dense, ids)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        e0.record()
        for _ in range(steps):
            encode(model, dense, ids)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"impl": "transformers.Qwen2Model bf16 + SDPA (stock torch/cuBLAS path)", "workload": what,
                          "ms_per_step": ms, "embeddings_per_s": B / (ms * 1e-3), "tokens_per_step": B * S,
                          "torch": torch.__version__}), flush=True)


if __name__ == "__main__":
    main()
