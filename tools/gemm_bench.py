"""Micro-benchmark of the tcgen05 GEMM on the encoder's shapes against cuBLAS (torch.matmul), every
tile variant.  Run under gpurun: python tools/gemm_bench.py [tokens]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
enc = importlib.import_module("abstracts-search_b200.encoder")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    shapes = [("qkv", T, 2048, 1536), ("o", T, 1536, 1536), ("ffn_in", T, 17920, 1536), ("ffn_down", T, 1536, 8960)]
    for name, M, N, K in shapes:
        A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
        B = torch.randn((N, K), device="cuda").to(torch.bfloat16)
        fl = 2.0 * M * N * K
        ms = timeit(lambda: torch.matmul(A, B.T))
        row = [f"{name:9s} M={M} N={N} K={K}  cuBLAS {ms*1e3:7.1f} us {fl/ms/1e9:7.0f} TF"]
        for v in (1, 2, 3, 4, 5):
            enc.gemm_set_variant(v)
            try:
                ms = timeit(lambda: enc.gemm_bf16(A, B))
                row.append(f"v{v} {ms*1e3:7.1f} us {fl/ms/1e9:7.0f} TF")
            except Exception as e:  # noqa: BLE001
                row.append(f"v{v} failed: {e}")
        enc.gemm_set_variant(0)
        print(" | ".join(row), flush=True)


def epilogues():
    """The four fused epilogues on their encoder shapes, as they run inside a layer."""
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    for name, epi, N, K in [("qkv bf16", 0, 2048, 1536), ("o  f32+=", 2, 1536, 1536), ("ffn swiglu", 3, 17920, 1536),
                            ("down f32+=", 2, 1536, 8960)]:
        A = torch.randn((T, K), device="cuda").to(torch.bfloat16)
        B = (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16)
        out = torch.zeros((T, N // 2 if epi == 3 else N), dtype=torch.bfloat16 if epi in (0, 3) else torch.float32, device="cuda")
        ms = timeit(lambda: enc.gemm_bf16_epi(A, B, epi, out=out))
        row = f"epi {name:11s} M={T} N={N} K={K}: auto {ms*1e3:7.1f} us {2.0*T*N*K/ms/1e9:7.0f} TF"
        if True:  # every epilogue under the pair (2, 3) and quad (4, 5) variants; 192-column tiles only where allowed
            for v in ((2, 3, 4, 5) if epi == 2 else (2, 4)):
                enc.gemm_set_variant(v)
                ms = timeit(lambda: enc.gemm_bf16_epi(A, B, epi, out=out))
                row += f" | v{v} {ms*1e3:7.1f} us {2.0*T*N*K/ms/1e9:7.0f} TF"
            enc.gemm_set_variant(0)
        print(row, flush=True)


def split_k():
    """Residual-add GEMMs with K cut into slices (absb_gemm_set_ksplit): auto choice against 1 / 2 / 4 / 8 slices."""
    for T in (2048, 4096, 8192, 16384):
        for name, N, K in [("o  f32+=", 1536, 1536), ("down f32+=", 1536, 8960)]:
            A = torch.randn((T, K), device="cuda").to(torch.bfloat16)
            B = (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16)
            out = torch.zeros((T, N), dtype=torch.float32, device="cuda")
            row = f"split-k {name:11s} M={T} N={N} K={K}:"
            for v in (0, 2, 3):
                enc.gemm_set_variant(v)
                row += f" || v{v}"
                for ks in (0, 1, 2, 4, 8, 16):
                    enc.gemm_set_ksplit(ks)
                    ms = timeit(lambda: enc.gemm_bf16_epi(A, B, 2, out=out))
                    row += f" | {'auto' if ks == 0 else ks}: {ms*1e3:6.1f}"
            enc.gemm_set_variant(0)
            enc.gemm_set_ksplit(0)
            print(row, flush=True)


if __name__ == "__main__":
    if "--split-k" in sys.argv:
        split_k()
        sys.exit(0)
    if "--epi-only" not in sys.argv:
        main()
    epilogues()
