"""How sensitive are the residual-add GEMMs to the depth of the operand ring?  Shrinks the shared-memory budget
(absb_gemm_set_smem_budget) so that the kernels run with fewer stages and times O-proj / FFN-down / QKV.
python tools/gemm_stages.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
enc = importlib.import_module("abstracts-search_b200.encoder")
from importlib import import_module

lib = import_module("abstracts-search_b200._lib")


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


for T in (16384, 2048):
    for name, epi, N, K in [("o  f32+=", 2, 1536, 1536), ("down f32+=", 2, 1536, 8960), ("qkv bf16", 0, 2048, 1536)]:
        A = torch.randn((T, K), device="cuda").to(torch.bfloat16)
        B = (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16)
        out = torch.zeros((T, N), dtype=torch.bfloat16 if epi == 0 else torch.float32, device="cuda")
        for v in (2, 3):
            if epi == 0 and v == 3:
                continue
            enc.gemm_set_variant(v)
            row = f"{name:11s} M={T} v{v}:"
            for kb in (227, 196, 164, 132, 227):
                lib.check(lib.lib().absb_gemm_set_smem_budget(kb * 1024))
                ms = timeit(lambda: enc.gemm_bf16_epi(A, B, epi, out=out))
                row += f" | {kb} KB {ms*1e3:6.1f} us"
            lib.check(lib.lib().absb_gemm_set_smem_budget(0))
            print(row, flush=True)
        enc.gemm_set_variant(0)
