"""One fused-epilogue GEMM launched a few times (for `ncu -k regex:gemm_bf16_tc_kernel -s 3 -c 1`):
python tools/gemm_one.py EPI M N K"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
enc = importlib.import_module("abstracts-search_b200.encoder")

epi, M, N, K = (int(a) for a in sys.argv[1:5])
A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
B = (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16)
out = torch.zeros((M, N // 2 if epi == 3 else N), dtype=torch.bfloat16 if epi in (0, 3) else torch.float32, device="cuda")
for _ in range(6):
    enc.gemm_bf16_epi(A, B, epi, out=out)
torch.cuda.synchronize()
