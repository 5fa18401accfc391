#!/usr/bin/env python
"""Index.add rate alone (assign GEMM with fused arg-max + append): 1M unit-norm rows per step into IVF65536."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
P = importlib.import_module("abstracts-search_b200")
d, nlist, n = 1024, 65536, 1 << 20
ix = P.index_factory(d, f"IVF{nlist},Flat", P.METRIC_INNER_PRODUCT)
ix.set_centroids(P.synth.centroids(1234, nlist, d))
xb = [P.synth.corpus(1234, s * n, n, d, nlist, unit=True) for s in range(4)]
ix.add(xb[0])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(1, 4):
    ix.add(xb[s])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"ABSB_PDL={os.environ.get('ABSB_PDL','1')} add: {ms:.1f} ms per 1M rows = {n/ms/1e3:.3f} M rows/s, assign {n*6*2*nlist*d/ms/1e9:.0f} TFLOP/s", flush=True)
# the assign GEMM alone
ix.set_profile(2)
x = xb[1]
for _ in range(2):
    ix.assign(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ix.assign(x)
torch.cuda.synchronize()
print(f"  assign alone: {(time.perf_counter()-t0)/3*1e3:.1f} ms")
# add_core alone (precomputed lists): sort + page table + scatter + allocation of new slabs
lists = P.synth.cluster_of(1234, 4 * n, n, nlist)
x5 = P.synth.corpus(1234, 4 * n, n, d, nlist, unit=True)
ids = torch.arange(4 * n, 5 * n, device="cuda")
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    ix.add_core(x5, ids, lists)
    torch.cuda.synchronize()
    print(f"  add_core alone: {(time.perf_counter()-t0)*1e3:.1f} ms (ntotal {ix.ntotal})", flush=True)
t0 = time.perf_counter()
ix.add(x5)
torch.cuda.synchronize()
print(f"  add (assign + add_core): {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
