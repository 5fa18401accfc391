"""Generates the golden fixtures in this directory FROM THE ORACLE (oracle/ivf.py, oracle/encoder.py).

The reference (colonelwatch/abstracts-search) holds no tests, fixtures or golden vectors for the
encode/search path, and faiss / sentence-transformers are not installable offline (SURVEY.md §4,
§8c), so these vectors pin the oracle against drift and give the CUDA path fixed targets; they are
not outputs of the reference itself ("parity unpinned").

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ivf as oivf  # noqa: E402
from oracle import synth as osynth  # noqa: E402

SEED = 1234
QSEED = 4321


def lattice_case():
    d, nlist, n, nq, nprobe, k = 1024, 128, 8192, 64, 8, 10
    x = osynth.corpus(SEED, 0, n, d, nlist)
    c = osynth.centroids(SEED, nlist, d)
    q = osynth.queries(SEED, 0, nq, d, nlist, n)
    ix = oivf.IVFFlat(d, nlist)
    ix.set_centroids(c)
    ix.add(x)
    Dc, Ic = ix.coarse(q, nprobe)
    D, I = ix.search_preassigned(q, k, Ic)
    D2, I2 = ix.search_preassigned(q, k, Ic, impl="c")
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    fl = oivf.FlatIP(d)
    fl.add(x)
    Df, If = fl.search(q, k)
    np.savez_compressed(os.path.join(HERE, "ivf_lattice_d1024.npz"), d=d, nlist=nlist, n=n, nq=nq, nprobe=nprobe, k=k,
                        seed=SEED, assign=ix.assign(x).astype(np.int32), sizes=ix.list_sizes(), Dc=Dc, Ic=Ic, D=D, I=I,
                        Df=Df, If=If)


def gauss_case():
    d, nlist, n, nq, nprobe, k = 64, 32, 5000, 40, 4, 5
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ix = oivf.IVFFlat(d, nlist)
    ix.train(x)
    ix.add(x)
    Dc, Ic = ix.coarse(q, nprobe)
    D, I = ix.search_preassigned(q, k, Ic)
    cm, fm = ix.ambiguity(q, k, nprobe)
    np.savez_compressed(os.path.join(HERE, "ivf_gauss_d64.npz"), d=d, nlist=nlist, n=n, nq=nq, nprobe=nprobe, k=k,
                        x=x, q=q, centroids=ix.centroids, assign=ix.assign(x).astype(np.int32), Dc=Dc, Ic=Ic, D=D, I=I,
                        coarse_margin=cm, fine_margin=fm)


def encoder_case():
    from oracle import encoder as oenc

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_cfg import TINY

    sd = oenc.random_state_dict(TINY, seed=0, std=0.05)
    rng = np.random.default_rng(11)
    ids = rng.integers(0, TINY.vocab_size, (5, 40)).astype(np.int64)
    mask = np.ones((5, 40), dtype=np.int64)
    mask[1, 33:] = 0
    mask[2, 7:] = 0
    mask[4, 1:] = 0
    emb, hidden = oenc.forward_plain(TINY, sd, ids, mask, normalize=True, return_hidden=True)
    np.savez_compressed(os.path.join(HERE, "encoder_tiny.npz"), ids=ids, mask=mask, emb=emb,
                        hidden_l2=np.linalg.norm(hidden, axis=-1).astype(np.float32))


if __name__ == "__main__":
    lattice_case()
    gauss_case()
    encoder_case()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
