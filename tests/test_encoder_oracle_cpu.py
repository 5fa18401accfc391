"""CPU tests of the encoder oracle: plain restatement vs transformers' Qwen2Model, golden vector."""
import numpy as np

from conftest import golden
from oracle import encoder as oenc
from tiny_cfg import TINY, TinyCfg


def _inputs():
    g = golden("encoder_tiny.npz")
    return g, g["ids"], g["mask"]


def test_plain_oracle_matches_golden():
    g, ids, mask = _inputs()
    sd = oenc.random_state_dict(TINY, seed=0, std=0.05)
    emb, hidden = oenc.forward_plain(TINY, sd, ids, mask, normalize=True, return_hidden=True)
    assert np.allclose(emb, g["emb"], atol=2e-6)
    assert np.allclose(np.linalg.norm(hidden, axis=-1), g["hidden_l2"], rtol=1e-5)
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)


def test_plain_oracle_matches_transformers_bidirectional_and_causal():
    g, ids, mask = _inputs()
    for causal in (False, True):
        cfg = TinyCfg(causal=causal)
        sd = oenc.random_state_dict(cfg, seed=0, std=0.05)
        e1, h1 = oenc.forward_plain(cfg, sd, ids, mask, normalize=True, return_hidden=True)
        e2, h2 = oenc.TransformersOracle(cfg, sd).forward(ids, mask, normalize=True, return_hidden=True)
        m = mask.astype(bool)
        assert np.abs(h1 - h2)[m].max() < 2e-5
        assert np.abs(e1 - e2).max() < 2e-6
    # the two readings genuinely differ (SURVEY §8a note on a2 causality)
    sd = oenc.random_state_dict(TINY, seed=0, std=0.05)
    eb = oenc.forward_plain(TinyCfg(causal=False), sd, ids, mask, normalize=True)
    ec = oenc.forward_plain(TinyCfg(causal=True), sd, ids, mask, normalize=True)
    assert np.abs(eb - ec).max() > 1e-3


def test_padding_does_not_change_the_embedding():
    g, ids, mask = _inputs()
    sd = oenc.random_state_dict(TINY, seed=0, std=0.05)
    full = oenc.forward_plain(TINY, sd, ids, mask, normalize=True)
    n2 = int(mask[2].sum())
    alone = oenc.forward_plain(TINY, sd, ids[2:3, :n2], mask[2:3, :n2], normalize=True)
    assert np.abs(full[2] - alone[0]).max() < 2e-6
