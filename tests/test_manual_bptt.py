"""The hand-derived BPTT (oracle/manual_bptt.py — the derivation the CUDA host code mirrors) against
autograd over the oracle forward, in fp64, with dropout masks replayed on both sides.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import manual_bptt as MB
from oracle import xgating_oracle as O
from tests.common import CONFIGS, fro_err, make_case, rel_err


def make_masks(cfg, batch, p, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    B, K, H = cfg["B"], cfg["K"], cfg["dims"]["H"]
    L = batch["seq"].shape[1]
    def mk(*shape):
        return (torch.rand(*shape, generator=g) >= p).to(dtype) / (1 - p)
    return {"enc_emb_rgb": mk(B, K, H), "enc_emb_opfl": mk(B, K, H), "enc_gate_rgb": mk(B, K, H),
            "enc_gate_opfl": mk(B, K, H), "enc_fusion": mk(B, K, H), "dec_gate": mk(L, B, H),
            "dec_h1": mk(L, B, H), "dec_h2": mk(L, B, H), "cls": mk(L, B, 128)}


@pytest.mark.parametrize("name", ["tiny", "mid"])
@pytest.mark.parametrize("train,p", [(True, 0.5), (True, 0.0), (False, 0.0)])
@pytest.mark.parametrize("activity", ["ReLU", "Tanh", "Sigmoid"])
def test_manual_matches_autograd(name, train, p, activity):
    dt = torch.float64
    cfg, P, b = make_case(name, dt)
    masks = make_masks(cfg, b, p, 5, dt) if p > 0 else None
    w = 0.5
    cls = b["seq"] % cfg["dims"]["C"]
    names = [n for n in P if not n.endswith(("running_mean", "running_var"))]
    Q = {n: (P[n].clone().requires_grad_(True) if n in names else P[n]) for n in P}
    logp, cat = O.forward(Q, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"],
                          train=train, masks=masks, activity=activity)
    Lp = logp.shape[1]
    loss = (O.language_model_criterion(logp, b["seq"][:, :Lp], b["seq_mask"][:, :Lp])
            + w * O.classifer_criterion(cat, cls[:, :Lp], b["seq_mask"][:, :Lp]))
    logp.retain_grad(); cat.retain_grad()
    loss.backward()

    with torch.no_grad():
        logp_m, cat_m, S = MB.forward_saved(P, b, train, masks, activity)
        assert S["Lp"] == Lp
        assert rel_err(logp_m.numpy(), logp.detach().numpy()) < 1e-10
        assert rel_err(cat_m.numpy(), cat.detach().numpy()) < 1e-10
        Gd = MB.backward(P, b, S, logp.grad, cat.grad)
    for n in names:
        ref = Q[n].grad if Q[n].grad is not None else torch.zeros_like(Q[n])
        if float(ref.norm()) < 1e-12:
            assert float(Gd[n].norm()) < 1e-10, n
        else:
            assert fro_err(Gd[n].numpy(), ref.numpy()) < 1e-9, n
