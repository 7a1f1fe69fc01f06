"""Optimizer step: the numpy oracle is pinned to torch.optim.Adam; the fused CUDA kernel is
checked against the oracle (both eps conventions), including the elementwise clamp of myutils.clip_gradient."""
import numpy as np
import pytest
import torch

from oracle import adam_oracle as AO


def _case(seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    shapes = [(7,), (3, 5), (1,), (130, 33), (4100,)]
    P = [rng.standard_normal(s).astype(dtype) for s in shapes]
    G = [[(rng.standard_normal(s) * 0.3).astype(dtype) for s in shapes] for _ in range(4)]
    return P, G


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_oracle_matches_torch_adam(wd):
    P, G = _case(1)
    tp = [torch.nn.Parameter(torch.tensor(p)) for p in P]
    opt = torch.optim.Adam(tp, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    p, m, v = [x.copy() for x in P], [np.zeros_like(x) for x in P], [np.zeros_like(x) for x in P]
    for step, g in enumerate(G, 1):
        for t, gi in zip(tp, g):
            t.grad = torch.tensor(gi)
        opt.step()
        p, m, v = AO.adam_step(p, g, m, v, step, lr=4e-4, weight_decay=wd, eps_mode=1)
    for a, b in zip(p, tp):
        assert np.allclose(a, b.detach().numpy(), rtol=1e-12, atol=1e-14)


def test_oracle_modes_agree_without_eps_and_clip_is_clamp():
    P, G = _case(2)
    z = [np.zeros_like(x) for x in P]
    a = AO.adam_step(P, G[0], z, z, 3, eps=0.0, eps_mode=0)[0]
    b = AO.adam_step(P, G[0], z, z, 3, eps=0.0, eps_mode=1)[0]
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-12)
    c = AO.clip_gradient(G[0], 0.1)
    assert all(float(np.abs(x).max()) <= 0.1 for x in c)
    assert all(np.array_equal(x[np.abs(g) < 0.1], g[np.abs(g) < 0.1]) for x, g in zip(c, G[0]))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("wd,clip", [(0.0, 0.1), (0.01, 0.0)])
def test_fused_adam_matches_oracle(mode, wd, clip):
    from controllable_xgating_b200.optim import FusedAdam
    P, G = _case(3, np.float32)
    tp = [torch.nn.Parameter(torch.tensor(p).cuda()) for p in P]
    opt = FusedAdam(tp, lr=4e-4, weight_decay=wd, grad_clip=clip, reference_eps=(mode == 0))
    p = [x.astype(np.float64) for x in P]
    m, v = [np.zeros_like(x) for x in p], [np.zeros_like(x) for x in p]
    for step, g in enumerate(G, 1):
        for t, gi in zip(tp, g):
            t.grad = torch.tensor(gi).cuda()
        opt.step()
        p, m, v = AO.adam_step(p, [x.astype(np.float64) for x in g], m, v, step, lr=4e-4, weight_decay=wd, grad_clip=clip,
                               eps_mode=mode)
        if clip > 0:      # clip_gradient clamps .grad in place
            assert all(float(t.grad.abs().max()) <= float(np.float32(clip)) for t in tp)
    for a, b in zip(p, tp):
        assert np.allclose(a, b.detach().cpu().numpy().astype(np.float64), rtol=2e-5, atol=2e-7)
    assert all(opt.state[t]["step"] == len(G) for t in tp)


@pytest.mark.gpu
def test_fused_adam_on_the_model_invalidates_derived_tables():
    """a training step + FusedAdam.step() + greedy decode: the decoder must see the updated weights
    (POS-gate token table and tf32 splits are rebuilt), i.e. equal a freshly constructed model."""
    import controllable_xgating_b200 as X
    from controllable_xgating_b200.optim import FusedAdam
    from tests.test_gpu_parity import _full_case, build_model, dev
    cfg, P, b = _full_case(8, seed=9); d = dev(b)
    m = build_model(cfg, P, drop=0.0)
    opt = FusedAdam(m.parameters(), lr=1e-2, grad_clip=0.1)
    o = {"sample_max": 1, "beam_size": 1}
    m.eval(); m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], o)      # builds the derived tables
    m.train()
    logp, _ = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    X.LanguageModelCriterion()(logp, d["seq"], d["seq_mask"]).backward()
    opt.step()
    m.eval(); seq1, lp1 = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], o)
    m2 = build_model(cfg, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}, drop=0.0).eval()
    seq2, lp2 = m2.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], o)
    assert torch.equal(seq1, seq2) and torch.equal(lp1, lp2)
