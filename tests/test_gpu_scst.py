"""sample() under model.train() — the self-critical path (starttrain.py:131, SAModel.py:163-219,
255-267): tokens drawn with dropout + batch-statistics BatchNorm active, returned log-probs carry gradients."""
import numpy as np
import pytest
import torch

from oracle import xgating_oracle as O
from tests.common import RTOL, make_case, rel_err
from tests.test_gpu_parity import _grad_check, build_model, dev

pytestmark = pytest.mark.gpu


def _masks(seed, p, B, K, H, L):
    from controllable_xgating_b200.engine import dropout_mask

    def mk(site, shape):
        return dropout_mask(seed, site, int(np.prod(shape)), p, "cuda").cpu().view(*shape)
    return {"enc_emb_rgb": mk("enc_emb_rgb", (B, K, H)), "enc_emb_opfl": mk("enc_emb_opfl", (B, K, H)),
            "enc_gate_rgb": mk("enc_gate_rgb", (K, B, H)).transpose(0, 1).contiguous(),
            "enc_gate_opfl": mk("enc_gate_opfl", (K, B, H)).transpose(0, 1).contiguous(),
            "enc_fusion": mk("enc_fusion", (K, B, H)).transpose(0, 1).contiguous(),
            "dec_gate": mk("dec_gate", (L, B, H)), "dec_h1": mk("dec_h1", (L, B, H)), "dec_h2": mk("dec_h2", (L, B, H)),
            "cls": mk("cls", (L, B, 128))}


@pytest.mark.parametrize("name,sample_max", [("mid", 0), ("mid", 1), ("tiny", 0)])
def test_training_mode_sample_matches_oracle_and_backpropagates(name, sample_max):
    import controllable_xgating_b200 as X
    cfg, P, b = make_case(name); d = dev(b)
    dims, B, K = cfg["dims"], cfg["B"], cfg["K"]
    p = 0.5
    m = build_model(cfg, P, drop=p).train()
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # the dropout seed Engine.next_seed() will draw first
    torch.manual_seed(77)
    seq, lp = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": sample_max, "beam_size": 1})
    assert lp.requires_grad and tuple(lp.shape) == tuple(seq.shape)
    steps = seq.shape[1]
    # (1) the teacher-forced replay reproduces the log-probs of the sampling pass itself (same masks, same BatchNorm)
    assert rel_err(lp.detach().cpu().numpy(), m._last_sample_logprobs.cpu().numpy()) < 1e-4
    # (2) oracle: training-mode forward on the sampled tokens with the kernels' dropout masks replayed
    seq_c = seq.cpu()
    seq_in = torch.cat([torch.zeros(B, 1, dtype=torch.long), seq_c[:, :steps - 1]], 1)
    mask = torch.cat([torch.ones(B, 1), (seq_c[:, :steps - 1] > 0).float()], 1)
    masks = _masks(seed, p, B, K, dims["H"], steps)
    Pg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in P.items()}
    logp_o, _ = O.forward(Pg, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], seq_in, mask, train=True, masks=masks)
    lp_o = logp_o[:, :steps].gather(2, seq_c.unsqueeze(2)).squeeze(2)
    live = torch.cat([torch.ones(B, 1, dtype=torch.bool), seq_c[:, :steps - 1] > 0], 1)   # positions RewardCriterion keeps
    assert rel_err((lp.detach().cpu() * live).numpy(), (lp_o.detach() * live).numpy()) < RTOL
    # (3) RewardCriterion (SAModel.py:255-267) and its gradients
    g = torch.Generator().manual_seed(5)
    reward = torch.rand(B, steps, generator=g) - 0.3
    loss = X.RewardCriterion()(lp, seq, reward.cuda())
    loss.backward()
    loss_o = O.reward_criterion(lp_o, seq_c, reward)
    assert abs(float(loss) - float(loss_o)) < 1e-4 * max(abs(float(loss_o)), 1e-3)
    names = [k for k, v in Pg.items() if v.requires_grad]
    grads = torch.autograd.grad(loss_o, [Pg[k] for k in names], allow_unused=True)
    _grad_check(m, {k: (gr.numpy() if gr is not None else np.zeros(tuple(Pg[k].shape), np.float32)) for k, gr in zip(names, grads)})


def test_training_mode_sample_updates_batchnorm_once():
    cfg, P, b = make_case("mid"); d = dev(b)
    m1 = build_model(cfg, P, drop=0.5).train()
    m2 = build_model(cfg, P, drop=0.5).train()
    m1.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 0, "beam_size": 1})
    m2(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    for (k1, v1), (k2, v2) in zip(m1.named_buffers(), m2.named_buffers()):
        assert torch.equal(v1, v2), k1


@pytest.mark.parametrize("name,ss", [("mid", 0.5), ("tiny", 1.0), ("mid", 0.25)])
def test_scheduled_sampling_forward_matches_oracle_on_the_tokens_used(name, ss):
    """ss_prob > 0 (SAModel.py:89-99): the token pass decides the inputs, the replayed forward must
    equal the oracle's training forward on exactly those tokens (dropout masks replayed), gradients included."""
    import controllable_xgating_b200 as X
    cfg, P, b = make_case(name); d = dev(b)
    dims, B, K = cfg["dims"], cfg["B"], cfg["K"]
    p = 0.5
    m = build_model(cfg, P, drop=p).train()
    m.ss_prob = ss
    torch.manual_seed(99)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    torch.manual_seed(99)
    logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    Lp = logp.shape[1]
    used = m._last_ss_tokens.cpu()
    gt = b["seq"]
    assert torch.equal(used[:, 0], gt[:, 0])                       # step 0 is never sampled
    changed = (used[:, 1:Lp] != gt[:, 1:Lp]).float().mean().item()
    if ss == 1.0:
        assert changed > 0.8                                       # every input drawn from the model (vocab >= 40)
    else:
        assert 0.0 < changed <= ss + 0.25                          # B*L Bernoulli draws around ss_prob
    L = gt.shape[1]
    masks = _masks(seed, p, B, K, dims["H"], L)
    Pg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in P.items()}
    logp_o, cat_o = O.forward(Pg, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], used[:, :Lp], b["seq_mask"][:, :Lp], train=True,
                              masks=masks)
    assert rel_err(logp.detach().cpu().numpy(), logp_o.detach().numpy()[:, :Lp]) < RTOL
    loss = X.LanguageModelCriterion()(logp, d["seq"][:, :Lp], d["seq_mask"][:, :Lp])     # targets stay the ground truth
    loss.backward()
    loss_o = O.language_model_criterion(logp_o[:, :Lp], gt[:, :Lp], b["seq_mask"][:, :Lp])
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    names = [k for k, v in Pg.items() if v.requires_grad]
    grads = torch.autograd.grad(loss_o, [Pg[k] for k in names], allow_unused=True)
    _grad_check(m, {k: (gr.numpy() if gr is not None else np.zeros(tuple(Pg[k].shape), np.float32)) for k, gr in zip(names, grads)})


@pytest.mark.parametrize("name,sample_max", [("mid", 0), ("mid", 1), ("c1", 0)])
def test_training_mode_sampling_loop_runs_on_the_grouped_kernel(name, sample_max):
    """The sampling word loop (multinomial draw and / or training dropout, SAModel.py:188-196 under model.train()) runs as
    one launch of decode_grouped_kernel<1> (asserted: strict handle + launch list) and draws the tokens the per-step
    launches draw from the same Philox streams (a draw that lands within rounding of a CDF step may differ)."""
    from tests.common import fused_path
    cfg, P, b = make_case(name); d = dev(b)
    res = []
    for persistent in (True, False):
        m = build_model(cfg, P, drop=0.5).train()
        m._engine.set_engine(True, persistent)
        torch.manual_seed(123)
        with torch.no_grad():
            if persistent:
                with fused_path(m, ["encode_persistent", "decode_persistent"]):
                    seq, lp = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": sample_max, "beam_size": 1})
            else:
                seq, lp = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": sample_max, "beam_size": 1})
        res.append((seq.cpu(), lp.cpu()))
    (s0, l0), (s1, l1) = res
    n = min(s0.shape[1], s1.shape[1])
    same_rows = (s0[:, :n] == s1[:, :n]).all(dim=1)
    assert same_rows.float().mean().item() >= 0.9, (s0, s1)
    assert rel_err(l0[:, :n][same_rows].numpy(), l1[:, :n][same_rows].numpy()) < 1e-3


def test_multinomial_sampling_distribution_on_the_grouped_kernel():
    """sample_max=0 in eval mode on decode_grouped_kernel<1>: the first token of 64 x 64 draws follows exp(logp / T)."""
    from tests.common import fused_path
    cfg, P, b = make_case("c1")
    B, reps, T = 64, 64, 0.8
    m = build_model(cfg, P).eval()
    one = {k: (v[:1].repeat(B, *([1] * (v.dim() - 1))) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    d = dev(one)
    first, lps0 = [], []
    with fused_path(m, ["encode_persistent", "decode_persistent"]):
        for i in range(reps):
            torch.manual_seed(1000 + i)
            seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 0, "temperature": T})
            first.append(seq[:, 0].cpu()); lps0.append(lps[:, 0].cpu())
    first = torch.cat(first).numpy(); lps0 = torch.cat(lps0).numpy()
    with torch.no_grad():
        V = O.encoder_fwd(P, b["rgb"][:1], b["opfl"][:1], b["feat_mask"][:1])
        st = O.init_hidden(P, V, b["feat_mask"][:1])
        lp0, _ = O.get_logprobs_state(P, torch.zeros(1, dtype=torch.long), V, b["pos"][:1], st)
    prob = torch.softmax(lp0[0] / T, 0).numpy()
    freq = np.bincount(first, minlength=prob.size) / (B * reps)
    assert np.abs(freq - prob).max() < 4 * np.sqrt(prob.max() / (B * reps)) + 5e-3
    nz = first > 0
    assert np.allclose(lps0[nz], lp0[0][torch.from_numpy(first)].numpy()[nz], atol=1e-4)      # log-probs at temperature 1


@pytest.mark.parametrize("name,ss", [("mid", 0.5), ("c1", 1.0)])
def test_scheduled_sampling_token_pass_runs_on_the_grouped_kernel(name, ss):
    """ss_prob > 0: the token pass (SAModel.py:89-99) runs as one launch of decode_grouped_kernel<1> (asserted) and feeds
    the tokens the per-step launches feed (same Philox streams; a draw within rounding of a CDF step may differ)."""
    from tests.common import fused_path
    cfg, P, b = make_case(name); d = dev(b)
    used = []
    for persistent in (True, False):
        m = build_model(cfg, P, drop=0.5).train()
        m._engine.set_engine(True, persistent)
        m.ss_prob = ss
        torch.manual_seed(321)
        with torch.no_grad():
            if persistent:
                with fused_path(m, ["encode_persistent", "decode_persistent", "train_decode_persistent"]):
                    m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
            else:
                m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        used.append(m._last_ss_tokens.cpu())
    same_rows = (used[0] == used[1]).all(dim=1)
    assert same_rows.float().mean().item() >= 0.9, (used[0], used[1])
    assert (used[0] != b["seq"][:, :used[0].shape[1]]).any()          # something was sampled
