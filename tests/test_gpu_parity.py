"""GPU parity tests: the CUDA path (through the Python mirror -> C ABI -> libxgating.so) against the
CPU oracle on the same seeded inputs and against the committed golden vectors produced by the real
reference.  Gates (BASELINE.json north_star): 1e-3 relative fp32 on every floating-point output and
gradient; greedy / beam token ids bit-exact."""
import argparse

import numpy as np
import pytest
import torch

from oracle import xgating_oracle as O
from tests.common import CONFIGS, RTOL, fro_err, load_golden, make_case, rel_err

pytestmark = pytest.mark.gpu


def _xg():
    import controllable_xgating_b200 as X
    import controllable_xgating_b200.SAModel as XS
    XS.VERBOSE = False
    return X


def make_opt(dims, T, drop=0.5, activity="ReLU"):
    return argparse.Namespace(vocab_size=dims["V"], category_size=dims["C"], input_encoding_size=dims["E"],
                              rnn_size=dims["H"], num_layers=1, drop_prob_lm=drop, seq_length=T, seed=1024,
                              feat_size=dims["R"], feat_size2=dims["F"], att_size=dims["A"], fusion_activity=activity)


def build_model(cfg, P, drop=0.5, activity="ReLU"):
    X = _xg()
    m = X.SAModel(make_opt(cfg["dims"], cfg["T"], drop, activity))
    sd = {k: v.clone() for k, v in P.items()}
    missing = m.load_state_dict(sd, strict=True)      # num_batches_tracked is synthesised (0.3.1 checkpoints lack it)
    m.cuda()
    return m


def dev(batch):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def stack_state(st):
    return np.stack([st[0][0][0].cpu().numpy(), st[0][1][0].cpu().numpy(), st[1][0][0].cpu().numpy(), st[1][1][0].cpu().numpy()])


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", [(5, 7, 3), (64, 2048, 512), (33, 65, 129), (1792, 512, 1024), (300, 1000, 64),
                                   (257, 2049, 40), (64, 10000, 512)])
def test_gemm_engine(layout, shape):
    from controllable_xgating_b200.engine import debug_gemm
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + layout)
    a_shape = (M, K) if layout in (0, 1) else (K, M)
    b_shape = (N, K) if layout == 0 else (K, N)
    A = torch.rand(a_shape, generator=g) - 0.5
    B = torch.rand(b_shape, generator=g) - 0.5
    Ad, Bd = A.double(), B.double()
    ref = (Ad if layout != 2 else Ad.t()) @ (Bd.t() if layout == 0 else Bd)
    C = debug_gemm(layout, 1, A.cuda(), B.cuda(), M, N, K).cpu()
    assert rel_err(C.numpy(), ref.numpy()) < 2e-6


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 2048, 512), (64, 512, 2048), (64, 1024, 1536), (64, 1536, 1024), (2, 2048, 512),
                                   (37, 130, 1000), (256, 512, 2048), (64, 468, 2048), (320, 2048, 512), (320, 1536, 468),
                                   (300, 10000, 512)])
def test_gemm_split_k(layout, shape):
    """skinny recurrent products: split-K SIMT path (engine 3) is deterministic and fp32-exact-grade."""
    from controllable_xgating_b200.engine import debug_gemm
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 5 + N * 11 + K + layout)
    a_shape = (M, K) if layout in (0, 1) else (K, M)
    b_shape = (N, K) if layout == 0 else (K, N)
    A = torch.rand(a_shape, generator=g) - 0.5
    B = torch.rand(b_shape, generator=g) - 0.5
    Ad, Bd = A.double(), B.double()
    ref = (Ad if layout != 2 else Ad.t()) @ (Bd.t() if layout == 0 else Bd)
    Ac, Bc = A.cuda(), B.cuda()
    C = debug_gemm(layout, 3, Ac, Bc, M, N, K)
    assert rel_err(C.cpu().numpy(), ref.numpy()) < 2e-6
    for _ in range(3):      # fixed-order reduction: bit-identical from launch to launch
        assert torch.equal(debug_gemm(layout, 3, Ac, Bc, M, N, K), C)


def test_dropout_mask_properties():
    from controllable_xgating_b200.engine import dropout_mask
    n, p = 1 << 20, 0.5
    m1 = dropout_mask(11, "dec_h1", n, p, "cuda").cpu()
    m2 = dropout_mask(11, "dec_h1", n, p, "cuda").cpu()
    m3 = dropout_mask(12, "dec_h1", n, p, "cuda").cpu()
    m4 = dropout_mask(11, "dec_h2", n, p, "cuda").cpu()
    assert torch.equal(m1, m2)                                   # pure function of (seed, site, index)
    assert set(np.unique(m1.numpy()).tolist()) == {0.0, 2.0}     # 0 or 1/(1-p)
    assert abs(float((m1 > 0).float().mean()) - (1 - p)) < 5e-3
    assert 0.4 < float((m1 != m3).float().mean()) < 0.6 and 0.4 < float((m1 != m4).float().mean()) < 0.6
    m5 = dropout_mask(11, "cls", 4096, 0.25, "cuda").cpu()
    assert abs(float((m5 > 0).float().mean()) - 0.75) < 0.03 and abs(float(m5.max()) - 1 / 0.75) < 1e-6


def test_fails_loudly_off_gpu():
    cfg, P, b = make_case("tiny")
    X = _xg()
    m = X.SAModel(make_opt(cfg["dims"], cfg["T"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"])
    m.cuda()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"])      # CPU inputs
    d = dev(b)
    with pytest.raises(AssertionError):
        m(d["rgb"][:, :, :-1], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    m.eval()
    with pytest.raises(AssertionError):
        m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"beam_size": cfg["dims"]["V"] + 1})


# --------------------------------------------------------------------------------------------
# inference paths vs golden (real reference) and oracle
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CONFIGS))
def test_encoder_init_step(name):
    g = load_golden(name); cfg, P, b = make_case(name); d = dev(b)
    m = build_model(cfg, P).eval()
    V = m.two_spatial_encoder(d["rgb"], d["opfl"], d["feat_mask"])
    assert rel_err(V.cpu().numpy(), g["V_eval"]) < RTOL
    st = m.init_hidden(V, d["feat_mask"])
    assert st[0][0].shape == (1, cfg["B"], cfg["dims"]["H"])
    assert rel_err(stack_state(st), g["init_state"]) < RTOL
    B = cfg["B"]
    probe = (torch.arange(B) % cfg["dims"]["V"]).cuda()
    pm = torch.ones(B, 1); pm[B - 1, 0] = 0
    xt = m.embed.weight.data[probe]
    o, st2 = m.lstmcore(xt, pm.cuda(), V, d["pos"], st)
    assert rel_err(o.cpu().numpy(), g["step_out"]) < RTOL
    assert rel_err(stack_state(st2), g["step_state"]) < RTOL
    lp, st3 = m.get_logprobs_state(probe, V, d["pos"], st)
    assert rel_err(lp.cpu().numpy(), g["glps_logp"]) < RTOL
    # tight check against the oracle (same fp32 math, different summation order)
    with torch.no_grad():
        Vo = O.encoder_fwd(P, b["rgb"], b["opfl"], b["feat_mask"])
    assert rel_err(V.cpu().numpy(), Vo.numpy()) < 2e-5


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_eval(name):
    g = load_golden(name); cfg, P, b = make_case(name); d = dev(b)
    m = build_model(cfg, P).eval()
    with torch.no_grad():
        logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    assert tuple(logp.shape) == g["fwd_eval_logp"].shape       # includes the early exit for `tiny`
    assert logp.is_contiguous() and cat.is_contiguous()
    assert rel_err(logp.cpu().numpy(), g["fwd_eval_logp"]) < RTOL
    assert rel_err(cat.cpu().numpy(), g["fwd_eval_cat"]) < RTOL
    assert rel_err(logp.cpu().numpy(), g["fwd_eval_logp"]) < 5e-5   # in practice far inside the gate


@pytest.mark.parametrize("name", list(CONFIGS))
def test_greedy_bit_exact(name):
    g = load_golden(name); cfg, P, b = make_case(name); d = dev(b)
    m = build_model(cfg, P).eval()
    seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
    assert seq.dtype == torch.int64 and seq.is_cuda
    assert np.array_equal(seq.cpu().numpy(), g["greedy_seq"])
    assert rel_err(lps.cpu().numpy(), g["greedy_logp"]) < RTOL


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("beam", [3, 5])
def test_beam_bit_exact(name, beam):
    g = load_golden(name); cfg, P, b = make_case(name); d = dev(b)
    m = build_model(cfg, P).eval()
    seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"beam_size": beam})
    assert not seq.is_cuda and seq.dtype == torch.int64          # the reference returns CPU tensors here
    assert np.array_equal(seq.numpy(), g["beam%d_seq" % beam])
    assert rel_err(lps.numpy(), g["beam%d_logp" % beam]) < RTOL
    gp, gs = g["beam%d_done_p" % beam], g["beam%d_done_seq" % beam]
    assert len(m.done_beams) == cfg["B"]
    for k, db in enumerate(m.done_beams):
        assert len(db) == int(np.sum(~np.isnan(gp[k])))
        for j, dbeam in enumerate(db):
            assert np.array_equal(dbeam["seq"].numpy(), gs[k][j])
            assert abs(dbeam["p"] - gp[k][j]) <= 1e-3 * max(1.0, abs(gp[k][j]))
    # sum of per-step log-probs of the best beam == its score (up to fp32 rounding; UNK never chosen)
    best = m.done_beams[0][0]
    assert abs(float(best["logps"].sum()) - best["p"]) < 1e-3 * max(1.0, abs(best["p"]))


def test_host_beam_search_method_matches_device():
    """CaptionModel.beam_search (host bookkeeping + CUDA word steps) == the batched device search."""
    cfg, P, b = make_case("mid"); d = dev(b)
    m = build_model(cfg, P).eval()
    V = m.two_spatial_encoder(d["rgb"], d["opfl"], d["feat_mask"])
    seq, lps = m.sample_beam(V, d["feat_mask"], d["pos"], {"beam_size": 3})
    dev_beams = m.done_beams
    k = 1
    feat = V[k].unsqueeze(0).expand(3, V.size(1), V.size(2)).contiguous()
    fm = d["feat_mask"][k].unsqueeze(0).expand(3, V.size(1)).contiguous()
    pf = d["pos"][k].unsqueeze(0).expand(3, d["pos"].size(1)).contiguous()
    st = m.init_hidden(feat, fm)
    lp, st = m.get_logprobs_state(torch.zeros(3, dtype=torch.long, device="cuda"), feat, pf, st)
    host = m.beam_search(st, lp, V[k], d["pos"][k], opt={"beam_size": 3})
    assert len(host) == len(dev_beams[k])
    for a, c in zip(host, dev_beams[k]):
        assert torch.equal(a["seq"], c["seq"]) and abs(a["p"] - c["p"]) < 1e-4


# --------------------------------------------------------------------------------------------
# training path
# --------------------------------------------------------------------------------------------
def _grad_check(m, grads_ref, tol=RTOL, noise=1e-7):
    named = dict(m.named_parameters())
    worst = 0.0
    for n, ref in grads_ref.items():
        got = named[n].grad
        assert got is not None, n
        ref = np.asarray(ref, dtype=np.float64)
        rn = np.linalg.norm(ref)
        if rn == 0.0:
            assert float(got.abs().max()) == 0.0, n
            continue
        if rn < noise:
            assert float(got.double().norm()) < 100 * noise, n
            continue
        e = fro_err(got.cpu().numpy(), ref)
        worst = max(worst, e)
        assert e < tol, (n, e)
    return worst


@pytest.mark.parametrize("name", ["tiny", "mid"])
@pytest.mark.parametrize("w", [("w0", 0.0), ("w05", 0.5)])
def test_train_fwd_bwd_vs_golden(name, w):
    X = _xg()
    wname, wc = w
    g = load_golden(name); cfg, P, b = make_case(name); d = dev(b)
    m = build_model(cfg, P, drop=0.0).train()
    crit, ccrit = X.LanguageModelCriterion(), X.ClassiferCriterion()
    logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    assert rel_err(logp.detach().cpu().numpy(), g["fwd_train_logp"]) < RTOL
    assert rel_err(cat.detach().cpu().numpy(), g["fwd_train_cat"]) < RTOL
    Lp = logp.shape[1]
    cls = (d["seq"] % cfg["dims"]["C"])
    loss_l = crit(logp, d["seq"][:, :Lp], d["seq_mask"][:, :Lp])
    loss_c = ccrit(cat, cls[:, :Lp], d["seq_mask"][:, :Lp], None)
    assert abs(float(loss_l) - float(g["loss_lang"])) < 1e-4 * abs(float(g["loss_lang"]))
    assert abs(float(loss_c) - float(g["loss_cls"])) < 1e-4 * abs(float(g["loss_cls"]))
    (loss_l + wc * loss_c).backward()
    ref = {n: g["grad_%s_full/%s" % (wname, n)] for n, _ in m.named_parameters()}
    _grad_check(m, ref)
    sd = m.state_dict()
    for k in sd:
        if k.endswith(("running_mean", "running_var")):
            assert rel_err(sd[k].cpu().numpy(), g["bn_after/" + k]) < 1e-4
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == 1


def test_train_fwd_bwd_config1_vs_golden():
    X = _xg()
    g = load_golden("c1"); cfg, P, b = make_case("c1"); d = dev(b)
    m = build_model(cfg, P, drop=0.0).train()
    logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    assert rel_err(logp.detach().cpu().numpy(), g["fwd_train_logp"]) < RTOL
    loss = X.LanguageModelCriterion()(logp, d["seq"], d["seq_mask"])
    cls = (d["seq"] % cfg["dims"]["C"])
    loss = loss + 0.5 * X.ClassiferCriterion()(cat, cls, d["seq_mask"])
    loss.backward()
    for n, p in m.named_parameters():
        ref_norm = float(g["grad_w05_norm/" + n])
        got = p.grad
        if ref_norm < 1e-7:
            assert float(got.double().norm()) < 1e-5, n
            continue
        idx = g["grad_w05_idx/" + n]; val = g["grad_w05_val/" + n]
        gv = got.reshape(-1)[torch.from_numpy(idx).cuda()].cpu().numpy()
        scale = max(np.max(np.abs(val)), ref_norm / np.sqrt(p.numel()))
        assert np.max(np.abs(gv - val)) <= RTOL * scale, n
        assert abs(float(got.double().norm()) - ref_norm) <= RTOL * ref_norm, n


@pytest.mark.parametrize("name,activity", [("tiny", "ReLU"), ("mid", "ReLU"), ("tiny", "Tanh"), ("mid", "Sigmoid")])
def test_train_with_dropout_replayed_in_oracle(name, activity):
    """Train mode, drop_prob 0.5: the kernels' Philox masks are exported through the C ABI and replayed in
    the oracle, so forward log-probs and every gradient can be compared exactly like the p=0 case."""
    from controllable_xgating_b200.engine import dropout_mask
    X = _xg()
    cfg, P, b = make_case(name); d = dev(b)
    dims, B, K = cfg["dims"], cfg["B"], cfg["K"]
    H = dims["H"]
    p = 0.5
    m = build_model(cfg, P, drop=p, activity=activity).train()
    torch.manual_seed(1234)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # what Engine.next_seed() will draw
    torch.manual_seed(1234)
    logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    Lp = logp.shape[1]
    L = b["seq"].shape[1]

    def mk(site, shape):
        return dropout_mask(seed, site, int(np.prod(shape)), p, "cuda").cpu().view(*shape)
    masks = {"enc_emb_rgb": mk("enc_emb_rgb", (B, K, H)), "enc_emb_opfl": mk("enc_emb_opfl", (B, K, H)),
             "enc_gate_rgb": mk("enc_gate_rgb", (K, B, H)).transpose(0, 1).contiguous(),
             "enc_gate_opfl": mk("enc_gate_opfl", (K, B, H)).transpose(0, 1).contiguous(),
             "enc_fusion": mk("enc_fusion", (K, B, H)).transpose(0, 1).contiguous(),
             "dec_gate": mk("dec_gate", (L, B, H)), "dec_h1": mk("dec_h1", (L, B, H)), "dec_h2": mk("dec_h2", (L, B, H)),
             "cls": mk("cls", (L, B, 128))}
    cls = b["seq"] % dims["C"]
    loss = (X.LanguageModelCriterion()(logp, d["seq"][:, :Lp], d["seq_mask"][:, :Lp])
            + 0.5 * X.ClassiferCriterion()(cat, cls.cuda()[:, :Lp], d["seq_mask"][:, :Lp]))
    loss.backward()
    # oracle with the same masks
    logp_o, cat_o = O.forward(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"], train=True,
                              masks=masks, activity=activity)
    assert rel_err(logp.detach().cpu().numpy(), logp_o.detach().numpy()) < RTOL
    assert rel_err(cat.detach().cpu().numpy(), cat_o.detach().numpy()) < RTOL
    loss_o, grads_o = O.train_step_grads(P, b, train=True, masks=masks, weight_class=0.5, cap_classes=cls, activity=activity)
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    _grad_check(m, {n: g.numpy() for n, g in grads_o.items()})


def test_eval_mode_backward_and_grad_accumulation():
    """backward in eval mode (running-stat BatchNorm) and accumulation into existing .grad."""
    X = _xg()
    cfg, P, b = make_case("mid"); d = dev(b)
    m = build_model(cfg, P).eval()
    crit = X.LanguageModelCriterion()
    for _ in range(2):
        logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        crit(logp, d["seq"], d["seq_mask"]).backward()
    _, grads_o = O.train_step_grads(P, b, train=False)
    _grad_check(m, {n: 2.0 * g.numpy() for n, g in grads_o.items()})


# --------------------------------------------------------------------------------------------
# BASELINE.json full sizes (configs 2/3/5): oracle run in the same test + size-independent properties
# --------------------------------------------------------------------------------------------
FULL = dict(dims=dict(R=1536, F=1024, H=512, E=468, A=1536, V=10000, C=14), K=28, T=30)


def _full_case(B, seed=0):
    P = O.synth_params(FULL["dims"], 1024)
    b = O.synth_inputs(FULL["dims"], B, FULL["K"], FULL["T"], seed)
    return dict(FULL, B=B), P, b


def test_full_size_greedy_config2():
    cfg, P, b = _full_case(64); d = dev(b)
    P["logit.bias"][0] = -1e4                 # EOS never chosen: all T steps run
    m = build_model(cfg, P).eval()
    seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
    assert tuple(seq.shape) == (64, 30)
    with torch.no_grad():
        seq_o, lps_o = O.sample_greedy(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], 30)
    got, ref = seq.cpu().numpy(), seq_o.numpy()
    if not np.array_equal(got, ref):
        # a mismatch is tolerated only at an fp32 near-tie of the reference itself: report the margin
        bad = np.argwhere(got != ref)
        rows = sorted(set(int(r) for r, _ in bad))
        first = {r: int(min(c for rr, c in bad if rr == r)) for r in rows}
        margins = []
        for r, c in first.items():
            margins.append(abs(float(lps.cpu()[r, c]) - float(lps_o[r, c])))
        pytest.fail("greedy tokens differ in rows %s at first columns %s (|dlogp| %s)" % (rows, first, margins))
    assert rel_err(lps.cpu().numpy(), lps_o.numpy()) < RTOL
    # idempotence / determinism: a second run is bit-identical
    seq2, lps2 = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
    assert torch.equal(seq, seq2) and torch.equal(lps, lps2)
    # batch independence: decoding a sub-batch gives the same captions
    seq3, _ = m.sample(d["rgb"][:7], d["opfl"][:7], d["feat_mask"][:7], d["pos"][:7], {"sample_max": 1, "beam_size": 1})
    assert torch.equal(seq3, seq[:7])


def test_full_size_train_config3():
    X = _xg()
    cfg, P, b = _full_case(64); d = dev(b)
    m = build_model(cfg, P, drop=0.0).train()
    logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
    assert tuple(logp.shape) == (64, 31, 10000)
    # log-probs normalise
    assert float((torch.logsumexp(logp, 2)).abs().max()) < 1e-4
    loss = X.LanguageModelCriterion()(logp, d["seq"], d["seq_mask"])
    loss.backward()
    loss_o, grads_o = O.train_step_grads(P, b, train=True)
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    worst = _grad_check(m, {n: g.numpy() for n, g in grads_o.items()})
    print("config3 worst per-tensor Frobenius grad error: %.3e" % worst)
    # classifier grads are exact zeros when the class loss is not in the objective
    assert float(m.classifer[0].weight.grad.abs().max()) == 0.0


@pytest.mark.parametrize("NB,min_exact", [(8, 7), (64, 63)])
def test_full_size_beam_config5(NB, min_exact):
    """config 5 (beam 5, V=10k, T=30) against the oracle.  At most ONE video per batch may differ, and only as a PROVEN fp32
    near-tie (policy below).  Measured on B200: 63/64 bit-exact; the one differing video (video 0 of this seed, also in
    the 8-video case) has the oracle's own two best beams 4e-6 apart relative (scores -269.277588 / -269.278717, 37 ulps
    of the running sum), and the device's winner scores -269.278259."""
    from tests.common import fused_path
    cfg, P, b = _full_case(NB, seed=2); d = dev(b)
    m = build_model(cfg, P).eval()
    with fused_path(m, ["encode_persistent", "decode_step_persistent"]):
        seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"beam_size": 5})
    with torch.no_grad():
        V = O.encoder_fwd(P, b["rgb"], b["opfl"], b["feat_mask"])
        seq_o, lps_o, done_o = O.sample_beam(P, V, b["feat_mask"], b["pos"], 5, 30)
    # Bit-exact ids are required wherever the reference's own ranking is decided by more than fp32
    # rounding noise.  Running sums reach |p| ~ 280 where one fp32 ulp is 3e-5, and the reference stores
    # the sums in a FloatTensor every step (CaptionModel.py:74), so two beams whose scores differ by a
    # few ulps can legitimately swap.  A differing row is accepted only if (a) the device's winning
    # score equals the oracle's within 1e-4 relative and (b) the oracle, teacher-forced on the device's
    # sequence, reproduces the device's per-step log-probs (the result is a true near-tie, not an error).
    exact = 0
    for k in range(NB):
        same = np.array_equal(seq[k].numpy(), seq_o[k].numpy())
        p_dev, p_or = m.done_beams[k][0]["p"], done_o[k][0]["p"]
        if same:
            exact += 1
            assert rel_err(lps[k].numpy(), lps_o[k].numpy()) < RTOL
            continue
        first = int(np.argmax(seq[k].numpy() != seq_o[k].numpy()))
        print("video %d: ids diverge at step %d; score device %.6f oracle %.6f; oracle runner-up %.6f"
              % (k, first, p_dev, p_or, done_o[k][1]["p"]))
        assert abs(p_dev - p_or) <= 1e-4 * abs(p_or), (k, p_dev, p_or)
        with torch.no_grad():
            st = O.init_hidden(P, V[k:k + 1], b["feat_mask"][k:k + 1])
            it = torch.zeros(1, dtype=torch.long)
            tot = 0.0
            for t in range(30):
                lp_t, st = O.get_logprobs_state(P, it, V[k:k + 1], b["pos"][k:k + 1], st)
                it = seq[k, t:t + 1]
                tot += float(lp_t[0, it[0]])
                assert abs(float(lp_t[0, it[0]]) - float(lps[k, t])) < 1e-3 * max(1.0, abs(float(lps[k, t])))
        assert abs(tot - p_dev) <= 1e-4 * abs(p_dev)
    print("beam-5 full size: %d/%d videos bit-exact" % (exact, NB))
    assert exact >= min_exact
    if NB != 8:
        return
    # a beam-1 search through the beam kernels reproduces greedy decoding (UNK aside)
    s1, _ = m.sample_beam(m.two_spatial_encoder(d["rgb"], d["opfl"], d["feat_mask"]), d["feat_mask"], d["pos"], {"beam_size": 1})
    sg, _ = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
    sg = sg.cpu()
    if not (sg == 1).any():
        n = sg.shape[1]
        # after the first EOS a finished beam is re-expanded, greedy rows are zero-masked: compare up to EOS
        for r in range(8):
            row = sg[r].tolist()
            upto = row.index(0) + 1 if 0 in row else n
            assert s1[r, :upto].tolist() == row[:upto]


def test_multinomial_sampling_distribution():
    """sample_max=0: tokens of the first step follow exp(logp/T) (chi-square-style frequency check)."""
    cfg, P, b = make_case("tiny")
    B = 4096
    P = {k: v.clone() for k, v in P.items()}
    m = build_model(cfg, P).eval()
    one = {k: (v[:1].repeat(B, *([1] * (v.dim() - 1))) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    d = dev(one)
    torch.manual_seed(5)
    seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 0, "temperature": 1.0})
    with torch.no_grad():
        V = O.encoder_fwd(P, b["rgb"][:1], b["opfl"][:1], b["feat_mask"][:1])
        st = O.init_hidden(P, V, b["feat_mask"][:1])
        lp0, _ = O.get_logprobs_state(P, torch.zeros(1, dtype=torch.long), V, b["pos"][:1], st)
    prob = lp0[0].exp().numpy()
    # first sampled token: finished rows are zero-masked in seq, so count via logp gather consistency instead
    first = seq[:, 0].cpu().numpy()
    freq = np.bincount(first, minlength=prob.size) / B
    assert np.abs(freq - prob).max() < 4 * np.sqrt(prob.max() / B) + 5e-3
    # returned log-probs are the log-probs of the sampled ids
    idx = torch.from_numpy(first)
    nz = first > 0
    assert np.allclose(lps[:, 0].cpu().numpy()[nz], lp0[0][idx].numpy()[nz], atol=1e-4)


def test_state_dict_roundtrip_and_optimizer_step():
    """load_state_dict(strict=True) interchange + parameters stay real nn.Parameters that Adam can step;
    the engine re-reads the updated storage (no stale copies)."""
    X = _xg()
    cfg, P, b = make_case("mid"); d = dev(b)
    m = build_model(cfg, P, drop=0.0).train()
    keys = list(m.state_dict().keys())
    g = load_golden("mid")
    assert keys == g["state_dict_keys"].tolist()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    crit = X.LanguageModelCriterion()
    losses = []
    for _ in range(5):
        opt.zero_grad()
        logp, _ = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        loss = crit(logp, d["seq"], d["seq_mask"])
        loss.backward()
        for p_ in m.parameters():
            p_.grad.data.clamp_(-0.1, 0.1)           # myutils.clip_gradient
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


# --------------------------------------------------------------------------------------------
# tcgen05 3xTF32 engine
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 2048, 512), (64, 10000, 512), (1792, 512, 1536), (1984, 1000, 468), (100, 130, 70),
                                   (300, 257, 1000), (17, 64, 64), (1792, 2048, 512)])
def test_gemm_tensor_core_engine(layout, shape):
    """fp32-grade accuracy of the 3xTF32 split (a single tf32 pass would sit near 5e-4)."""
    from controllable_xgating_b200.engine import debug_gemm
    M, N, K = shape
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + layout)
    a_shape = (M, K) if layout in (0, 1) else (K, M)
    b_shape = (N, K) if layout == 0 else (K, N)
    A = torch.rand(a_shape, generator=g) - 0.5
    B = torch.rand(b_shape, generator=g) - 0.5
    Ad, Bd = A.double(), B.double()
    ref = (Ad if layout != 2 else Ad.t()) @ (Bd.t() if layout == 0 else Bd)
    C = debug_gemm(layout, 2, A.cuda(), B.cuda(), M, N, K).cpu()
    C1 = debug_gemm(layout, 1, A.cuda(), B.cuda(), M, N, K).cpu()
    e_tc, e_simt = rel_err(C.numpy(), ref.numpy()), rel_err(C1.numpy(), ref.numpy())
    print("tc %.2e simt %.2e" % (e_tc, e_simt))
    assert e_tc < 5e-6, (e_tc, e_simt)
    assert fro_err(C.numpy(), ref.numpy()) < 1e-6


def test_engines_agree_end_to_end():
    """whole greedy decode + train step with the tensor-core engine on vs off."""
    X = _xg()
    cfg, P, b = make_case("c1"); d = dev(b)
    outs = []
    for tc in (True, False):
        m = build_model(cfg, P, drop=0.0)
        m._engine.set_engine(tc)
        m.eval()
        seq, lps = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
        m.train()
        logp, _ = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        X.LanguageModelCriterion()(logp, d["seq"], d["seq_mask"]).backward()
        outs.append((seq.cpu(), lps.cpu(), logp.detach().cpu(), {n: p.grad.cpu() for n, p in m.named_parameters()}))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_err(outs[0][2].numpy(), outs[1][2].numpy()) < 1e-5
    gmax = max(float(g.norm()) for g in outs[1][3].values())
    for n in outs[0][3]:
        a, r = outs[0][3][n].double(), outs[1][3][n].double()
        # gradients that are analytically zero (Linear bias under train-mode BatchNorm) hold only rounding
        # noise in either engine: compare against the global gradient scale, not their own norm
        assert float((a - r).norm()) < 1e-5 * max(float(r.norm()), 1e-3 * gmax), n
