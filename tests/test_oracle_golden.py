"""Pin the CPU oracle (oracle/xgating_oracle.py) to the golden vectors produced by the real
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import xgating_oracle as O
from tests.common import CONFIGS, RTOL, fro_err, load_golden, make_case, rel_err

NOISE = 1e-7   # |grad| below this is rounding noise around an analytic zero
TIGHT = 2e-5   # fp32 restatement vs fp32 reference: same math, different op order


@pytest.mark.parametrize("name", list(CONFIGS))
def test_state_dict_keys(name):
    g = load_golden(name); cfg = CONFIGS[name]; d = cfg["dims"]
    keys = [k for k in g["state_dict_keys"].tolist() if not k.endswith("num_batches_tracked")]
    mine = [n for n, _ in O.param_shapes(d["R"], d["F"], d["H"], d["E"], d["A"], d["V"], d["C"])]
    learn = [k for k in keys if not k.endswith(("running_mean", "running_var"))]
    assert learn == mine
    assert [k for k in keys if k.endswith(("running_mean", "running_var"))] == O.buffer_names()
    assert len(g["state_dict_keys"]) == 63


@pytest.mark.parametrize("name", list(CONFIGS))
def test_encoder_init_step(name):
    g = load_golden(name); cfg, P, b = make_case(name)
    with torch.no_grad():
        V = O.encoder_fwd(P, b["rgb"], b["opfl"], b["feat_mask"])
        assert rel_err(V.numpy(), g["V_eval"]) < TIGHT
        st = O.init_hidden(P, V, b["feat_mask"])
        got = np.stack([st[0][0][0].numpy(), st[0][1][0].numpy(), st[1][0][0].numpy(), st[1][1][0].numpy()])
        assert rel_err(got, g["init_state"]) < TIGHT
        B = cfg["B"]
        probe = torch.arange(B) % cfg["dims"]["V"]
        pm = torch.ones(B, 1); pm[B - 1, 0] = 0
        o, st2 = O.decoder_step(P, P["embed.weight"][probe], pm, V, b["pos"], st)
        assert rel_err(o.numpy(), g["step_out"]) < TIGHT
        got2 = np.stack([st2[0][0][0].numpy(), st2[0][1][0].numpy(), st2[1][0][0].numpy(), st2[1][1][0].numpy()])
        assert rel_err(got2, g["step_state"]) < TIGHT
        lp, _ = O.get_logprobs_state(P, probe, V, b["pos"], st)
        assert rel_err(lp.numpy(), g["glps_logp"]) < TIGHT
        # the loop-invariant v2a(V) hoist is algebraically identical
        o_h, _ = O.decoder_step(P, P["embed.weight"][probe], pm, V, b["pos"], st,
                                hoist_v2a=O._linear(V, P, "lstmcore.v2a"))
        assert rel_err(o_h.numpy(), g["step_out"]) < TIGHT


@pytest.mark.parametrize("name", list(CONFIGS))
def test_forward_eval(name):
    g = load_golden(name); cfg, P, b = make_case(name)
    with torch.no_grad():
        logp, cat = O.forward(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"])
    assert logp.shape == g["fwd_eval_logp"].shape            # includes the early exit (tiny: L' < L)
    assert int(g["Lprime"]) == logp.shape[1]
    if name == "tiny":
        assert int(g["Lprime"]) < int(g["Lfull"])
    assert rel_err(logp.numpy(), g["fwd_eval_logp"]) < TIGHT
    assert rel_err(cat.numpy(), g["fwd_eval_cat"]) < TIGHT


@pytest.mark.parametrize("name", list(CONFIGS))
def test_greedy_bit_exact(name):
    g = load_golden(name); cfg, P, b = make_case(name)
    with torch.no_grad():
        seq, lps = O.sample_greedy(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], cfg["T"])
    assert np.array_equal(seq.numpy(), g["greedy_seq"])
    assert rel_err(lps.numpy(), g["greedy_logp"]) < TIGHT
    if name == "tiny":                                        # EOS reached at ragged steps
        assert (g["greedy_seq"] == 0).any() and (g["greedy_seq"][:, 0] != 0).all()


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("beam", [3, 5])
def test_beam_bit_exact(name, beam):
    g = load_golden(name); cfg, P, b = make_case(name)
    with torch.no_grad():
        V = O.encoder_fwd(P, b["rgb"], b["opfl"], b["feat_mask"])
        seq, lps, done = O.sample_beam(P, V, b["feat_mask"], b["pos"], beam, cfg["T"])
    assert np.array_equal(seq.numpy(), g["beam%d_seq" % beam])
    assert rel_err(lps.numpy(), g["beam%d_logp" % beam]) < TIGHT
    for k, db in enumerate(done):
        gp = g["beam%d_done_p" % beam][k]
        gs = g["beam%d_done_seq" % beam][k]
        for j, d in enumerate(db):
            assert abs(d["p"] - gp[j]) <= 1e-4 * max(1.0, abs(gp[j]))
            assert np.array_equal(d["seq"].numpy(), gs[j])


@pytest.mark.parametrize("name", list(CONFIGS))
def test_train_forward_backward(name):
    g = load_golden(name); cfg, P, b = make_case(name)
    d = cfg["dims"]
    stats = {}
    logp, cat = O.forward(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"],
                          train=True, new_stats=stats)
    assert rel_err(logp.detach().numpy(), g["fwd_train_logp"]) < TIGHT
    for k, v in stats.items():
        assert rel_err(v.numpy(), g["bn_after/" + k]) < TIGHT
    Lp = logp.shape[1]
    cls = (b["seq"] % d["C"])
    for wname, w in (("w0", 0.0), ("w05", 0.5)):
        loss, grads = O.train_step_grads(P, b, train=True, weight_class=w, cap_classes=cls)
        assert abs(float(loss) - (float(g["loss_lang"]) + w * float(g["loss_cls"]))) < 1e-5
        for n, gr in grads.items():
            ref_norm = float(g["grad_%s_norm/%s" % (wname, n)])
            if ref_norm == 0.0:
                assert float(gr.norm()) == 0.0, n             # classifier grads are exact zeros at weight 0
                continue
            if ref_norm < NOISE:                              # Linear bias feeding train-mode BN: d/db == 0
                assert float(gr.norm()) < 10 * NOISE, n       # analytically; the reference holds rounding noise
                continue
            if name == "c1":
                idx = g["grad_%s_idx/%s" % (wname, n)]; val = g["grad_%s_val/%s" % (wname, n)]
                got = gr.reshape(-1)[idx].numpy()
                assert np.max(np.abs(got - val)) <= 1e-4 * max(np.max(np.abs(val)), ref_norm / np.sqrt(gr.numel())), n
                assert abs(float(gr.double().norm()) - ref_norm) <= 1e-4 * ref_norm, n
            else:
                assert fro_err(gr.numpy(), g["grad_%s_full/%s" % (wname, n)]) < 1e-4, n


def test_encoder_gets_no_grad_through_init_state():
    """init_hidden detaches the graph (SAModel.py:59-62): with the attention context removed from the
    loss path the encoder must receive exactly zero gradient."""
    cfg, P, b = make_case("tiny")
    Q = {k: v.clone().requires_grad_(not k.endswith(("running_mean", "running_var"))) for k, v in P.items()}
    V = O.encoder_fwd(Q, b["rgb"], b["opfl"], b["feat_mask"])
    st = O.init_hidden(Q, V, b["feat_mask"])
    (st[0][0].sum() + st[1][1].sum()).backward()
    assert Q["two_spatial_encoder.fusion.late_fusion.0.weight"].grad is None
    assert Q["img_embed_h_1.weight"].grad is not None


def test_fp64_yardstick():
    """fp32 oracle vs the same restatement in fp64: quantifies the reference's own rounding noise,
    the floor under every 1e-3 gate."""
    cfg, P, b = make_case("mid")
    _, P64, b64 = make_case("mid", torch.float64)
    with torch.no_grad():
        l32, _ = O.forward(P, b["rgb"], b["opfl"], b["feat_mask"], b["pos"], b["seq"], b["seq_mask"])
        l64, _ = O.forward(P64, b64["rgb"], b64["opfl"], b64["feat_mask"], b64["pos"], b64["seq"], b64["seq_mask"])
    assert rel_err(l32.numpy(), l64.numpy()) < 1e-5 < RTOL
