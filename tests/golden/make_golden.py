#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference (vsislab/Controllable_XGating,
imported read-only from /root/reference/caption_src) on committed synthetic inputs.

Run once in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Outputs `tests/golden/<config>.npz`.  Inputs and weights are NOT stored: they are
regenerated bit-identically from `oracle.xgating_oracle.synth_params/synth_inputs`
(numpy PCG64 streams), so the fixtures stay small.

Shims needed to execute the Python-2.7 / PyTorch-0.3.1 sources under torch 2.x
(none of them touches the arithmetic):
  1. `h5py` stub module                    (data_io.py:16 imports it; SAModel.py:6 star-imports data_io)
  2. `.cuda()` -> identity                 (hard-coded at SAModel.py:62,121,152,194,213,215; sub_modules.py:114-115)
  3. `Tensor.narrow(dimension=)` -> `dim=` (sub_modules.py:753)
  4. beam search only: PyTorch-0.3 scalar indexing.  In 0.3.1 `t[i]` on a 1-D tensor (and
     `t[i,j]` on 2-D) returns a Python number, i.e. a COPY.  Under torch>=0.4 it returns a 0-dim
     VIEW, so `final_beam['p'] = beam_logprobs_sum[vix]` (CaptionModel.py:114) aliases the buffer and
     is overwritten by the `-1000` two lines later, which scrambles the final ranking of done beams.
     The shim makes integer indexing that yields a 0-dim tensor return `.item()` while
     `beam_search` runs, restoring the behaviour of the PyTorch the reference was written for.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/caption_src"
sys.path.insert(0, ROOT)

from oracle import xgating_oracle as O  # noqa: E402  (input/weight generators only)

CONFIGS = {
    # SURVEY.md section 8d "config 1" (plumbing)
    "c1": dict(dims=dict(R=1536, F=1024, H=512, E=468, A=1536, V=1000, C=14), B=2, K=28, T=20,
               pseed=1024, dseed=0),
    # ragged, odd sizes; padded row (b=3), short captions, EOS reached during greedy
    "tiny": dict(dims=dict(R=24, F=20, H=16, E=12, A=20, V=40, C=5), B=5, K=6, T=7,
                 pseed=7, dseed=3),
    # moderate size with 8 rows so two rows are padded and lengths are ragged
    "mid": dict(dims=dict(R=96, F=64, H=64, E=36, A=80, V=300, C=14), B=8, K=12, T=10,
                pseed=11, dseed=5),
}


def load_reference():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _narrow = torch.Tensor.narrow

    def narrow(self, *a, **k):
        if "dimension" in k:
            k["dim"] = k.pop("dimension")
        return _narrow(self, *a, **k)
    torch.Tensor.narrow = narrow
    sys.path.insert(0, REF)
    sys.argv = ["x"]
    import SAModel as RS  # the reference module
    return RS


class legacy_scalar_indexing:
    """shim 4 (see module docstring)."""

    def __enter__(self):
        self._orig = torch.Tensor.__getitem__
        orig = self._orig

        def getitem(t, idx):
            r = orig(t, idx)
            if isinstance(r, torch.Tensor) and r.dim() == 0 and not r.requires_grad:
                return r.item()
            return r
        torch.Tensor.__getitem__ = getitem

    def __exit__(self, *a):
        torch.Tensor.__getitem__ = self._orig


def make_opt(dims, T, drop):
    return argparse.Namespace(vocab_size=dims["V"], category_size=dims["C"], input_encoding_size=dims["E"],
                              rnn_size=dims["H"], num_layers=1, drop_prob_lm=drop, seq_length=T, seed=1024,
                              feat_size=dims["R"], feat_size2=dims["F"], att_size=dims["A"],
                              fusion_activity="ReLU")


def build_ref_model(RS, dims, T, drop, P):
    model = RS.SAModel(make_opt(dims, T, drop))
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k in P:
            assert tuple(v.shape) == tuple(P[k].shape), (k, v.shape, P[k].shape)
            new[k] = P[k].clone()
        else:
            assert k.endswith("num_batches_tracked"), k
            new[k] = v
    model.load_state_dict(new, strict=True)
    return model


def sample_grad(g: torch.Tensor, n=256):
    flat = g.reshape(-1)
    idx = np.linspace(0, flat.numel() - 1, num=min(n, flat.numel())).astype(np.int64)
    return idx, flat[idx].numpy()


def run_config(RS, name, cfg):
    dims, B, K, T = cfg["dims"], cfg["B"], cfg["K"], cfg["T"]
    P = O.synth_params(dims, cfg["pseed"])
    if name == "tiny":                       # make EOS reachable at ragged times in greedy/beam
        P["logit.bias"][0] = 0.02
    batch = O.synth_inputs(dims, B, K, T, cfg["dseed"])
    if name == "tiny":                       # trailing all-zero column -> early exit at SAModel.py:103
        batch["seq"] = torch.cat([batch["seq"], torch.zeros(B, 2, dtype=torch.long)], 1)
        batch["seq_mask"] = torch.cat([batch["seq_mask"], torch.zeros(B, 2)], 1)
    out = {}
    names = [n for n, _ in O.param_shapes(dims["R"], dims["F"], dims["H"], dims["E"], dims["A"], dims["V"], dims["C"])]

    # ---------------- eval mode ----------------
    model = build_ref_model(RS, dims, T, 0.5, P)
    assert [k for k in model.state_dict() if not k.endswith("num_batches_tracked")][:len(P)] is not None
    sd_keys = list(model.state_dict().keys())
    out["state_dict_keys"] = np.array(sd_keys)
    model.eval()
    with torch.no_grad():
        V = model.two_spatial_encoder(batch["rgb"], batch["opfl"], batch["feat_mask"])
        out["V_eval"] = V.numpy()
        st = model.init_hidden(V, batch["feat_mask"])
        out["init_state"] = np.stack([st[0][0][0].numpy(), st[0][1][0].numpy(), st[1][0][0].numpy(), st[1][1][0].numpy()])
        # one decoder step from the init state on token ids `probe` with a mixed mask
        probe = torch.arange(B, dtype=torch.long) % dims["V"]
        pm = torch.ones(B, 1); pm[B - 1, 0] = 0.0
        o, st2 = model.lstmcore(model.embed(probe), pm, V, batch["pos"], st)
        out["step_out"] = o.numpy()
        out["step_state"] = np.stack([st2[0][0][0].numpy(), st2[0][1][0].numpy(), st2[1][0][0].numpy(), st2[1][1][0].numpy()])
        lp, st3 = model.get_logprobs_state(probe, V, batch["pos"], st)
        out["glps_logp"] = lp.numpy()
        logp, cat = model(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], batch["seq"], batch["seq_mask"])
        out["fwd_eval_logp"] = logp.numpy(); out["fwd_eval_cat"] = cat.numpy()
        seq, slp = model.sample(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], {"sample_max": 1, "beam_size": 1})
        out["greedy_seq"] = seq.numpy(); out["greedy_logp"] = slp.numpy()
        # margin between best and second-best log-prob at every greedy step (tie diagnostics)
        for bs in (3, 5):
            if bs > dims["V"]:
                continue
            with legacy_scalar_indexing():
                bseq, blp = model.sample(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], {"beam_size": bs})
            out["beam%d_seq" % bs] = bseq.numpy().copy(); out["beam%d_logp" % bs] = blp.numpy().copy()
            out["beam%d_done_p" % bs] = np.array([[d["p"] for d in db] + [np.nan] * (bs - len(db)) for db in model.done_beams], dtype=np.float64)
            out["beam%d_done_seq" % bs] = np.stack([np.stack([np.asarray(d["seq"]) for d in db] + [np.full(T, -1)] * (bs - len(db))) for db in model.done_beams])

    # ---------------- train mode, dropout off (parity of fwd+bwd; SURVEY 8d parity gates) ----------------
    model = build_ref_model(RS, dims, T, 0.0, P)
    model.train()
    crit = RS.LanguageModelCriterion(); ccrit = RS.ClassiferCriterion()
    Lfull = batch["seq"].shape[1]
    logp, cat = model(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], batch["seq"], batch["seq_mask"])
    Lp = logp.shape[1]
    out["fwd_train_logp"] = logp.detach().numpy(); out["fwd_train_cat"] = cat.detach().numpy()
    seq_c, mask_c = batch["seq"][:, :Lp], batch["seq_mask"][:, :Lp]
    cap_classes = (batch["seq"] % dims["C"])[:, :Lp]
    loss_l = crit(logp, seq_c, mask_c)
    loss_c = ccrit(cat, cap_classes, mask_c, None)
    for wname, w in (("w0", 0.0), ("w05", 0.5)):
        model.zero_grad()
        (loss_l + w * loss_c).backward(retain_graph=True)
        out["loss_lang"] = np.float64(loss_l.item()); out["loss_cls"] = np.float64(loss_c.item())
        sd = dict(model.named_parameters())
        for n in names:
            g = sd[n].grad
            g = torch.zeros_like(sd[n]) if g is None else g.detach()
            out["grad_%s_norm/%s" % (wname, n)] = np.float64(g.double().norm().item())
            if name == "c1":
                idx, val = sample_grad(g)
                out["grad_%s_idx/%s" % (wname, n)] = idx; out["grad_%s_val/%s" % (wname, n)] = val
            else:
                out["grad_%s_full/%s" % (wname, n)] = g.numpy().copy()
    for k, v in model.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            out["bn_after/" + k] = v.numpy().copy()
    out["Lprime"] = np.int64(Lp); out["Lfull"] = np.int64(Lfull)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "saved", {k: getattr(v, "shape", None) for k, v in list(out.items())[:12]})


def main():
    RS = load_reference()
    torch.set_num_threads(8)
    for name, cfg in CONFIGS.items():
        run_config(RS, name, cfg)


if __name__ == "__main__":
    main()
