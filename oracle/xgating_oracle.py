"""CPU oracle for the gated-fusion caption decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `controllable_xgating_b200/` may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs do, and only as the checker / the timed CPU baseline.

This is an independent *functional* restatement (plain tensors in a dict keyed by the
reference's `state_dict` names, no nn.Module) of the algorithm in
vsislab/Controllable_XGating `caption_src/{SAModel,sub_modules,CaptionModel}.py`.
Every function cites the reference file:line it follows.  Arithmetic is torch CPU in
the dtype of the weights handed in (fp32 for parity / timing, fp64 for an error
yardstick).  Backward is torch autograd over this restatement, exactly as the
reference's backward is autograd over its own forward (`starttrain.py:134`).

Pinning: `tests/golden/make_golden.py` ran the *real* reference (imported from
/root/reference under the shims listed there) on the committed synthetic inputs and
stored its outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this
restatement against those vectors.  The reference itself ships no tests or golden
vectors, so that is the strongest pin available.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

BN_EPS = 1e-5          # torch.nn.BatchNorm1d default (sub_modules.py:98,102)
BN_MOMENTUM = 0.1

# ----------------------------------------------------------------------------------
# parameter table: the reference's state_dict order (SAModel.py:14-50 registration
# order; verified against an instantiated reference model in make_golden.py).
# ----------------------------------------------------------------------------------


def param_shapes(R: int, F: int, H: int, E: int, A: int, V: int, C: int,
                 cls_hidden: int = 128) -> "List[Tuple[str, Tuple[int, ...]]]":
    """(name, shape) of the 57 learnable tensors, in state_dict order."""
    enc = "two_spatial_encoder."
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for s, d in (("rgb", R), ("opfl", F)):
        out += [(enc + "visual_emb_%s.0.weight" % s, (H, d)),
                (enc + "visual_emb_%s.0.bias" % s, (H,)),
                (enc + "visual_emb_%s.1.weight" % s, (H,)),
                (enc + "visual_emb_%s.1.bias" % s, (H,))]
    for s in ("rgb", "opfl"):
        out += [(enc + "lstmcell_%s.weight_ih" % s, (4 * H, H)),
                (enc + "lstmcell_%s.weight_hh" % s, (4 * H, H)),
                (enc + "lstmcell_%s.bias_ih" % s, (4 * H,)),
                (enc + "lstmcell_%s.bias_hh" % s, (4 * H,))]
    for s in ("rgb", "opfl"):
        out += [(enc + "gate_%s.gate.0.weight" % s, (H, H)),
                (enc + "gate_%s.gate.0.bias" % s, (H,))]
    out += [(enc + "fusion.late_fusion.0.weight", (H, 2 * H)),
            (enc + "fusion.late_fusion.0.bias", (H,))]
    for n in ("h_1", "c_1", "h_2", "c_2"):
        out += [("img_embed_%s.weight" % n, (H, H)), ("img_embed_%s.bias" % n, (H,))]
    out += [("lstmcore.gate.gate.0.weight", (H, E)), ("lstmcore.gate.gate.0.bias", (H,))]
    for cell, d1 in (("lstm_1", E), ("lstm_2", H)):
        out += [("lstmcore.%s.i2h.weight" % cell, (4 * H, d1)), ("lstmcore.%s.i2h.bias" % cell, (4 * H,)),
                ("lstmcore.%s.a2h.weight" % cell, (4 * H, H)), ("lstmcore.%s.a2h.bias" % cell, (4 * H,)),
                ("lstmcore.%s.h2h.weight" % cell, (4 * H, H)), ("lstmcore.%s.h2h.bias" % cell, (4 * H,))]
    out += [("lstmcore.v2a.weight", (A, H)), ("lstmcore.v2a.bias", (A,)),
            ("lstmcore.h2a.weight", (A, 2 * H)), ("lstmcore.h2a.bias", (A,)),
            ("lstmcore.a2w.weight", (1, A)), ("lstmcore.a2w.bias", (1,)),
            ("embed.weight", (V, E)),
            ("logit.weight", (V, H)), ("logit.bias", (V,)),
            ("classifer.0.weight", (cls_hidden, H)), ("classifer.0.bias", (cls_hidden,)),
            ("classifer.3.weight", (C, cls_hidden)), ("classifer.3.bias", (C,))]
    return out


def buffer_names() -> List[str]:
    enc = "two_spatial_encoder."
    return [enc + "visual_emb_rgb.1.running_mean", enc + "visual_emb_rgb.1.running_var",
            enc + "visual_emb_opfl.1.running_mean", enc + "visual_emb_opfl.1.running_var"]


def synth_params(dims: dict, seed: int = 1024, dtype=torch.float32) -> Params:
    """Deterministic weights with the reference's init *distributions*
    (torch defaults U(-1/sqrt(fan_in), +); embed/logit U(-0.1,0.1), logit.bias 0:
    SAModel.py:52-56; BN gamma 1 / beta 0).  numpy PCG64 so the values do not depend
    on the torch version.  BN gamma/beta and running stats are perturbed away from
    their trivial init so that parity tests exercise them."""
    rng = np.random.Generator(np.random.PCG64(seed))
    P: Params = {}
    H = dims["H"]
    for name, shape in param_shapes(dims["R"], dims["F"], H, dims["E"], dims["A"], dims["V"],
                                    dims["C"], dims.get("cls_hidden", 128)):
        if name in ("embed.weight", "logit.weight"):
            a = rng.uniform(-0.1, 0.1, size=shape)
        elif name == "logit.bias":
            a = np.zeros(shape)
        elif ".1.weight" in name:           # BN gamma
            a = rng.uniform(0.5, 1.5, size=shape)
        elif ".1.bias" in name:             # BN beta
            a = rng.uniform(-0.2, 0.2, size=shape)
        elif "lstmcell_" in name:           # nn.LSTMCell: U(-1/sqrt(H), 1/sqrt(H))
            k = 1.0 / math.sqrt(H)
            a = rng.uniform(-k, k, size=shape)
        else:                                # nn.Linear: bound 1/sqrt(fan_in) for W and b
            if len(shape) == 2:
                fan_in = shape[1]
                last_fan_in = fan_in
            else:
                fan_in = last_fan_in
            k = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-k, k, size=shape)
        P[name] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    for name in buffer_names():
        if name.endswith("running_mean"):
            a = rng.uniform(-0.1, 0.1, size=(H,))
        else:
            a = rng.uniform(0.05, 0.15, size=(H,))
        P[name] = torch.from_numpy(a).to(dtype)
    return P


def synth_inputs(dims: dict, B: int, K: int, T: int, seed: int = 0, dtype=torch.float32,
                 full_length: bool = False) -> dict:
    """Synthetic MSRVTT-shape batch (mirrors the tensor contract of
    data_io.collate_fn, data_io.py:330-374): rgb/opfl ~ U[0,1); every 4th row has its
    last min(8,K-1) frames zero-padded with feat_mask 0; pos ~ U(-1,1); seq column 0 = 0,
    tokens ~ U{2..V-1}; row 0 has full length T, others U{min(5,T)..T}; rows sorted by
    length descending; seq_mask has len+1 ones."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, F, H, V = dims["R"], dims["F"], dims["H"], dims["V"]
    rgb = rng.random((B, K, R))
    opfl = rng.random((B, K, F))
    fmask = np.ones((B, K))
    npad = min(8, K - 1)
    for b in range(B):
        if b % 4 == 3 and npad > 0:
            rgb[b, K - npad:] = 0.0
            opfl[b, K - npad:] = 0.0
            fmask[b, K - npad:] = 0.0
    pos = rng.uniform(-1.0, 1.0, size=(B, H))
    L = T + 1
    lens = [T] + [int(rng.integers(min(5, T), T + 1)) for _ in range(B - 1)]
    if full_length:
        lens = [T] * B
    lens = sorted(lens, reverse=True)
    seq = np.zeros((B, L), dtype=np.int64)
    smask = np.zeros((B, L))
    for b, n in enumerate(lens):
        seq[b, 1:n + 1] = rng.integers(2, V, size=n)
        smask[b, :n + 1] = 1.0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return {"rgb": t(rgb), "opfl": t(opfl), "feat_mask": t(fmask), "pos": t(pos),
            "seq": torch.from_numpy(seq), "seq_mask": t(smask), "lens": lens}


# ----------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------


def _linear(x: Tensor, P: Params, prefix: str) -> Tensor:
    return x @ P[prefix + ".weight"].t() + P[prefix + ".bias"]


def _drop(x: Tensor, masks: Optional[dict], site: str) -> Tensor:
    """nn.Dropout in train mode with an *explicit* pre-scaled mask (values 0 or 1/(1-p))
    so a CUDA run's Philox masks can be replayed here.  masks=None == eval / p=0."""
    if masks is None or site not in masks:
        return x
    return x * masks[site]


def gate_fwd(P: Params, prefix: str, source: Tensor, target: Tensor,
             masks: Optional[dict] = None, site: str = "") -> Tensor:
    """Gate, simple=True branch: target * (1 + dropout(relu(Linear(source)))).
    sub_modules.py:29-32,42-47."""
    g = _drop(torch.relu(_linear(source, P, prefix + ".gate.0")), masks, site)
    return g * target + target


_ACT = {"ReLU": torch.relu, "Tanh": torch.tanh, "Sigmoid": torch.sigmoid}


def fusion_fwd(P: Params, f1: Tensor, f2: Tensor, activity: str = "ReLU",
               masks: Optional[dict] = None) -> Tensor:
    """Fusion: dropout(act(Linear(cat[f1,f2]))).  sub_modules.py:62-72."""
    x = torch.cat([f1, f2], -1)
    return _drop(_ACT[activity](_linear(x, P, "two_spatial_encoder.fusion.late_fusion.0")),
                 masks, "enc_fusion")


def _bn_relu(x: Tensor, P: Params, prefix: str, train: bool,
             new_stats: Optional[dict]) -> Tensor:
    """Linear -> BatchNorm1d -> ReLU over the flattened (B*K, d) rows, padded rows
    included.  sub_modules.py:97-104,121,126."""
    y = _linear(x, P, prefix + ".0")
    if train:
        mean = y.mean(0)
        var = y.var(0, unbiased=False)
        if new_stats is not None:
            n = y.shape[0]
            with torch.no_grad():
                new_stats[prefix + ".1.running_mean"] = (
                    (1 - BN_MOMENTUM) * P[prefix + ".1.running_mean"] + BN_MOMENTUM * mean)
                new_stats[prefix + ".1.running_var"] = (
                    (1 - BN_MOMENTUM) * P[prefix + ".1.running_var"]
                    + BN_MOMENTUM * var * (n / max(n - 1, 1)))
    else:
        mean = P[prefix + ".1.running_mean"]
        var = P[prefix + ".1.running_var"]
    yhat = (y - mean) / torch.sqrt(var + BN_EPS)
    return torch.relu(yhat * P[prefix + ".1.weight"] + P[prefix + ".1.bias"])


def _lstmcell_ifgo(x: Tensor, h: Tensor, c: Tensor, P: Params, prefix: str) -> Tuple[Tensor, Tensor]:
    """torch.nn.LSTMCell (gate order i,f,g,o) as used at sub_modules.py:106-107,138,145."""
    H = h.shape[1]
    z = (x @ P[prefix + ".weight_ih"].t() + P[prefix + ".bias_ih"]
         + h @ P[prefix + ".weight_hh"].t() + P[prefix + ".bias_hh"])
    i = torch.sigmoid(z[:, 0:H]); f = torch.sigmoid(z[:, H:2 * H])
    g = torch.tanh(z[:, 2 * H:3 * H]); o = torch.sigmoid(z[:, 3 * H:4 * H])
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def encoder_fwd(P: Params, rgb: Tensor, opfl: Tensor, fmask: Tensor, train: bool = False,
                masks: Optional[dict] = None, activity: str = "ReLU",
                new_stats: Optional[dict] = None) -> Tensor:
    """EncoderLstm_two_fc.forward = the Cross-Gating block.  sub_modules.py:118-159.
    Dropout sites (train only): enc_emb_rgb, enc_emb_opfl (B,K,H); enc_gate_rgb,
    enc_gate_opfl (B,K,H; frame-major replay of the per-frame masks); enc_fusion."""
    enc = "two_spatial_encoder."
    B, K = rgb.shape[0], rgb.shape[1]
    H = P[enc + "gate_rgb.gate.0.bias"].shape[0]
    m3 = fmask.unsqueeze(-1)
    e_rgb = _bn_relu(rgb.reshape(B * K, -1), P, enc + "visual_emb_rgb", train, new_stats).view(B, K, H)
    e_rgb = _drop(e_rgb, masks, "enc_emb_rgb") * m3
    e_of = _bn_relu(opfl.reshape(B * K, -1), P, enc + "visual_emb_opfl", train, new_stats).view(B, K, H)
    e_of = _drop(e_of, masks, "enc_emb_opfl") * m3
    h_r = rgb.new_zeros(B, H); c_r = rgb.new_zeros(B, H)
    h_o = rgb.new_zeros(B, H); c_o = rgb.new_zeros(B, H)
    out_r, out_o = [], []
    for t in range(K):
        mt = fmask[:, t].unsqueeze(-1)
        h_r, c_r = _lstmcell_ifgo(e_rgb[:, t], h_r, c_r, P, enc + "lstmcell_rgb")
        h_r = h_r * mt; c_r = c_r * mt                               # :139-140 (zeroing mask)
        h_o, c_o = _lstmcell_ifgo(e_of[:, t], h_o, c_o, P, enc + "lstmcell_opfl")
        h_o = h_o * mt; c_o = c_o * mt                               # :146-147
        mk_r = None if masks is None or "enc_gate_rgb" not in masks else {"s": masks["enc_gate_rgb"][:, t]}
        mk_o = None if masks is None or "enc_gate_opfl" not in masks else {"s": masks["enc_gate_opfl"][:, t]}
        out_r.append(gate_fwd(P, enc + "gate_rgb", h_o, h_r, mk_r, "s"))    # :151
        out_o.append(gate_fwd(P, enc + "gate_opfl", h_r, h_o, mk_o, "s"))   # :152
    g_r = torch.stack(out_r, 1); g_o = torch.stack(out_o, 1)
    return fusion_fwd(P, g_r, g_o, activity, masks)                  # :158


def init_hidden(P: Params, V: Tensor, fmask: Tensor) -> List[Tuple[Tensor, Tensor]]:
    """SAModel.init_hidden: all-frame sum / valid-frame count, DETACHED (numpy round trip
    at SAModel.py:59-62), then four Linears.  Returns [(h1,c1),(h2,c2)], each (1,B,H)."""
    with torch.no_grad():
        mean = (V.detach().sum(1) / fmask.sum(1).unsqueeze(-1)).unsqueeze(0)
    s1 = (_linear(mean, P, "img_embed_h_1"), _linear(mean, P, "img_embed_c_1"))
    s2 = (_linear(mean, P, "img_embed_h_2"), _linear(mean, P, "img_embed_c_2"))
    return [s1, s2]


def two_inputs_lstmcell(P: Params, prefix: str, x1: Tensor, x2: Tensor,
                        state: Tuple[Tensor, Tensor], mask: Optional[Tensor],
                        masks: Optional[dict] = None, site: str = "") -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """two_inputs_lstmcell.forward: gate order i,f,o,g; mask *carries* state; dropout on
    the carried h.  sub_modules.py:750-770."""
    h_prev, c_prev = state[0][-1], state[1][-1]
    H = h_prev.shape[1]
    z = _linear(x1, P, prefix + ".i2h") + _linear(x2, P, prefix + ".a2h") + _linear(h_prev, P, prefix + ".h2h")
    sg = torch.sigmoid(z[:, :3 * H])
    i, f, o = sg[:, :H], sg[:, H:2 * H], sg[:, 2 * H:3 * H]
    g = torch.tanh(z[:, 3 * H:])
    c = f * c_prev + i * g
    if mask is not None:
        c = c * mask + c_prev * (1.0 - mask)
    h = o * torch.tanh(c)
    if mask is not None:
        h = h * mask + h_prev * (1.0 - mask)
    h = _drop(h, masks, site)
    return h, (h.unsqueeze(0), c.unsqueeze(0))


def attention(P: Params, V: Tensor, h1: Tensor, h2: Tensor, Uv: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Temporal soft attention, softmax over ALL K frames (no mask).
    sub_modules.py:677-680.  `Uv` may carry a precomputed v2a(V) (loop invariant)."""
    if Uv is None:
        Uv = _linear(V, P, "lstmcore.v2a")
    a = _linear(torch.cat([h1, h2], 1), P, "lstmcore.h2a").unsqueeze(1) + Uv
    s = _linear(torch.tanh(a), P, "lstmcore.a2w")                  # (B,K,1)
    alpha = torch.softmax(s, dim=1)
    return (alpha * V).sum(1), alpha.squeeze(-1)


def decoder_step(P: Params, xt: Tensor, xt_mask: Tensor, V: Tensor, pos: Tensor,
                 state: Sequence[Tuple[Tensor, Tensor]], masks: Optional[dict] = None,
                 hoist_v2a: Optional[Tensor] = None):
    """LSTMCore_two_layer_gate.forward: one word step.  sub_modules.py:671-687.
    The attention reads the PREVIOUS step's h1,h2."""
    assert len(state) == 2
    st1, st2 = state
    af, _ = attention(P, V, st1[0][-1], st2[0][-1], hoist_v2a)
    gp = gate_fwd(P, "lstmcore.gate", xt, pos, masks, "dec_gate")
    o1, st1n = two_inputs_lstmcell(P, "lstmcore.lstm_1", xt, gp, st1, xt_mask, masks, "dec_h1")
    o2, st2n = two_inputs_lstmcell(P, "lstmcore.lstm_2", o1, af, st2, xt_mask, masks, "dec_h2")
    return o2, [st1n, st2n]


def word_logprobs(P: Params, out: Tensor) -> Tensor:
    """log_softmax(logit(out)).  SAModel.py:109,125,154,217."""
    return torch.log_softmax(_linear(out, P, "logit"), dim=1)


def category_logprobs(P: Params, out: Tensor, masks: Optional[dict] = None) -> Tensor:
    """classifer = Linear(H,128) ReLU Dropout Linear(128,C); log_softmax.  SAModel.py:46-49,110."""
    x = _drop(torch.relu(_linear(out, P, "classifer.0")), masks, "cls")
    return torch.log_softmax(_linear(x, P, "classifer.3"), dim=1)


def _step_masks(masks: Optional[dict], i: int) -> Optional[dict]:
    if masks is None:
        return None
    return {k: v[i] for k, v in masks.items() if k in ("dec_gate", "dec_h1", "dec_h2", "cls")}


def forward(P: Params, rgb: Tensor, opfl: Tensor, fmask: Tensor, pos: Tensor, seq: Tensor,
            seq_mask: Tensor, train: bool = False, masks: Optional[dict] = None,
            activity: str = "ReLU", hoist_v2a: bool = False,
            new_stats: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """SAModel.forward, teacher forced, ss_prob == 0.  SAModel.py:67-115.
    Decoder dropout masks are indexed [step] -> (B, .)."""
    V = encoder_fwd(P, rgb, opfl, fmask, train, masks, activity, new_stats)
    state = init_hidden(P, V, fmask)
    Uv = _linear(V, P, "lstmcore.v2a") if hoist_v2a else None
    outs, cats = [], []
    for i in range(seq.shape[1]):
        if i >= 1 and int(seq[:, i].sum()) == 0:                   # :103
            break
        xt = P["embed.weight"][seq[:, i]]
        m = seq_mask[:, i].unsqueeze(1)
        sm = _step_masks(masks, i)
        out, state = decoder_step(P, xt, m, V, pos, state, sm, Uv)
        outs.append(word_logprobs(P, out))
        cats.append(category_logprobs(P, out, sm))
    return torch.stack(outs, 1).contiguous(), torch.stack(cats, 1).contiguous()


def get_logprobs_state(P: Params, it: Tensor, V: Tensor, pos: Tensor, state):
    """SAModel.get_logprobs_state (mask == 1).  SAModel.py:117-127."""
    xt = P["embed.weight"][it]
    m = V.new_ones(it.shape[0], 1)
    out, state = decoder_step(P, xt, m, V, pos, state)
    return word_logprobs(P, out), state


def sample_greedy(P: Params, rgb: Tensor, opfl: Tensor, fmask: Tensor, pos: Tensor,
                  seq_length: int, activity: str = "ReLU", hoist_v2a: bool = False):
    """SAModel.sample with sample_max=1, beam_size=1 (eval mode).  SAModel.py:163-219.
    Quirks kept: the raw argmax is embedded even for finished rows (:198 before :208);
    returned log-probs are not masked; loop breaks when no row is unfinished (:206)."""
    V = encoder_fwd(P, rgb, opfl, fmask, False, None, activity)
    B = V.shape[0]
    state = init_hidden(P, V, fmask)
    Uv = _linear(V, P, "lstmcore.v2a") if hoist_v2a else None
    seq, lps = [], []
    logprobs = None
    unfinished = None
    for t in range(seq_length + 1):
        if t == 0:
            it = torch.zeros(B, dtype=torch.long)
        else:
            slp, it = torch.max(logprobs, 1)
        xt = P["embed.weight"][it]
        if t >= 1:
            unfinished = (it > 0) if t == 1 else (unfinished & (it > 0))
            if int(unfinished.sum()) == 0:
                break
            it = it * unfinished.to(it.dtype)
            seq.append(it); lps.append(slp)
        m = V.new_ones(B, 1) if t == 0 else unfinished.unsqueeze(-1).to(V.dtype)
        out, state = decoder_step(P, xt, m, V, pos, state, None, Uv)
        logprobs = word_logprobs(P, out)
    if not seq:
        return torch.zeros(B, 0, dtype=torch.long), V.new_zeros(B, 0)
    return torch.stack(seq, 1), torch.stack(lps, 1)


# ----------------------------------------------------------------------------------
# beam search (CaptionModel.py:22-128, SAModel.py:129-161), PyTorch-0.3 scalar
# semantics: indexing a tensor down to one element yields a Python number (a COPY, in
# double precision for floats), so candidate sums are double adds of fp32 values and
# `final_beam['p']` is a snapshot.  (Under torch>=0.4 the same source aliases
# `beam_logprobs_sum[vix]` as a 0-dim view and `p` is later overwritten by -1000; the
# golden script restores the 0.3 semantics, see tests/golden/make_golden.py.)
# ----------------------------------------------------------------------------------


def beam_search_one(P: Params, state, logprobs: Tensor, feat: Tensor, pos_feat: Tensor,
                    beam_size: int, seq_length: int):
    T = seq_length
    beam_seq = torch.zeros(T, beam_size, dtype=torch.long)
    beam_lp = torch.zeros(T, beam_size, dtype=torch.float32)
    beam_sum = [0.0] * beam_size                     # fp32-representable doubles
    done = []
    for t in range(T):
        lpf = logprobs.detach().to(torch.float32).clone()
        lpf[:, 1] = lpf[:, 1] - 1000                 # CaptionModel.py:94 (UNK suppression)
        ys, ix = torch.sort(lpf, 1, True)            # :39
        cols = min(beam_size, ys.shape[1])
        rows = 1 if t == 0 else beam_size
        cand = []
        for c in range(cols):                        # :45-50 (column-major enumeration)
            for q in range(rows):
                r = float(ys[q, c])
                cand.append((beam_sum[q] + r, int(ix[q, c]), q, r))
        cand.sort(key=lambda x: -x[0])               # :51 (stable)
        new_state = [[s.clone() for s in state[0]], [s.clone() for s in state[1]]]
        prev_seq = beam_seq[:t].clone(); prev_lp = beam_lp[:t].clone()
        new_sum = list(beam_sum)
        for vix in range(beam_size):                 # :60-74
            p, c, q, r = cand[vix]
            if t >= 1:
                beam_seq[:t, vix] = prev_seq[:, q]
                beam_lp[:t, vix] = prev_lp[:, q]
            for layer in range(2):
                for si in range(2):
                    new_state[layer][si][:, vix] = state[layer][si][:, q]
            beam_seq[t, vix] = c
            beam_lp[t, vix] = r
            new_sum[vix] = float(np.float32(p))      # stored into a FloatTensor (:74)
        beam_sum = new_sum
        state = [tuple(new_state[0]), tuple(new_state[1])]
        for vix in range(beam_size):                 # :108-118
            if int(beam_seq[t, vix]) == 0 or t == T - 1:
                done.append({"seq": beam_seq[:, vix].clone(), "logps": beam_lp[:, vix].clone(),
                             "p": beam_sum[vix]})
                beam_sum[vix] = -1000.0
        it = beam_seq[t]
        feat_ = feat.unsqueeze(0).expand(beam_size, feat.shape[0], feat.shape[1])
        pos_ = pos_feat.unsqueeze(0).expand(beam_size, pos_feat.shape[0])
        logprobs, state = get_logprobs_state(P, it, feat_, pos_, state)     # :125
    done.sort(key=lambda x: -x["p"])                 # :127
    return done[:beam_size]


def sample_beam(P: Params, V: Tensor, fmask: Tensor, pos: Tensor, beam_size: int, seq_length: int):
    """SAModel.sample_beam (takes the FUSED feats).  SAModel.py:129-161."""
    B = V.shape[0]
    assert beam_size <= P["logit.bias"].shape[0]
    seq = torch.zeros(seq_length, B, dtype=torch.long)
    lps = torch.zeros(seq_length, B, dtype=torch.float32)
    done_beams = []
    for k in range(B):
        feat = V[k].unsqueeze(0).expand(beam_size, V.shape[1], V.shape[2])
        fm = fmask[k].unsqueeze(0).expand(beam_size, fmask.shape[1])
        pf = pos[k].unsqueeze(0).expand(beam_size, pos.shape[1])
        state = init_hidden(P, feat, fm)
        it = torch.zeros(beam_size, dtype=torch.long)
        xt = P["embed.weight"][it]
        out, state = decoder_step(P, xt, V.new_ones(beam_size, 1), feat, pf, state)
        logprobs = word_logprobs(P, out)
        db = beam_search_one(P, state, logprobs, V[k], pos[k], beam_size, seq_length)
        done_beams.append(db)
        seq[:, k] = db[0]["seq"]; lps[:, k] = db[0]["logps"]
    return seq.t(), lps.t(), done_beams


def sample(P: Params, rgb, opfl, fmask, pos, seq_length: int, opt: Optional[dict] = None,
           activity: str = "ReLU"):
    """SAModel.sample dispatch (greedy / beam).  SAModel.py:163-176."""
    opt = opt or {}
    beam = opt.get("beam_size", 1)
    if beam > 1:
        V = encoder_fwd(P, rgb, opfl, fmask, False, None, activity)
        s, l, _ = sample_beam(P, V, fmask, pos, beam, seq_length)
        return s, l
    return sample_greedy(P, rgb, opfl, fmask, pos, seq_length, activity)


# ----------------------------------------------------------------------------------
# criterions (SAModel.py:221-267)
# ----------------------------------------------------------------------------------


def language_model_criterion(logp: Tensor, target: Tensor, mask: Tensor) -> Tensor:
    """LanguageModelCriterion: target rotated left by one; -sum(logp[target]*mask)/sum(mask).
    SAModel.py:225-234.  (As in the reference, logp must cover all target columns.)"""
    Vn = logp.shape[2]
    tgt = torch.cat((target[:, 1:], target[:, 0:1]), dim=1).reshape(-1, 1)
    out = -1.0 * logp.reshape(-1, Vn).gather(1, tgt) * mask.reshape(-1, 1)
    return out.sum() / mask.sum()


def classifer_criterion(logp: Tensor, target: Tensor, mask: Tensor, class_mask: Optional[Tensor] = None) -> Tensor:
    """ClassiferCriterion (no rotation; optional class_mask).  SAModel.py:241-253."""
    Cn = logp.shape[2]
    out = -1.0 * logp.reshape(-1, Cn).gather(1, target.reshape(-1, 1)) * mask.reshape(-1, 1)
    if class_mask is None:
        return out.sum() / mask.sum()
    cm = class_mask.reshape(-1, 1)
    return (out * cm).sum() / (mask.reshape(-1, 1) * cm).sum()


def reward_criterion(logp: Tensor, seq: Tensor, reward: Tensor) -> Tensor:
    """RewardCriterion (SCST).  SAModel.py:259-267."""
    mask = (seq > 0).to(logp.dtype)
    mask = torch.cat([mask.new_ones(mask.shape[0], 1), mask[:, :-1]], 1).reshape(-1)
    out = -logp.reshape(-1) * reward.reshape(-1) * mask
    return out.sum() / mask.sum()


def train_step_grads(P: Params, batch: dict, train: bool = True, masks: Optional[dict] = None,
                     weight_class: float = 0.0, cap_classes: Optional[Tensor] = None,
                     class_mask: Optional[Tensor] = None, activity: str = "ReLU"):
    """forward + LanguageModelCriterion (+ weight_class * ClassiferCriterion) + backward,
    as starttrain.py:125-134.  Returns (loss, {name: grad}) for the 57 learnables."""
    names = [n for n in P if not n.endswith(("running_mean", "running_var"))]
    Q = {n: (P[n].detach().clone().requires_grad_(True) if n in names else P[n]) for n in P}
    logp, cat = forward(Q, batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"],
                        batch["seq"], batch["seq_mask"], train=train, masks=masks, activity=activity)
    L = logp.shape[1]
    loss = language_model_criterion(logp, batch["seq"][:, :L], batch["seq_mask"][:, :L])
    if cap_classes is None:
        cap_classes = torch.zeros_like(batch["seq"])
    loss_c = classifer_criterion(cat, cap_classes[:, :L], batch["seq_mask"][:, :L], class_mask)
    total = loss + weight_class * loss_c
    total.backward()
    grads = {n: (Q[n].grad if Q[n].grad is not None else torch.zeros_like(Q[n])) for n in names}
    return total.detach(), grads
