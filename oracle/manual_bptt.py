"""Explicit (autograd-free) forward-with-saved-activations and BPTT for the caption path.

TEST INFRASTRUCTURE ONLY (same rules as xgating_oracle.py).

The reference has no backward source: its backward is torch autograd (`starttrain.py:134`).
The CUDA product implements the backward by hand, so this file writes the same hand
derivation in plain torch tensor ops, at the granularity and in the buffer layouts the
CUDA host code uses (csrc/xg_train.cu), and `tests/test_manual_bptt.py` checks it against
autograd over `xgating_oracle.forward`.  It exists to separate "the derivation is wrong"
from "the kernel is wrong".

Layouts: encoder recurrent buffers are frame-major (K, B, .); decoder per-step buffers
are step-major (L, B, .); V and Uv are batch-major (B, K, .).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import xgating_oracle as O

Tensor = torch.Tensor


def _m(masks, key, shape, like):
    if masks is None or key not in masks:
        return torch.ones(shape, dtype=like.dtype)
    return masks[key]


def forward_saved(P, batch, train: bool, masks: Optional[dict] = None, activity: str = "ReLU"):
    """Returns (logp (B,L',V), cat (B,L',C), saved)."""
    enc = "two_spatial_encoder."
    rgb, opfl, fmask, pos, seq, smask = (batch[k] for k in ("rgb", "opfl", "feat_mask", "pos", "seq", "seq_mask"))
    B, K = rgb.shape[0], rgb.shape[1]
    H = P["img_embed_h_1.bias"].shape[0]
    S: Dict[str, Tensor] = {}
    fm_tb = fmask.t().contiguous()                                       # (K,B)

    # ---- encoder: embed + BN + relu + drop + mask, frame-major output ----
    for s, x in (("rgb", rgb), ("opfl", opfl)):
        pre = enc + "visual_emb_%s" % s
        Y = x.reshape(B * K, -1) @ P[pre + ".0.weight"].t() + P[pre + ".0.bias"]      # rows (b,k)
        if train:
            mean = Y.mean(0); var = Y.var(0, unbiased=False)
        else:
            mean = P[pre + ".1.running_mean"]; var = P[pre + ".1.running_var"]
        invstd = 1.0 / torch.sqrt(var + O.BN_EPS)
        xhat = (Y - mean) * invstd
        bn = xhat * P[pre + ".1.weight"] + P[pre + ".1.bias"]
        dm = _m(masks, "enc_emb_" + s, (B, K, H), Y).reshape(B * K, H)
        E = torch.relu(bn) * dm * fmask.reshape(B * K, 1)
        S["xhat_" + s] = xhat; S["invstd_" + s] = invstd; S["bnpos_" + s] = (bn > 0)
        S["E_" + s] = E.view(B, K, H).transpose(0, 1).contiguous()        # (K,B,H)
        lp = enc + "lstmcell_%s" % s
        XG = S["E_" + s].reshape(K * B, H) @ P[lp + ".weight_ih"].t() + (P[lp + ".bias_ih"] + P[lp + ".bias_hh"])
        XG = XG.view(K, B, 4 * H)
        h = Y.new_zeros(B, H); c = Y.new_zeros(B, H)
        G = Y.new_zeros(K, B, 4 * H); Hs = Y.new_zeros(K, B, H); Cs = Y.new_zeros(K, B, H)
        for t in range(K):
            z = XG[t] + h @ P[lp + ".weight_hh"].t()
            i = torch.sigmoid(z[:, :H]); f = torch.sigmoid(z[:, H:2 * H])
            g = torch.tanh(z[:, 2 * H:3 * H]); o = torch.sigmoid(z[:, 3 * H:])
            mt = fm_tb[t].unsqueeze(1)
            c = (f * c + i * g) * mt
            h = o * torch.tanh(c) * mt
            G[t] = torch.cat([i, f, g, o], 1); Hs[t] = h; Cs[t] = c
        S["G_" + s] = G; S["H_" + s] = Hs; S["C_" + s] = Cs
    # ---- cross gates, batched over all frames (rows (k,b)) ----
    GG = rgb.new_zeros(K * B, 2 * H)
    for idx, (tgt, src) in enumerate((("rgb", "opfl"), ("opfl", "rgb"))):
        gp = enc + "gate_%s.gate.0" % tgt
        dm = _m(masks, "enc_gate_" + tgt, (B, K, H), rgb).transpose(0, 1).reshape(K * B, H)
        R = torch.relu(S["H_" + src].reshape(K * B, H) @ P[gp + ".weight"].t() + P[gp + ".bias"]) * dm
        S["R_" + tgt] = R
        GG[:, idx * H:(idx + 1) * H] = S["H_" + tgt].reshape(K * B, H) * (1 + R)
    S["GG"] = GG
    fp = enc + "fusion.late_fusion.0"
    Fpre = GG @ P[fp + ".weight"].t() + P[fp + ".bias"]
    Fact = O._ACT[activity](Fpre)
    dmf = _m(masks, "enc_fusion", (B, K, H), rgb).transpose(0, 1).reshape(K * B, H)
    S["Fact"] = Fact
    V = (Fact * dmf).view(K, B, H).transpose(0, 1).contiguous()          # (B,K,H)
    S["V"] = V
    Uv = V.reshape(B * K, H) @ P["lstmcore.v2a.weight"].t() + P["lstmcore.v2a.bias"]
    A = Uv.shape[1]
    S["Uv"] = Uv.view(B, K, A)
    mean = V.sum(1) / fmask.sum(1, keepdim=True)
    S["mean"] = mean
    h1 = mean @ P["img_embed_h_1.weight"].t() + P["img_embed_h_1.bias"]
    c1 = mean @ P["img_embed_c_1.weight"].t() + P["img_embed_c_1.bias"]
    h2 = mean @ P["img_embed_h_2.weight"].t() + P["img_embed_h_2.bias"]
    c2 = mean @ P["img_embed_c_2.weight"].t() + P["img_embed_c_2.bias"]

    # ---- decoder ----
    L = seq.shape[1]
    Lp = L
    for i in range(1, L):
        if int(seq[:, i].sum()) == 0:
            Lp = i
            break
    S["Lp"] = Lp
    E_ = P["embed.weight"].shape[1]
    tok = seq[:, :Lp].t().contiguous()                                   # (L',B)
    XT = P["embed.weight"][tok.reshape(-1)]                              # (L'B,E) rows (i,b)
    dg = _m(masks, "dec_gate", (Lp, B, H), rgb)[:Lp].reshape(Lp * B, H)
    RG = torch.relu(XT @ P["lstmcore.gate.gate.0.weight"].t() + P["lstmcore.gate.gate.0.bias"]) * dg
    GP = pos.repeat(Lp, 1) * (1 + RG)
    l1, l2 = "lstmcore.lstm_1", "lstmcore.lstm_2"
    Z1pre = (XT @ P[l1 + ".i2h.weight"].t() + GP @ P[l1 + ".a2h.weight"].t()
             + (P[l1 + ".i2h.bias"] + P[l1 + ".a2h.bias"] + P[l1 + ".h2h.bias"])).view(Lp, B, 4 * H)
    b2 = P[l2 + ".i2h.bias"] + P[l2 + ".a2h.bias"] + P[l2 + ".h2h.bias"]
    H12 = rgb.new_zeros(Lp + 1, B, 2 * H); C1 = rgb.new_zeros(Lp + 1, B, H); C2 = rgb.new_zeros(Lp + 1, B, H)
    H12[0, :, :H] = h1; H12[0, :, H:] = h2; C1[0] = c1; C2[0] = c2
    G1 = rgb.new_zeros(Lp, B, 4 * H); G2 = rgb.new_zeros(Lp, B, 4 * H)
    AH = rgb.new_zeros(Lp, B, A); ALPHA = rgb.new_zeros(Lp, B, K); AF = rgb.new_zeros(Lp, B, H)
    d1 = _m(masks, "dec_h1", (Lp, B, H), rgb); d2 = _m(masks, "dec_h2", (Lp, B, H), rgb)
    wa = P["lstmcore.a2w.weight"][0]; ba = P["lstmcore.a2w.bias"][0]

    def cell(z, c_prev, h_prev, m, dmask):
        sg = torch.sigmoid(z[:, :3 * H]); i, f, o = sg[:, :H], sg[:, H:2 * H], sg[:, 2 * H:]
        g = torch.tanh(z[:, 3 * H:])
        c = f * c_prev + i * g
        c = c * m + c_prev * (1 - m)
        h = o * torch.tanh(c)
        h = (h * m + h_prev * (1 - m)) * dmask
        return torch.cat([i, f, o, g], 1), c, h

    for i in range(Lp):
        m = smask[:, i].unsqueeze(1)
        AH[i] = H12[i] @ P["lstmcore.h2a.weight"].t() + P["lstmcore.h2a.bias"]
        s = (torch.tanh(AH[i].unsqueeze(1) + S["Uv"]) * wa).sum(-1) + ba     # (B,K)
        ALPHA[i] = torch.softmax(s, dim=1)
        AF[i] = (ALPHA[i].unsqueeze(-1) * V).sum(1)
        z1 = Z1pre[i] + H12[i, :, :H] @ P[l1 + ".h2h.weight"].t()
        G1[i], C1[i + 1], h1n = cell(z1, C1[i], H12[i, :, :H], m, d1[i])
        H12[i + 1, :, :H] = h1n
        z2 = h1n @ P[l2 + ".i2h.weight"].t() + AF[i] @ P[l2 + ".a2h.weight"].t() + H12[i, :, H:] @ P[l2 + ".h2h.weight"].t() + b2
        G2[i], C2[i + 1], h2n = cell(z2, C2[i], H12[i, :, H:], m, d2[i])
        H12[i + 1, :, H:] = h2n
    OUT = H12[1:, :, H:].reshape(Lp * B, H)
    logits = OUT @ P["logit.weight"].t() + P["logit.bias"]
    logp_ib = torch.log_softmax(logits, 1)                               # rows (i,b)
    dc = _m(masks, "cls", (Lp, B, P["classifer.0.bias"].shape[0]), rgb)[:Lp].reshape(Lp * B, -1)
    Hc = torch.relu(OUT @ P["classifer.0.weight"].t() + P["classifer.0.bias"]) * dc
    cat_ib = torch.log_softmax(Hc @ P["classifer.3.weight"].t() + P["classifer.3.bias"], 1)
    Vn = logits.shape[1]
    S.update(dict(tok=tok, XT=XT, RG=RG, GP=GP, H12=H12, C1=C1, C2=C2, G1=G1, G2=G2, AH=AH, ALPHA=ALPHA,
                  AF=AF, OUT=OUT, Hc=Hc, logp_ib=logp_ib, cat_ib=cat_ib, masks=masks, train=train,
                  activity=activity))
    logp = logp_ib.view(Lp, B, Vn).transpose(0, 1).contiguous()
    cat = cat_ib.view(Lp, B, -1).transpose(0, 1).contiguous()
    return logp, cat, S


def _cell_bwd(G, c_new, c_prev, dh_out, dc_carry, m, dmask, H):
    """backward of the decoder cell (order i,f,o,g; mask carries; dropout on carried h)."""
    i, f, o, g = G[:, :H], G[:, H:2 * H], G[:, 2 * H:3 * H], G[:, 3 * H:]
    dhb = dh_out * dmask
    dh_t = dhb * m
    dh_prev = dhb * (1 - m)
    tc = torch.tanh(c_new)
    do = dh_t * tc
    dcn = dc_carry + dh_t * o * (1 - tc * tc)
    dc_t = dcn * m
    dc_prev = dcn * (1 - m) + dc_t * f
    di = dc_t * g; dg_ = dc_t * i; df = dc_t * c_prev
    dz = torch.cat([di * i * (1 - i), df * f * (1 - f), do * o * (1 - o), dg_ * (1 - g * g)], 1)
    return dz, dc_prev, dh_prev


def backward(P, batch, S, dlogp: Tensor, dcat: Tensor):
    """Returns {name: grad} for the 57 learnables given d(loss)/d(logp), d(loss)/d(cat)."""
    enc = "two_spatial_encoder."
    rgb, opfl, fmask, pos, smask = (batch[k] for k in ("rgb", "opfl", "feat_mask", "pos", "seq_mask"))
    B, K = rgb.shape[0], rgb.shape[1]
    H = P["img_embed_h_1.bias"].shape[0]
    Lp = S["Lp"]; masks = S["masks"]; train = S["train"]
    A = S["Uv"].shape[2]
    Gd: Dict[str, Tensor] = {}
    l1, l2 = "lstmcore.lstm_1", "lstmcore.lstm_2"
    V = S["V"]; Uv = S["Uv"]

    # ---- heads ----
    dlp = dlogp.transpose(0, 1).reshape(Lp * B, -1)                      # rows (i,b)
    dlogits = dlp - torch.exp(S["logp_ib"]) * dlp.sum(1, keepdim=True)
    Gd["logit.weight"] = dlogits.t() @ S["OUT"]; Gd["logit.bias"] = dlogits.sum(0)
    dOUT = dlogits @ P["logit.weight"]
    dct = dcat.transpose(0, 1).reshape(Lp * B, -1)
    dcl = dct - torch.exp(S["cat_ib"]) * dct.sum(1, keepdim=True)
    Gd["classifer.3.weight"] = dcl.t() @ S["Hc"]; Gd["classifer.3.bias"] = dcl.sum(0)
    dcm = _m(masks, "cls", (Lp, B, S["Hc"].shape[1]), rgb)[:Lp].reshape(Lp * B, -1)
    dHc = (dcl @ P["classifer.3.weight"]) * dcm * (S["Hc"] > 0)
    Gd["classifer.0.weight"] = dHc.t() @ S["OUT"]; Gd["classifer.0.bias"] = dHc.sum(0)
    dOUT = (dOUT + dHc @ P["classifer.0.weight"]).view(Lp, B, H)

    # ---- decoder BPTT ----
    d1 = _m(masks, "dec_h1", (Lp, B, H), rgb); d2 = _m(masks, "dec_h2", (Lp, B, H), rgb)
    wa = P["lstmcore.a2w.weight"][0]
    dh1c = rgb.new_zeros(B, H); dc1c = rgb.new_zeros(B, H); dh2c = rgb.new_zeros(B, H); dc2c = rgb.new_zeros(B, H)
    DZ1 = rgb.new_zeros(Lp, B, 4 * H); DZ2 = rgb.new_zeros(Lp, B, 4 * H); DAH = rgb.new_zeros(Lp, B, A)
    dV = torch.zeros_like(V); dUv = torch.zeros_like(Uv)
    dwa = torch.zeros_like(wa); dba = rgb.new_zeros(())
    for i in range(Lp - 1, -1, -1):
        m = smask[:, i].unsqueeze(1)
        dz2, dc2c, dh2_prev = _cell_bwd(S["G2"][i], S["C2"][i + 1], S["C2"][i], dOUT[i] + dh2c, dc2c, m, d2[i], H)
        DZ2[i] = dz2
        dh1_new = dh1c + dz2 @ P[l2 + ".i2h.weight"]
        dAF = dz2 @ P[l2 + ".a2h.weight"]
        dh2_prev = dh2_prev + dz2 @ P[l2 + ".h2h.weight"]
        # attention backward
        al = S["ALPHA"][i]
        dal = (dAF.unsqueeze(1) * V).sum(-1)                              # (B,K)
        dV += al.unsqueeze(-1) * dAF.unsqueeze(1)
        ds = al * (dal - (al * dal).sum(1, keepdim=True))
        th = torch.tanh(S["AH"][i].unsqueeze(1) + Uv)                     # (B,K,A) recomputed
        dpre = ds.unsqueeze(-1) * wa * (1 - th * th)
        dwa += (ds.unsqueeze(-1) * th).sum((0, 1)); dba += ds.sum()
        dUv += dpre
        DAH[i] = dpre.sum(1)
        dH12 = DAH[i] @ P["lstmcore.h2a.weight"]                          # (B,2H)
        dz1, dc1c, dh1_prev = _cell_bwd(S["G1"][i], S["C1"][i + 1], S["C1"][i], dh1_new, dc1c, m, d1[i], H)
        DZ1[i] = dz1
        dh1c = dh1_prev + dz1 @ P[l1 + ".h2h.weight"] + dH12[:, :H]
        dh2c = dh2_prev + dH12[:, H:]
    # init-state linears (mean is detached: SAModel.py:59-62)
    for n, d in (("h_1", dh1c), ("c_1", dc1c), ("h_2", dh2c), ("c_2", dc2c)):
        Gd["img_embed_%s.weight" % n] = d.t() @ S["mean"]; Gd["img_embed_%s.bias" % n] = d.sum(0)
    # batched weight gradients over all steps
    Z1 = DZ1.reshape(Lp * B, 4 * H); Z2 = DZ2.reshape(Lp * B, 4 * H)
    H12 = S["H12"]
    Hprev = H12[:Lp].reshape(Lp * B, 2 * H); Hnew = H12[1:].reshape(Lp * B, 2 * H)
    Gd[l2 + ".i2h.weight"] = Z2.t() @ Hnew[:, :H]
    Gd[l2 + ".a2h.weight"] = Z2.t() @ S["AF"].reshape(Lp * B, H)
    Gd[l2 + ".h2h.weight"] = Z2.t() @ Hprev[:, H:]
    Gd[l1 + ".i2h.weight"] = Z1.t() @ S["XT"]
    Gd[l1 + ".a2h.weight"] = Z1.t() @ S["GP"]
    Gd[l1 + ".h2h.weight"] = Z1.t() @ Hprev[:, :H]
    for nm in ("i2h", "a2h", "h2h"):
        Gd[l1 + ".%s.bias" % nm] = Z1.sum(0); Gd[l2 + ".%s.bias" % nm] = Z2.sum(0)
    DAHf = DAH.reshape(Lp * B, A)
    Gd["lstmcore.h2a.weight"] = DAHf.t() @ Hprev; Gd["lstmcore.h2a.bias"] = DAHf.sum(0)
    Gd["lstmcore.a2w.weight"] = dwa.unsqueeze(0); Gd["lstmcore.a2w.bias"] = dba.reshape(1)
    dGP = Z1 @ P[l1 + ".a2h.weight"]
    dgm = _m(masks, "dec_gate", (Lp, B, H), rgb)[:Lp].reshape(Lp * B, H)
    dRG = dGP * pos.repeat(Lp, 1) * dgm * (S["RG"] > 0)
    Gd["lstmcore.gate.gate.0.weight"] = dRG.t() @ S["XT"]; Gd["lstmcore.gate.gate.0.bias"] = dRG.sum(0)
    dXT = Z1 @ P[l1 + ".i2h.weight"] + dRG @ P["lstmcore.gate.gate.0.weight"]
    dE = torch.zeros_like(P["embed.weight"])
    dE.index_add_(0, S["tok"].reshape(-1), dXT)
    Gd["embed.weight"] = dE
    dUvf = dUv.reshape(B * K, A)
    Gd["lstmcore.v2a.weight"] = dUvf.t() @ V.reshape(B * K, H); Gd["lstmcore.v2a.bias"] = dUvf.sum(0)
    dV = dV + (dUvf @ P["lstmcore.v2a.weight"]).view(B, K, H)

    # ---- encoder backward (rows (k,b)) ----
    dVkb = dV.transpose(0, 1).reshape(K * B, H)
    dmf = _m(masks, "enc_fusion", (B, K, H), rgb).transpose(0, 1).reshape(K * B, H)
    Fact = S["Fact"]; act = S["activity"]
    if act == "ReLU":
        dact = (Fact > 0).to(rgb.dtype)
    elif act == "Tanh":
        dact = 1 - Fact * Fact
    else:
        dact = Fact * (1 - Fact)
    dF = dVkb * dmf * dact
    fp = enc + "fusion.late_fusion.0"
    Gd[fp + ".weight"] = dF.t() @ S["GG"]; Gd[fp + ".bias"] = dF.sum(0)
    dGG = dF @ P[fp + ".weight"]                                          # (KB,2H)
    dH = {"rgb": rgb.new_zeros(K * B, H), "opfl": rgb.new_zeros(K * B, H)}
    for idx, (tgt, src) in enumerate((("rgb", "opfl"), ("opfl", "rgb"))):
        gp = enc + "gate_%s.gate.0" % tgt
        dg_ = dGG[:, idx * H:(idx + 1) * H]
        R = S["R_" + tgt]
        dm = _m(masks, "enc_gate_" + tgt, (B, K, H), rgb).transpose(0, 1).reshape(K * B, H)
        dH[tgt] += dg_ * (1 + R)
        dR = dg_ * S["H_" + tgt].reshape(K * B, H) * dm * (R > 0)
        Gd[gp + ".weight"] = dR.t() @ S["H_" + src].reshape(K * B, H); Gd[gp + ".bias"] = dR.sum(0)
        dH[src] += dR @ P[gp + ".weight"]
    fm_tb = fmask.t().contiguous()
    for s, x in (("rgb", rgb), ("opfl", opfl)):
        lp = enc + "lstmcell_%s" % s
        G = S["G_" + s]; Cs = S["C_" + s]; Hs = S["H_" + s]
        dHs = dH[s].view(K, B, H)
        DZ = rgb.new_zeros(K, B, 4 * H)
        dhc = rgb.new_zeros(B, H); dcc = rgb.new_zeros(B, H)
        for t in range(K - 1, -1, -1):
            i, f, g, o = G[t, :, :H], G[t, :, H:2 * H], G[t, :, 2 * H:3 * H], G[t, :, 3 * H:]
            mt = fm_tb[t].unsqueeze(1)
            dh = (dHs[t] + dhc) * mt
            tc = torch.tanh(Cs[t])
            do = dh * tc
            dc = (dcc * mt) + dh * o * (1 - tc * tc)
            c_prev = Cs[t - 1] if t > 0 else torch.zeros_like(dc)
            di = dc * g; dg_ = dc * i; df = dc * c_prev
            dcc = dc * f
            DZ[t] = torch.cat([di * i * (1 - i), df * f * (1 - f), dg_ * (1 - g * g), do * o * (1 - o)], 1)
            dhc = DZ[t] @ P[lp + ".weight_hh"]
        DZf = DZ.reshape(K * B, 4 * H)
        Gd[lp + ".weight_hh"] = DZf[B:].t() @ Hs.reshape(K * B, H)[:(K - 1) * B] if K > 1 else torch.zeros_like(P[lp + ".weight_hh"])
        Gd[lp + ".weight_ih"] = DZf.t() @ S["E_" + s].reshape(K * B, H)
        Gd[lp + ".bias_ih"] = DZf.sum(0); Gd[lp + ".bias_hh"] = DZf.sum(0)
        dE = (DZf @ P[lp + ".weight_ih"]).view(K, B, H).transpose(0, 1).reshape(B * K, H)   # rows (b,k)
        pre = enc + "visual_emb_%s" % s
        dm = _m(masks, "enc_emb_" + s, (B, K, H), rgb).reshape(B * K, H)
        dy = dE * fmask.reshape(B * K, 1) * dm * S["bnpos_" + s]
        xhat = S["xhat_" + s]; invstd = S["invstd_" + s]; gamma = P[pre + ".1.weight"]
        Gd[pre + ".1.weight"] = (dy * xhat).sum(0); Gd[pre + ".1.bias"] = dy.sum(0)
        M = B * K
        if train:
            dY = (gamma * invstd / M) * (M * dy - dy.sum(0) - xhat * (dy * xhat).sum(0))
        else:
            dY = dy * gamma * invstd
        Gd[pre + ".0.weight"] = dY.t() @ x.reshape(B * K, -1); Gd[pre + ".0.bias"] = dY.sum(0)
    return Gd
