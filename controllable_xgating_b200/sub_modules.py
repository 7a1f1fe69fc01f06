"""Parameter containers with the reference's module tree (caption_src/sub_modules.py), so that
`state_dict()` keys, shapes and default initialisation are identical and checkpoints interchange
with `load_state_dict(strict=True)`.

These classes hold nn.Parameters only.  The arithmetic of the live classes
(Gate 18-47, Fusion 51-72, EncoderLstm_two_fc 78-159, LSTMCore_two_layer_gate 641-687,
two_inputs_lstmcell 732-770) is fused inside libxgating.so; the two modules the reference's callers
invoke directly (`two_spatial_encoder(...)`, `lstmcore(...)`) forward to the C ABI through the owning
SAModel's engine.  The reference's ablation variants (sub_modules.py:162-592) are dead code in the
shipped configuration and are not provided.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def to_contiguous(tensor):
    """sub_modules.py:10-14"""
    return tensor if tensor.is_contiguous() else tensor.contiguous()


class _Fused(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise NotImplementedError(
            "%s has no standalone forward in controllable_xgating_b200: it is fused into the CUDA path "
            "(xg_encode_fwd / xg_decode_step). Call it through SAModel, its two_spatial_encoder or its lstmcore."
            % type(self).__name__)


class Gate(_Fused):
    """target * (1 + dropout(relu(Linear(source))))  — sub_modules.py:18-47 (simple=True branch)."""

    def __init__(self, seed, source_size, target_size, drop_lm, simple=True):
        super().__init__()
        torch.manual_seed(seed)
        if not simple:
            raise NotImplementedError("Gate(simple=False) is never instantiated by the reference's SAModel")
        self.source_size, self.target_size, self.drop_prob_lm = source_size, target_size, drop_lm
        self.middle_size = 2 * source_size
        self.gate = nn.Sequential(nn.Linear(source_size, target_size), nn.ReLU(), nn.Dropout(drop_lm))


class Fusion(_Fused):
    """dropout(act(Linear(cat[f1, f2])))  — sub_modules.py:51-72."""

    def __init__(self, seed, feat_size1, feat_size2, fusion_size, drop_lm=0.5, activity=None):
        super().__init__()
        torch.manual_seed(seed)
        if activity not in ("ReLU", "Tanh", "Sigmoid"):
            raise ValueError("fusion_activity must be ReLU, Tanh or Sigmoid (myopts.py:29), got %r" % (activity,))
        self.feat_size1, self.feat_size2, self.fusion_size = feat_size1, feat_size2, fusion_size
        self.drop_prob_lm, self.activity = drop_lm, activity
        self.late_fusion = nn.Sequential(nn.Linear(feat_size1 + feat_size2, fusion_size), getattr(nn, activity)(),
                                         nn.Dropout(drop_lm))


class two_inputs_lstmcell(_Fused):
    """LSTM cell with two inputs, gate order i,f,o,g, mask-carried state — sub_modules.py:732-770."""

    def __init__(self, input_size, visual_size, rnn_size, drop_lm=0.5):
        super().__init__()
        self.input_size, self.visual_size, self.rnn_size, self.drop_lm = input_size, visual_size, rnn_size, drop_lm
        self.i2h = nn.Linear(input_size, 4 * rnn_size)
        self.a2h = nn.Linear(visual_size, 4 * rnn_size)
        self.h2h = nn.Linear(rnn_size, 4 * rnn_size)
        if drop_lm is not None:
            self.dropout = nn.Dropout(drop_lm)


class EncoderLstm_two_fc(nn.Module):
    """The Cross-Gating block — sub_modules.py:78-159."""

    def __init__(self, opt):
        super().__init__()
        torch.manual_seed(opt.seed)
        self.feat_size_rgb, self.feat_size_opfl = opt.feat_size, opt.feat_size2
        self.embed_size = self.rnn_size = opt.rnn_size
        self.drop_prob_lm = opt.drop_prob_lm
        H = opt.rnn_size
        self.visual_emb_rgb = nn.Sequential(nn.Linear(opt.feat_size, H), nn.BatchNorm1d(H), nn.ReLU(True))
        self.visual_emb_opfl = nn.Sequential(nn.Linear(opt.feat_size2, H), nn.BatchNorm1d(H), nn.ReLU(True))
        self.drop_out = nn.Dropout(opt.drop_prob_lm)
        self.lstmcell_rgb = nn.LSTMCell(H, H)
        self.lstmcell_opfl = nn.LSTMCell(H, H)
        self.gate_rgb = Gate(opt.seed, H, H, opt.drop_prob_lm)
        self.gate_opfl = Gate(opt.seed, H, H, opt.drop_prob_lm)
        self.fusion = Fusion(opt.seed, H, H, H, opt.drop_prob_lm, opt.fusion_activity)
        self._owner = None   # set by SAModel (plain attribute: not a sub-module, not in state_dict)

    def forward(self, feats_rgb, feats_opfl, feats_mask):
        """(m,K,R), (m,K,F), (m,K) -> fused feats (m,K,H).  Inference/no-grad use; training gradients flow
        through SAModel.forward, which runs the encoder inside the fused train path."""
        owner = object.__getattribute__(self, "_owner")
        if owner is None:
            raise RuntimeError("EncoderLstm_two_fc must be owned by an SAModel to run (it needs the xgating engine)")
        return owner()._encode(feats_rgb, feats_opfl, feats_mask)[0]


class LSTMCore_two_layer_gate(nn.Module):
    """One word step: attention + POS gate + two LSTM cells — sub_modules.py:641-687."""

    def __init__(self, opt):
        super().__init__()
        torch.manual_seed(opt.seed)
        self.input_encoding_size, self.rnn_size = opt.input_encoding_size, opt.rnn_size
        self.visual_size = self.globalpos_size = opt.rnn_size
        self.att_size, self.drop_prob_lm = opt.att_size, opt.drop_prob_lm
        H, E, A = opt.rnn_size, opt.input_encoding_size, opt.att_size
        self.gate = Gate(opt.seed, E, H, opt.drop_prob_lm)
        self.lstm_1 = two_inputs_lstmcell(E, H, H, opt.drop_prob_lm)
        self.lstm_2 = two_inputs_lstmcell(H, H, H, opt.drop_prob_lm)
        self.dropout = nn.Dropout(opt.drop_prob_lm)
        self.v2a = nn.Linear(H, A)
        self.h2a = nn.Linear(H + H, A)
        self.a2w = nn.Linear(A, 1)
        self._owner = None

    def forward(self, xt, xt_mask, V, pos_feat, state):
        """xt (m,E), xt_mask (m,1), V (m,K,H), pos_feat (m,H), state [(h1,c1),(h2,c2)] each (1,m,H)
        -> output (m,H), new state.  Eval-mode (no dropout, no grad) word step."""
        assert len(state) == 2, "input parameters 'state' expect a list with 2 elements"   # sub_modules.py:673
        owner = object.__getattribute__(self, "_owner")
        if owner is None:
            raise RuntimeError("LSTMCore_two_layer_gate must be owned by an SAModel to run")
        return owner()._word_step(None, xt, xt_mask, V, pos_feat, state, want_logp=False)[0::2]
