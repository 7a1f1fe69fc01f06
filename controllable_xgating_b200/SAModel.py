"""Drop-in mirror of the reference's caption_src/SAModel.py surface on top of libxgating.so.

Same constructor (`SAModel(opt)`), methods, attribute names and `state_dict` keys as the reference,
so `starttrain.py` / `eval.py` / `eval_utils.py` keep working when this module is the one named
`SAModel` on sys.path (see controllable_xgating_b200/compat/).  The bodies are different: every
method hands device pointers to the CUDA library; nothing here computes with torch ops.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import _lib as L
from .CaptionModel import CaptionModel
from .engine import Engine, _stream
from .sub_modules import (EncoderLstm_two_fc, Fusion, Gate, LSTMCore_two_layer_gate, to_contiguous,  # noqa: F401
                          two_inputs_lstmcell)

__all__ = ["SAModel", "LanguageModelCriterion", "ClassiferCriterion", "RewardCriterion", "to_contiguous",
           "CaptionModel", "Gate", "Fusion", "EncoderLstm_two_fc", "LSTMCore_two_layer_gate", "two_inputs_lstmcell"]

VERBOSE = True   # default of SAModel.VERBOSE: the reference prints 'sampling with greedy search' / '... beam search' (SAModel.py:131,177)


class _TrainForward(torch.autograd.Function):
    """autograd node of the fused teacher-forced forward; backward = hand-written BPTT (xg_train_bwd)."""

    @staticmethod
    def forward(ctx, model, rgb, opfl, fmask, pos, seq, smask, seed, *params):
        eng = model._engine
        eng.split_event_wanted = bool(getattr(model._grad_hook, "overlap", False))
        logp, cat, c = eng.train_fwd(rgb, opfl, fmask, pos, seq, smask, model._train_flags(), seed, keep=True,
                                     steps=getattr(model, "_forced_steps", None))
        ctx.model = model
        ctx.c = c
        ctx.save_for_backward(logp, cat)
        ctx.set_materialize_grads(False)
        return logp, cat

    @staticmethod
    def backward(ctx, dlogp, dcat):
        logp, cat = ctx.saved_tensors
        model = ctx.model
        if ctx.c is None:
            raise RuntimeError("xgating: backward through the same forward twice is not supported "
                               "(the saved activations are consumed in place)")
        grads = model._engine.train_bwd(ctx.c, logp, cat, dlogp, dcat)
        ctx.c = None
        hook = model._grad_hook
        if hook is not None:
            hook(model._engine.last_flat_grad, model._engine.last_split)
        plist = model._engine.params()
        out = [g if p.requires_grad else None for g, p in zip(grads, plist)]
        return (None,) * 8 + tuple(out)


class PendingSample:
    """Result of SAModel.sample_async: full-length (B, seq_length) device tensors whose trailing columns are zero."""

    def __init__(self, seq, lps):
        self.seq, self.lps = seq, lps
        self.event = torch.cuda.Event()
        self.event.record()

    @staticmethod
    def steps_of(seq_host) -> int:
        """columns the reference would have returned: leading columns with a non-zero id (SAModel.py:199-210)"""
        alive = (seq_host != 0).any(dim=0)
        n = int(alive.shape[0])
        for t in range(n):
            if not bool(alive[t]):
                return t
        return n

    def to_host(self, host_seq, host_lps):
        """queue the copies into (pinned) host buffers on the current stream; finish with .event / a stream sync"""
        host_seq.copy_(self.seq, non_blocking=True)
        host_lps.copy_(self.lps, non_blocking=True)

    def result(self):
        self.event.synchronize()
        steps = self.steps_of(self.seq.cpu())
        if steps == 0:
            raise ValueError("torch.cat(): expected a non-empty list of Tensors (every caption ended at the first step)")
        return self.seq[:, :steps], self.lps[:, :steps]


class SAModel(CaptionModel):
    VERBOSE = VERBOSE   # set SAModel.VERBOSE = False (or on an instance) to silence the reference-style prints

    def __init__(self, opt):
        super(SAModel, self).__init__()
        torch.manual_seed(opt.seed)                                  # SAModel.py:16
        self.vocab_size = opt.vocab_size
        self.category_size = opt.category_size
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_size = opt.rnn_size
        self.visual_size = opt.rnn_size
        self.num_layers = opt.num_layers
        self.drop_prob_lm = opt.drop_prob_lm
        self.seq_length = opt.seq_length
        self.ss_prob = 0.0                                           # written by starttrain.py:100
        self.two_spatial_encoder = EncoderLstm_two_fc(opt)
        self.img_embed_h_1 = nn.Linear(self.visual_size, self.rnn_size)
        self.img_embed_c_1 = nn.Linear(self.visual_size, self.rnn_size)
        self.img_embed_h_2 = nn.Linear(self.visual_size, self.rnn_size)
        self.img_embed_c_2 = nn.Linear(self.visual_size, self.rnn_size)
        self.lstmcore = LSTMCore_two_layer_gate(opt)
        self.embed = nn.Embedding(self.vocab_size, self.input_encoding_size)
        self.logit = nn.Linear(self.rnn_size, self.vocab_size)
        self.classifer = nn.Sequential(nn.Linear(self.rnn_size, 128), nn.ReLU(), nn.Dropout(self.drop_prob_lm),
                                       nn.Linear(128, self.category_size))
        self.init_weights()
        dims = dict(R=opt.feat_size, F=opt.feat_size2, H=opt.rnn_size, E=opt.input_encoding_size, A=opt.att_size,
                    V=opt.vocab_size, C=opt.category_size, cls_hidden=128,
                    fusion_activity=getattr(opt, "fusion_activity", "ReLU"), drop_prob=opt.drop_prob_lm)
        object.__setattr__(self, "_engine", Engine(self, dims))
        object.__setattr__(self, "_grad_hook", None)
        ref = weakref.ref(self)
        object.__setattr__(self.two_spatial_encoder, "_owner", ref)
        object.__setattr__(self.lstmcore, "_owner", ref)
        self.done_beams = []

    def init_weights(self):
        """SAModel.py:52-56"""
        initrange = 0.1
        self.embed.weight.data.uniform_(-initrange, initrange)
        self.logit.bias.data.fill_(0)
        self.logit.weight.data.uniform_(-initrange, initrange)
        self.params_changed()

    def params_changed(self):
        """Tell the engine that parameter storage was rewritten through `.data` (p.data.copy_/uniform_/add_, an
        old-style optimizer, dist.broadcast(p.data)): such writes do not bump Tensor._version, so the tables the
        library derives from the parameters (tf32 weight splits, POS-gate token table) would otherwise go stale.
        In-place updates through the parameter itself (torch.optim, load_state_dict, FusedAdam) are detected."""
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.params_changed()

    def __deepcopy__(self, memo):
        """a copy gets its own Engine (and native handle): two Python objects never share / double-free one xg_handle"""
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in ("_engine", "_grad_hook"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        object.__setattr__(new, "_engine", Engine(new, dict(self._engine.dims)))
        new._engine._engine_mode = self._engine._engine_mode
        new._engine._strict = self._engine._strict
        object.__setattr__(new, "_grad_hook", None)
        ref = weakref.ref(new)
        object.__setattr__(new.two_spatial_encoder, "_owner", ref)
        object.__setattr__(new.lstmcore, "_owner", ref)
        return new

    # nn.Module plumbing: parameter storage can move (.cuda(), .to()); re-bind lazily
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine.invalidate_param_cache()
        return r

    def load_state_dict(self, state_dict, strict=True, **kw):
        """PyTorch-0.3.1 checkpoints have no BatchNorm `num_batches_tracked` entries; synthesise them so
        `strict=True` interchange works in both directions."""
        sd = dict(state_dict)
        for k, v in self.state_dict().items():
            if k.endswith("num_batches_tracked") and k not in sd:
                sd[k] = v
        r = super().load_state_dict(sd, strict=strict, **kw)
        self._engine.invalidate_param_cache()
        return r

    # ---- helpers ------------------------------------------------------------------------
    def _encode(self, feats_rgb, feats_opfl, feat_mask, want_state=True):
        seed = self._engine.next_seed() if (self.training and self.drop_prob_lm > 0) else 0
        with torch.no_grad():
            V, Uv, st = self._engine.encode(feats_rgb, feats_opfl, feat_mask, self.training, seed, want_state)
        if self.training:
            self._bump_bn_counters()
        return V, Uv, st

    def _train_flags(self):
        """`train` argument of xg_train_fwd: bit 0 training mode, bit 1 keep the BatchNorm running statistics
        (set while sample() replays a batch it has already encoded in training mode)."""
        return (1 if self.training else 0) | (2 if getattr(self, "_bn_frozen", False) else 0)

    def _bump_bn_counters(self):
        enc = self.two_spatial_encoder
        enc.visual_emb_rgb[1].num_batches_tracked += 1
        enc.visual_emb_opfl[1].num_batches_tracked += 1

    @staticmethod
    def _pack_state(st):
        return [(st[0].unsqueeze(0), st[1].unsqueeze(0)), (st[2].unsqueeze(0), st[3].unsqueeze(0))]

    @staticmethod
    def _unpack_state(state):
        return [state[0][0][-1], state[0][1][-1], state[1][0][-1], state[1][1][-1]]

    def _word_step(self, it, xt, xt_mask, feats, pos_feats, state, want_logp, Uv=None):
        with torch.no_grad():
            out, logp, st = self._engine.decode_step(it, xt, xt_mask, feats, Uv, pos_feats, self._unpack_state(state),
                                                     want_logp)
        return out, logp, self._pack_state(st)

    # ---- reference surface --------------------------------------------------------------
    def init_hidden(self, feat, feat_mask):
        """SAModel.py:58-65 -> [(h1,c1),(h2,c2)], each (1,m,rnn_size).  The mean over frames is detached
        in the reference (numpy round trip), so no graph is attached here either."""
        with torch.no_grad():
            st = self._engine.init_hidden(feat, feat_mask)
        return self._pack_state(st)

    def forward(self, feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask):
        """SAModel.py:67-115 -> (word log-probs (m,L',V), category log-probs (m,L',C))."""
        if self.training and self.ss_prob > 0.0 and not getattr(self, "_bn_frozen", False):
            return self._forward_scheduled(feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask)
        seed = getattr(self, "_forced_seed", None)
        if seed is None:
            seed = self._engine.next_seed() if (self.training and self.drop_prob_lm > 0) else 0
        plist = self._engine.params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
            logp, cat = _TrainForward.apply(self, feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask, seed, *plist)
        else:
            logp, cat, _ = self._engine.train_fwd(feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask,
                                                  self._train_flags(), seed, keep=False,
                                                  steps=getattr(self, "_forced_steps", None))
        if self.training and not getattr(self, "_bn_frozen", False):
            self._bump_bn_counters()
        return logp, cat

    def get_logprobs_state(self, it, feats, pos_feats, state):
        """SAModel.py:117-127"""
        _, logp, st = self._word_step(it, None, None, feats, pos_feats, state, want_logp=True)
        return logp, st

    def sample_beam(self, feats, feat_masks, pos_feats, opt={}):
        """SAModel.py:129-161 (+ CaptionModel.beam_search), the whole batch in one device call.
        Returns CPU tensors like the reference (seq (m,T) int64, seqLogprobs (m,T) float)."""
        beam_size = opt.get("beam_size", 5)
        if self.VERBOSE:
            print("sampling with beam search ( beam_size = {} )".format(beam_size))
        assert beam_size <= self.vocab_size, ("lets assume this for now, otherwise this corner case causes a few headaches "
                                              "down the road. can be dealt with in future if needed")
        with torch.no_grad():
            dev_out = self._engine.sample_beam(feats, feat_masks, pos_feats, self.seq_length, beam_size)
        # six results -> pinned host buffers with ONE synchronisation (six pageable .cpu() copies cost 0.5 ms per batch)
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in dev_out]
        for h, t in zip(host, dev_out):
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream(dev_out[0].device).synchronize()
        seq, lps, dseq, dlps, dp, dn = host
        # done_beams (SAModel.py:158: one list of {seq, logps, p} per video) is built when it is first read: the 2 x B x beam
        # row views + dicts cost a few hundred microseconds of interpreter time per batch
        self._done_beams = None
        self._done_raw = (dseq, dlps, dp, dn)
        return seq, lps

    @property
    def done_beams(self):
        if self._done_beams is None and getattr(self, "_done_raw", None) is not None:
            dseq, dlps, dp, dn = self._done_raw
            sq = [r.unbind(0) for r in dseq.unbind(0)]
            lp = [r.unbind(0) for r in dlps.unbind(0)]
            dpl, dnl = dp.tolist(), dn.tolist()
            self._done_beams = [[{"seq": sq[k][j], "logps": lp[k][j], "p": dpl[k][j]} for j in range(dnl[k])]
                                for k in range(dseq.size(0))]
        return self._done_beams

    @done_beams.setter
    def done_beams(self, value):
        object.__setattr__(self, "_done_beams", value)
        object.__setattr__(self, "_done_raw", None)

    def sample(self, feats_rgb, feats_opfl, feat_mask, pos_feats, opt={}):
        """SAModel.py:163-219"""
        sample_max = opt.get("sample_max", 1)
        beam_size = opt.get("beam_size", 1)
        temperature = opt.get("temperature", 1.0)
        if self.training:
            if beam_size > 1:
                # SAModel.py:169-175 under model.train(): the encoder runs in training mode (batch-statistics BatchNorm,
                # dropout, running statistics updated); the beam search then walks the word steps.  The word steps here
                # are the deterministic ones (the reference also drops units inside every step with torch's RNG, which
                # no other implementation can reproduce mask for mask).
                feats, _, _ = self._encode(feats_rgb, feats_opfl, feat_mask, want_state=False)
                return self.sample_beam(feats, feat_mask, pos_feats, opt)
            return self._sample_training(feats_rgb, feats_opfl, feat_mask, pos_feats, sample_max, temperature)
        with self._engine.bound():
            feats, Uv, st = self._encode(feats_rgb, feats_opfl, feat_mask)
            if beam_size > 1:
                return self.sample_beam(feats, feat_mask, pos_feats, opt)
            if self.VERBOSE:
                print("sampling with greedy search")
            seed = 0 if sample_max else self._engine.next_seed()
            with torch.no_grad():
                seq, lps, steps = self._engine.sample_greedy(feats, Uv, pos_feats, st, self.seq_length, sample_max,
                                                             temperature, seed)
        if steps == 0:
            # the reference crashes here (torch.cat of an empty list, SAModel.py:219)
            raise ValueError("torch.cat(): expected a non-empty list of Tensors (every caption ended at the first step)")
        return seq[:, :steps], lps[:, :steps]


    def sample_async(self, feats_rgb, feats_opfl, feat_mask, pos_feats, opt={}):
        """Greedy / multinomial sample() in eval mode WITHOUT the host synchronisation: the work is queued on the current
        stream and a PendingSample comes back at once, so a serving loop can queue batch i+1 before it reads batch i
        (sample() has to wait for the word loop to learn how many columns the reference would return, SAModel.py:206).
        PendingSample.result() -> (seq, seqLogprobs) exactly as sample() returns them."""
        if self.training or opt.get("beam_size", 1) > 1:
            raise ValueError("sample_async covers eval-mode greedy / multinomial decoding (use sample())")
        sample_max = opt.get("sample_max", 1)
        with self._engine.bound():
            feats, Uv, st = self._encode(feats_rgb, feats_opfl, feat_mask)
            seed = 0 if sample_max else self._engine.next_seed()
            with torch.no_grad():
                seq, lps, _ = self._engine.sample_greedy(feats, Uv, pos_feats, st, self.seq_length, sample_max,
                                                         opt.get("temperature", 1.0), seed, want_steps=False)
        return PendingSample(seq, lps)

    def _forward_scheduled(self, feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask):
        """forward() with scheduled sampling (ss_prob > 0, SAModel.py:89-99).  The sampled inputs carry no gradient
        in the reference either, so: (1) encoder once in training mode; (2) a token pass walks the word loop
        without gradients and decides every step's input token (ground truth, or a draw from the previous step's
        distribution with probability ss_prob); (3) the teacher-forced forward on THOSE tokens with the same
        dropout seed and frozen running statistics gives the log-probs and their graph."""
        eng = self._engine
        seed = eng.next_seed()
        ss_seed = eng.next_seed()
        with torch.no_grad():
            feats, Uv, st = eng.encode(feats_rgb, feats_opfl, feat_mask, True, seed, True)
            self._bump_bn_counters()
            used, Lp = eng.scheduled_tokens(feats, Uv, pos_feats, st, seq, seq_mask, self.ss_prob, ss_seed, drop_seed=seed)
        self._last_ss_tokens = used                       # diagnostics / tests
        self._forced_seed, self._bn_frozen, self._forced_steps = seed, True, Lp
        try:
            return self.forward(feats_rgb, feats_opfl, feat_mask, pos_feats, used, seq_mask)
        finally:
            self._forced_seed, self._bn_frozen, self._forced_steps = None, False, None

    def _sample_training(self, feats_rgb, feats_opfl, feat_mask, pos_feats, sample_max, temperature):
        """sample() under model.train() — the self-critical path (starttrain.py:131, myutils.py:41-77): tokens are
        drawn with dropout and batch-statistics BatchNorm active, and the returned log-probs carry gradients.

        1. the encoder runs ONCE in training mode (running statistics updated once, like the reference);
        2. the word loop samples with the training dropout of the step (Philox seed s);
        3. a teacher-forced forward on the sampled tokens with the SAME seed (running statistics frozen) reproduces
           those activations and provides the autograd graph: seqLogprobs = gather(logp, seq)."""
        eng = self._engine
        seed = eng.next_seed()
        with torch.no_grad():
            feats, Uv, st = eng.encode(feats_rgb, feats_opfl, feat_mask, True, seed, True)
            self._bump_bn_counters()
            sseed = 0 if sample_max else eng.next_seed()
            seq, lps, steps = eng.sample_greedy(feats, Uv, pos_feats, st, self.seq_length, sample_max, temperature, sseed,
                                                drop_seed=seed)
        if steps == 0:
            raise ValueError("torch.cat(): expected a non-empty list of Tensors (every caption ended at the first step)")
        seq = seq[:, :steps]
        B = seq.size(0)
        if not torch.is_grad_enabled():
            # the greedy baseline of get_self_critical_reward (myutils.py:45) runs under no_grad: nothing to replay
            return seq, lps[:, :steps]
        # inputs of the steps: <bos>, then the tokens just sampled; the state mask of step t >= 1 is `unfinished`
        seq_in = torch.cat([seq.new_zeros(B, 1), seq[:, :steps - 1]], 1)
        mask = torch.cat([torch.ones(B, 1, device=seq.device), (seq[:, :steps - 1] > 0).float()], 1)
        self._forced_seed, self._bn_frozen = seed, True
        try:
            logp, _ = self.forward(feats_rgb, feats_opfl, feat_mask, pos_feats, seq_in, mask)
        finally:
            self._forced_seed, self._bn_frozen = None, False
        lps = lps[:, :steps]
        self._last_sample_logprobs = lps                 # from the sampling pass itself (diagnostics / tests)
        gathered = logp[:, :steps].gather(2, seq.unsqueeze(2)).squeeze(2)
        # a caption that finished before step t keeps recording the log-prob of its RAW sample (SAModel.py:209-210)
        # while seq holds 0 there; those positions are exactly the ones RewardCriterion masks out (SAModel.py:262)
        live = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=seq.device), seq[:, :steps - 1] > 0], 1)
        return seq, torch.where(live, gathered, lps)


# ------------------------------------------------------------------------------------------
# criterions (SAModel.py:221-267)
# ------------------------------------------------------------------------------------------
class _NLLCriterion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp, target, mask, class_mask, rotate):
        if not logp.is_cuda:
            raise RuntimeError("criterion input is on %s: the xgating path is CUDA-only" % logp.device)
        lib = L.load()
        B, Lp, N = logp.shape
        logp_c = logp.contiguous()
        ld = int(target.shape[1])
        assert ld >= Lp and mask.shape[1] == ld, "target / mask must have at least as many columns as the log-probs"
        target_c = target.contiguous().to(torch.int64)
        mask_c = mask.contiguous().to(torch.float32)
        cm_c = None if class_mask is None else class_mask.contiguous().to(torch.float32)
        out = torch.empty(2, device=logp.device)
        scratch = torch.empty(2 * B * Lp, device=logp.device)
        L.check(lib.xg_nll_criterion_fwd(logp_c.data_ptr(), N, target_c.data_ptr(), mask_c.data_ptr(),
                                         None if cm_c is None else cm_c.data_ptr(), ld, int(rotate), B, Lp,
                                         out.data_ptr(), out[1:].data_ptr(), scratch.data_ptr(), _stream()),
                "xg_nll_criterion_fwd")
        ctx.save_for_backward(target_c, mask_c, out)
        ctx.cm = cm_c
        ctx.meta = (B, Lp, N, ld, int(rotate))
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        target_c, mask_c, out = ctx.saved_tensors
        B, Lp, N, ld, rotate = ctx.meta
        lib = L.load()
        g = g.contiguous().to(torch.float32)
        dlogp = torch.empty(B, Lp, N, device=out.device)
        L.check(lib.xg_nll_criterion_bwd(N, target_c.data_ptr(), mask_c.data_ptr(),
                                         None if ctx.cm is None else ctx.cm.data_ptr(), ld, rotate, B, Lp,
                                         out[1:].data_ptr(), g.data_ptr(), dlogp.data_ptr(), _stream()),
                "xg_nll_criterion_bwd")
        return dlogp, None, None, None, None


class LanguageModelCriterion(nn.Module):
    """SAModel.py:221-234: target rotated left by one; -sum(logp[target]*mask)/sum(mask)."""

    def forward(self, input, target, mask):
        return _NLLCriterion.apply(input, target, mask, None, True)


class ClassiferCriterion(nn.Module):
    """SAModel.py:236-253"""

    def forward(self, input, target, mask, class_mask=None):
        return _NLLCriterion.apply(input, target, mask, class_mask, False)


class RewardCriterion(nn.Module):
    """SAModel.py:255-267 (SCST loss; adjacent to the measured path, a (m,T) elementwise expression)."""

    def forward(self, input, seq, reward):
        input = to_contiguous(input).view(-1)
        reward = to_contiguous(reward).view(-1)
        mask = (seq > 0).float()
        mask = to_contiguous(torch.cat([mask.new_ones(mask.size(0), 1), mask[:, :-1]], 1)).view(-1)
        output = -input * reward * mask
        return torch.sum(output) / torch.sum(mask)
