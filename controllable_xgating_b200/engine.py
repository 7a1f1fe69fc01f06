"""Host-side engine: owns the xg_handle, binds nn.Parameter storage, caches workspaces, and wraps each
C-ABI entry with shape checks.  torch is used only for device memory, streams and autograd plumbing;
all arithmetic happens inside libxgating.so."""
from __future__ import annotations

import contextlib
import ctypes
from ctypes import c_int, c_void_p
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

PARAM_NAMES: List[str] = []


def _build_param_names() -> List[str]:
    enc = "two_spatial_encoder."
    n: List[str] = []
    for s in ("rgb", "opfl"):
        n += [enc + "visual_emb_%s.0.weight" % s, enc + "visual_emb_%s.0.bias" % s,
              enc + "visual_emb_%s.1.weight" % s, enc + "visual_emb_%s.1.bias" % s]
    for s in ("rgb", "opfl"):
        n += [enc + "lstmcell_%s.weight_ih" % s, enc + "lstmcell_%s.weight_hh" % s,
              enc + "lstmcell_%s.bias_ih" % s, enc + "lstmcell_%s.bias_hh" % s]
    for s in ("rgb", "opfl"):
        n += [enc + "gate_%s.gate.0.weight" % s, enc + "gate_%s.gate.0.bias" % s]
    n += [enc + "fusion.late_fusion.0.weight", enc + "fusion.late_fusion.0.bias"]
    for q in ("h_1", "c_1", "h_2", "c_2"):
        n += ["img_embed_%s.weight" % q, "img_embed_%s.bias" % q]
    n += ["lstmcore.gate.gate.0.weight", "lstmcore.gate.gate.0.bias"]
    for cell in ("lstm_1", "lstm_2"):
        for lin in ("i2h", "a2h", "h2h"):
            n += ["lstmcore.%s.%s.weight" % (cell, lin), "lstmcore.%s.%s.bias" % (cell, lin)]
    for lin in ("v2a", "h2a", "a2w"):
        n += ["lstmcore.%s.weight" % lin, "lstmcore.%s.bias" % lin]
    n += ["embed.weight", "logit.weight", "logit.bias",
          "classifer.0.weight", "classifer.0.bias", "classifer.3.weight", "classifer.3.bias"]
    assert len(n) == L.XG_NUM_PARAMS
    return n


PARAM_NAMES = _build_param_names()
DECODER_FIRST = PARAM_NAMES.index("img_embed_h_1.weight")     # XG_P_INIT_H1_W: first parameter of the decoder side
BN_BUFFER_NAMES = ["two_spatial_encoder.visual_emb_rgb.1.running_mean", "two_spatial_encoder.visual_emb_rgb.1.running_var",
                   "two_spatial_encoder.visual_emb_opfl.1.running_mean", "two_spatial_encoder.visual_emb_opfl.1.running_var"]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, shape: Sequence[int], dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the xgating path runs on CUDA only (there is no CPU fallback)" % (name, t.device))
    assert tuple(t.shape) == tuple(shape), "%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class Engine:
    """One per SAModel instance."""
    _live = None        # weak set of engines (FusedAdam notifies them after rewriting parameter storage)

    @classmethod
    def notify_params_changed(cls):
        for e in list(cls._live or ()):
            e._force_changed = True

    def __init__(self, module: torch.nn.Module, dims: dict):
        self.module = module
        self.dims = dims
        self.lib = L.load()
        self.handle = c_void_p()
        self._bound_key = None
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}
        self._device = None
        self._param_version = None
        self._bind_held = False
        self._engine_mode = 2
        self._strict = None          # None: the library default (XG_STRICT_PERSIST in the environment)
        self._force_changed = False
        if Engine._live is None:
            import weakref
            Engine._live = weakref.WeakSet()
        Engine._live.add(self)

    def set_engine(self, tensor_cores: bool, persistent: bool = True):
        """tensor_cores (default True): dense contractions above the size gate use the tcgen05 3xTF32 engine;
        persistent (default True): greedy decoding runs in the fused persistent word-step kernel."""
        self._engine_mode = (2 if persistent else 1) if tensor_cores else 0
        if self.handle:
            L.check(self.lib.xg_set_engine(self.handle, self._engine_mode), "xg_set_engine", self.handle)

    def set_strict(self, strict: bool = True):
        """strict: a serial loop that cannot run on its persistent kernel raises instead of quietly taking the
        per-step launches (xg_set_strict)."""
        self._strict = bool(strict)
        if self.handle:
            L.check(self.lib.xg_set_strict(self.handle, int(self._strict)), "xg_set_strict", self.handle)

    def path_counters(self) -> Tuple[int, int]:
        """(serial loops run by a persistent kernel, serial loops run as per-step launches) so far."""
        if not self.handle:
            return 0, 0
        a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
        L.check(self.lib.xg_path_counters(self.handle, ctypes.byref(a), ctypes.byref(b)), "xg_path_counters", self.handle)
        return int(a.value), int(b.value)

    def profile(self, on: bool):
        self.bind()
        L.check(self.lib.xg_profile_enable(self.handle, int(on)), "xg_profile_enable", self.handle)

    def profile_report(self) -> List[dict]:
        """per-kernel CUDA-event timings since profile(True): [{"name", "launches", "ms"}, ...]"""
        import json
        buf = ctypes.create_string_buffer(1 << 20)
        L.check(self.lib.xg_profile_report(self.handle, buf, len(buf)), "xg_profile_report", self.handle)
        return json.loads(buf.value.decode())

    # ---- handle lifecycle -------------------------------------------------------------
    def _ensure_handle(self, device: torch.device):
        if self.handle and self._device == device:
            return
        if self.handle:
            self.lib.xg_destroy(self.handle)
            self.handle = c_void_p()
        if device.type != "cuda":
            raise RuntimeError("SAModel parameters are on %s: call .cuda() first — the xgating path is CUDA-only "
                               "(sm_100a), there is no CPU fallback" % device)
        d = self.dims
        xd = L.XgDims(d["R"], d["F"], d["H"], d["E"], d["A"], d["V"], d["C"], d["cls_hidden"],
                      L.XG_ACT[d["fusion_activity"]], float(d["drop_prob"]), 1e-5, 0.1)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        L.check(self.lib.xg_create(ctypes.byref(xd), idx, ctypes.byref(self.handle)), "xg_create")
        L.check(self.lib.xg_set_engine(self.handle, self._engine_mode), "xg_set_engine", self.handle)
        if self._strict is not None:
            L.check(self.lib.xg_set_strict(self.handle, int(self._strict)), "xg_set_strict", self.handle)
        self._device = device
        self._bound_key = None
        self._ws.clear()

    def __del__(self):
        try:
            if self.handle:
                self.lib.xg_destroy(self.handle)
        except Exception:
            pass

    def params(self) -> List[torch.nn.Parameter]:
        if getattr(self, "_plist", None) is None:
            named = dict(self.module.named_parameters())
            self._plist = [named[n] for n in PARAM_NAMES]
            bufs = dict(self.module.named_buffers())
            self._blist = [bufs[n] for n in BN_BUFFER_NAMES]
        return self._plist

    def params_changed(self):
        """parameter storage was rewritten behind autograd's back (p.data writes): rebuild derived tables on next use"""
        self._force_changed = True

    def invalidate_param_cache(self):
        self._plist = None
        self._bound_key = None

    @contextlib.contextmanager
    def bound(self):
        """bind() once for a public call that enters the library several times (encoder, word loop, ...): the check walks
        57 parameters (~35 us of interpreter time), and nothing rewrites parameters between those entries."""
        if self._bind_held:
            yield
            return
        self.bind()
        self._bind_held = True
        try:
            yield
        finally:
            self._bind_held = False

    def bind(self):
        if self._bind_held:
            return
        plist = self.params()
        dev = plist[0].device
        self._ensure_handle(dev)
        key = tuple(p.data_ptr() for p in plist) + tuple(b.data_ptr() for b in self._blist)
        # In-place updates THROUGH THE PARAMETER (optimizer.step, load_state_dict, p.copy_ under no_grad) bump
        # Tensor._version: the derived copies inside the library (tf32 hi/lo splits, POS-gate token table) are then
        # dropped.  Writes through `p.data` (p.data.copy_/uniform_/add_) do NOT bump it and cannot be seen from here:
        # whoever does that calls SAModel.params_changed() (init_weights, DataParallelSAModel and FusedAdam do).
        ver = sum(p._version for p in plist)
        if key == self._bound_key:
            if ver != self._param_version or self._force_changed:
                self._force_changed = False
                L.check(self.lib.xg_params_changed(self.handle), "xg_params_changed", self.handle)
                self._param_version = ver
            return
        self._param_version = ver
        if self._bound_key is not None:
            L.check(self.lib.xg_params_changed(self.handle), "xg_params_changed", self.handle)
        for n, p in zip(PARAM_NAMES, plist):
            if p.device != dev or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameter %s must be contiguous fp32 on %s" % (n, dev))
        r, c = c_int(), c_int()
        for i, (n, p) in enumerate(zip(PARAM_NAMES, plist)):
            L.check(self.lib.xg_param_shape(self.handle, i, ctypes.byref(r), ctypes.byref(c)), "xg_param_shape", self.handle)
            assert p.numel() == r.value * c.value, "parameter %s has %d elements, expected %dx%d" % (n, p.numel(), r.value, c.value)
        table = (c_void_p * L.XG_NUM_PARAMS)(*[p.data_ptr() for p in plist])
        L.check(self.lib.xg_bind_params(self.handle, table, L.XG_NUM_PARAMS), "xg_bind_params", self.handle)
        b = self._blist
        for t in b:
            if t.device != dev or t.dtype != torch.float32:
                raise RuntimeError("BatchNorm running statistics must be fp32 on %s" % dev)
        L.check(self.lib.xg_bind_bn_buffers(self.handle, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr()),
                "xg_bind_bn_buffers", self.handle)
        self._bound_key = key

    # ---- workspaces -------------------------------------------------------------------
    def workspace(self, kind: int, B: int, K: int, LT: int = 0, beam: int = 0, fresh: bool = False) -> torch.Tensor:
        n = int(self.lib.xg_workspace_bytes(self.handle, kind, B, K, LT, beam))
        if n == 0:
            raise AssertionError("xg_workspace_bytes: unsupported shape (kind=%d B=%d K=%d L/T=%d beam=%d)" % (kind, B, K, LT, beam))
        if fresh:
            return torch.empty(n, dtype=torch.uint8, device=self._device)
        key = (kind, 0)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < n:
            ws = torch.empty(n, dtype=torch.uint8, device=self._device)
            self._ws[key] = ws
        return ws

    def next_seed(self) -> int:
        """One dropout seed per forward call, drawn from torch's CPU generator so that
        torch.manual_seed() makes runs reproducible."""
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    # ---- entries ----------------------------------------------------------------------
    def encode(self, rgb, opfl, fmask, train: bool, seed: int = 0, want_state: bool = True, want_uv: bool = True):
        self.bind()
        d = self.dims
        B, K = int(rgb.shape[0]), int(rgb.shape[1])
        rgb = _req(rgb, (B, K, d["R"]), torch.float32, "feats_rgb")
        opfl = _req(opfl, (B, K, d["F"]), torch.float32, "feats_opfl")
        fmask = _req(fmask, (B, K), torch.float32, "feat_mask")
        dev = rgb.device
        V = torch.empty(B, K, d["H"], device=dev)
        Uv = torch.empty(B, K, d["A"], device=dev) if want_uv else None
        st = [torch.empty(B, d["H"], device=dev) for _ in range(4)] if want_state else None
        ws = self.workspace(L.XG_WS_ENCODE, B, K)
        stp = (c_void_p * 4)(*[t.data_ptr() for t in st]) if st else None
        L.check(self.lib.xg_encode_fwd(self.handle, rgb.data_ptr(), opfl.data_ptr(), fmask.data_ptr(), B, K, int(train),
                                       seed, V.data_ptr(), _ptr(Uv), stp, ws.data_ptr(), ws.numel(), _stream()),
                "xg_encode_fwd", self.handle)
        return V, Uv, st

    def init_hidden(self, V, fmask):
        self.bind()
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        V = _req(V, (B, K, d["H"]), torch.float32, "feat")
        fmask = _req(fmask, (B, K), torch.float32, "feat_mask")
        st = [torch.empty(B, d["H"], device=V.device) for _ in range(4)]
        ws = self.workspace(L.XG_WS_DECODE_STEP, B, K)
        stp = (c_void_p * 4)(*[t.data_ptr() for t in st])
        L.check(self.lib.xg_init_hidden(self.handle, V.data_ptr(), fmask.data_ptr(), B, K, stp, ws.data_ptr(), ws.numel(),
                                        _stream()), "xg_init_hidden", self.handle)
        return st

    def attend_precompute(self, V):
        self.bind()
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        V = _req(V, (B, K, d["H"]), torch.float32, "feats")
        Uv = torch.empty(B, K, d["A"], device=V.device)
        L.check(self.lib.xg_attend_precompute(self.handle, V.data_ptr(), B, K, Uv.data_ptr(), _stream()),
                "xg_attend_precompute", self.handle)
        return Uv

    def decode_step(self, tokens, xt, xt_mask, V, Uv, pos, state, want_logp: bool):
        """state: 4 tensors (B,H).  Returns (out (B,H), logp or None, new_state)."""
        self.bind()
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        V = _req(V, (B, K, d["H"]), torch.float32, "feats")
        pos = _req(pos, (B, d["H"]), torch.float32, "pos_feats")
        if tokens is not None:
            tokens = _req(tokens.reshape(-1), (B,), torch.int64, "it")
        if xt is not None:
            xt = _req(xt, (B, d["E"]), torch.float32, "xt")
        if xt_mask is not None:
            xt_mask = _req(xt_mask.reshape(-1), (B,), torch.float32, "xt_mask")
        if Uv is not None:
            Uv = _req(Uv, (B, K, d["A"]), torch.float32, "Uv")
        st_in = [_req(s.reshape(B, d["H"]), (B, d["H"]), torch.float32, "state") for s in state]
        st_out = [torch.empty(B, d["H"], device=V.device) for _ in range(4)]
        logp = torch.empty(B, d["V"], device=V.device) if want_logp else None
        ws = self.workspace(L.XG_WS_DECODE_STEP, B, K)
        pin = (c_void_p * 4)(*[t.data_ptr() for t in st_in])
        pout = (c_void_p * 4)(*[t.data_ptr() for t in st_out])
        L.check(self.lib.xg_decode_step(self.handle, _ptr(tokens), _ptr(xt), _ptr(xt_mask), V.data_ptr(), _ptr(Uv),
                                        pos.data_ptr(), pin, pout, None, _ptr(logp), B, K, ws.data_ptr(), ws.numel(),
                                        _stream()), "xg_decode_step", self.handle)
        return st_out[2], logp, st_out

    def scheduled_tokens(self, V, Uv, pos, state, seq, smask, ss_prob: float, ss_seed: int, drop_seed: Optional[int]):
        """input tokens of every step under scheduled sampling (SAModel.py:89-99); no gradients."""
        self.bind()
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        Lmax = int(seq.shape[1])
        pos = _req(pos, (B, d["H"]), torch.float32, "pos_feats")
        seq = _req(seq, (B, Lmax), torch.int64, "seq")
        smask = _req(smask, (B, Lmax), torch.float32, "seq_mask")
        Lp = self.seq_steps(seq)
        out = torch.empty(B, Lmax, dtype=torch.int64, device=V.device)
        ws = self.workspace(L.XG_WS_GREEDY, B, K, Lmax)
        stp = (c_void_p * 4)(*[t.data_ptr() for t in state])
        if drop_seed is not None:
            L.check(self.lib.xg_set_decode_dropout(self.handle, 1, drop_seed), "xg_set_decode_dropout", self.handle)
        try:
            L.check(self.lib.xg_scheduled_tokens(self.handle, V.data_ptr(), _ptr(Uv), pos.data_ptr(), stp, seq.data_ptr(),
                                                 smask.data_ptr(), B, K, Lmax, Lp, float(ss_prob), ss_seed, out.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _stream()), "xg_scheduled_tokens", self.handle)
        finally:
            if drop_seed is not None:
                L.check(self.lib.xg_set_decode_dropout(self.handle, 0, 0), "xg_set_decode_dropout", self.handle)
        return out, Lp

    def sample_greedy(self, V, Uv, pos, state, T: int, sample_max: int, temperature: float, seed: int,
                      drop_seed: Optional[int] = None, want_steps: bool = True):
        """drop_seed: apply the TRAINING dropout of the word step with this Philox seed (self-critical sampling).
        want_steps=False: asynchronous call (steps_out = NULL): returns (seq, lps, None) without synchronising."""
        self.bind()
        if drop_seed is not None:
            L.check(self.lib.xg_set_decode_dropout(self.handle, 1, drop_seed), "xg_set_decode_dropout", self.handle)
            try:
                return self.sample_greedy(V, Uv, pos, state, T, sample_max, temperature, seed, want_steps=want_steps)
            finally:
                L.check(self.lib.xg_set_decode_dropout(self.handle, 0, 0), "xg_set_decode_dropout", self.handle)
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        pos = _req(pos, (B, d["H"]), torch.float32, "pos_feats")
        seq = torch.empty(B, T, dtype=torch.int64, device=V.device)
        lps = torch.empty(B, T, dtype=torch.float32, device=V.device)
        ws = self.workspace(L.XG_WS_GREEDY, B, K, T)
        stp = (c_void_p * 4)(*[t.data_ptr() for t in state])
        steps = c_int(0)
        L.check(self.lib.xg_sample_greedy(self.handle, V.data_ptr(), _ptr(Uv), pos.data_ptr(), stp, B, K, T, int(sample_max),
                                          float(temperature), seed, seq.data_ptr(), lps.data_ptr(),
                                          ctypes.byref(steps) if want_steps else None,
                                          ws.data_ptr(), ws.numel(), _stream()), "xg_sample_greedy", self.handle)
        return seq, lps, (steps.value if want_steps else None)

    def sample_beam(self, V, fmask, pos, T: int, beam: int):
        self.bind()
        d = self.dims
        B, K = int(V.shape[0]), int(V.shape[1])
        V = _req(V, (B, K, d["H"]), torch.float32, "feats")
        fmask = _req(fmask, (B, K), torch.float32, "feat_masks")
        pos = _req(pos, (B, d["H"]), torch.float32, "pos_feats")
        assert beam <= d["V"], ("lets assume this for now, otherwise this corner case causes a few headaches down the "
                                "road. can be dealt with in future if needed")
        dev = V.device
        seq = torch.empty(B, T, dtype=torch.int64, device=dev)
        lps = torch.empty(B, T, dtype=torch.float32, device=dev)
        dseq = torch.empty(B, beam, T, dtype=torch.int64, device=dev)
        dlps = torch.empty(B, beam, T, dtype=torch.float32, device=dev)
        dp = torch.empty(B, beam, dtype=torch.float32, device=dev)
        dn = torch.empty(B, dtype=torch.int32, device=dev)
        ws = self.workspace(L.XG_WS_BEAM, B, K, T, beam)
        L.check(self.lib.xg_sample_beam(self.handle, V.data_ptr(), fmask.data_ptr(), pos.data_ptr(), B, K, T, beam,
                                        seq.data_ptr(), lps.data_ptr(), dseq.data_ptr(), dlps.data_ptr(), dp.data_ptr(),
                                        dn.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "xg_sample_beam", self.handle)
        return seq, lps, dseq, dlps, dp, dn

    def seq_steps(self, seq) -> int:
        self._ensure_handle(seq.device)
        B, Lmax = int(seq.shape[0]), int(seq.shape[1])
        out = c_int(0)
        L.check(self.lib.xg_seq_steps(self.handle, seq.data_ptr(), B, Lmax, ctypes.byref(out), _stream()),
                "xg_seq_steps", self.handle)
        return out.value

    def train_fwd(self, rgb, opfl, fmask, pos, seq, smask, train: bool, seed: int, keep: bool, steps: Optional[int] = None):
        self.bind()
        d = self.dims
        B, K = int(rgb.shape[0]), int(rgb.shape[1])
        Lmax = int(seq.shape[1])
        rgb = _req(rgb, (B, K, d["R"]), torch.float32, "feats_rgb")
        opfl = _req(opfl, (B, K, d["F"]), torch.float32, "feats_opfl")
        fmask = _req(fmask, (B, K), torch.float32, "feat_mask")
        pos = _req(pos, (B, d["H"]), torch.float32, "pos_feats")
        seq = _req(seq, (B, Lmax), torch.int64, "seq")
        smask = _req(smask, (B, Lmax), torch.float32, "seq_mask")
        Lp = self.seq_steps(seq) if steps is None else int(steps)   # `steps`: loop length decided on another sequence
        dev = rgb.device
        logp = torch.empty(B, Lp, d["V"], device=dev)
        cat = torch.empty(B, Lp, d["C"], device=dev)
        saved = self.workspace(L.XG_WS_TRAIN_SAVED, B, K, Lmax, fresh=keep)
        L.check(self.lib.xg_train_fwd(self.handle, rgb.data_ptr(), opfl.data_ptr(), fmask.data_ptr(), pos.data_ptr(),
                                      seq.data_ptr(), smask.data_ptr(), B, K, Lmax, Lp, int(train), seed, logp.data_ptr(),
                                      cat.data_ptr(), saved.data_ptr(), saved.numel(), None, 0, _stream()),
                "xg_train_fwd", self.handle)
        ctx = dict(rgb=rgb, opfl=opfl, fmask=fmask, pos=pos, seq=seq, smask=smask, B=B, K=K, L=Lmax, Lp=Lp,
                   train=int(train), seed=seed, saved=saved)
        return logp, cat, ctx

    def train_bwd(self, ctx: dict, logp, cat, dlogp, dcat) -> List[torch.Tensor]:
        """Returns the 57 gradients as views into one flat fp32 buffer (PARAM_NAMES order)."""
        self.bind()
        plist = self.params()
        sizes = [p.numel() for p in plist]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=logp.device)
        views, off = [], 0
        for p, n in zip(plist, sizes):
            views.append(flat[off:off + n].view_as(p))
            off += n
        B, K, Lmax, Lp = ctx["B"], ctx["K"], ctx["L"], ctx["Lp"]
        ws = self.workspace(L.XG_WS_TRAIN_BWD, B, K, Lmax)
        table = (c_void_p * L.XG_NUM_PARAMS)(*[v.data_ptr() for v in views])
        if dlogp is not None:
            dlogp = _req(dlogp, tuple(logp.shape), torch.float32, "grad of logp")
        if dcat is not None:
            dcat = _req(dcat, tuple(cat.shape), torch.float32, "grad of cat")
        saved = ctx["saved"]
        # data-parallel overlap: an event recorded inside xg_train_bwd when the decoder-side gradients (the tail of `flat`
        # from PARAM_NAMES[DECODER_FIRST] on) are final; the gradient hook all-reduces that tail under the encoder backward
        ev = None
        if getattr(self, "split_event_wanted", False):
            ev = torch.cuda.Event()
            ev.record()                                   # materialises the cudaEvent_t; re-recorded by the library
            L.check(self.lib.xg_set_bwd_split_event(self.handle, c_void_p(ev.cuda_event)), "xg_set_bwd_split_event", self.handle)
        self.last_split = None
        L.check(self.lib.xg_train_bwd(self.handle, ctx["rgb"].data_ptr(), ctx["opfl"].data_ptr(), ctx["fmask"].data_ptr(),
                                      ctx["pos"].data_ptr(), ctx["seq"].data_ptr(), ctx["smask"].data_ptr(), B, K, Lmax, Lp,
                                      ctx["train"], ctx["seed"], logp.data_ptr(), cat.data_ptr(), _ptr(dlogp), _ptr(dcat),
                                      saved.data_ptr(), saved.numel(), table, 0, ws.data_ptr(), ws.numel(), _stream()),
                "xg_train_bwd", self.handle)
        if ev is not None:
            L.check(self.lib.xg_set_bwd_split_event(self.handle, None), "xg_set_bwd_split_event", self.handle)
            self.last_split = (ev, sum(sizes[:DECODER_FIRST]))
        self.last_flat_grad = flat
        return views


def dropout_mask(seed: int, site: str, n: int, p: float, device) -> torch.Tensor:
    """The mask (0 or 1/(1-p)) the kernels use at `site` for logical indices [0,n) — for oracle replay."""
    lib = L.load()
    out = torch.empty(n, dtype=torch.float32, device=device)
    L.check(lib.xg_debug_dropout_mask(seed, L.DROP_SITES[site], n, float(p), out.data_ptr(), _stream()),
            "xg_debug_dropout_mask")
    return out


def debug_gemm(layout: int, engine: int, A: torch.Tensor, Bm: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    lib = L.load()
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    L.check(lib.xg_debug_gemm(layout, engine, A.data_ptr(), Bm.data_ptr(), C.data_ptr(), M, N, K, _stream()), "xg_debug_gemm")
    return C
