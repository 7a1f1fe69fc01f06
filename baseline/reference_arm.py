"""The reference's OWN files as the CPU arm of bench.py (`--impl reference`, cpu_baseline.kind = "reference").

`stage()` (called by __graft_entry__.build() in the build container, where /root/reference exists) copies the five
modules the path needs — caption_src/{SAModel,sub_modules,CaptionModel,data_io,myopts}.py, unmodified — into
baseline/_ref/caption_src/.  That directory is git-ignored (the reference's sources never enter the history) but it
is NOT gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  The reference is
un-packaged Python-2.7 / PyTorch-0.3.1 source: there is nothing to pip-install; it imports under torch 2.x with the
three shims below, none of which touches the arithmetic.

Test / bench infrastructure only: nothing under controllable_xgating_b200/ imports this module.
"""
import argparse
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "caption_src")
SRC_DIR = "/root/reference/caption_src"
FILES = ("SAModel.py", "sub_modules.py", "CaptionModel.py", "data_io.py", "myopts.py")


def stage(verbose=False):
    """copy the reference's files next to the bench (build container only); returns True if baseline/_ref is usable"""
    if os.path.isdir(SRC_DIR):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            src, dst = os.path.join(SRC_DIR, f), os.path.join(REF_DIR, f)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copyfile(src, dst)
                if verbose:
                    print("[reference arm] staged", dst)
    return available()


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def load():
    """import the staged reference as a module (its file is called SAModel.py and star-imports its siblings)"""
    import torch
    if not available():
        raise ImportError("baseline/_ref is not staged (run __graft_entry__.build() in the build container)")
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))             # data_io.py:16 imports it; never used here
    torch.Tensor.cuda = lambda self, *a, **k: self                        # hard-coded .cuda() (SAModel.py:62,121,...)
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _narrow = torch.Tensor.narrow

    def narrow(self, *a, **k):                                            # torch-0.3 keyword (sub_modules.py:753)
        if "dimension" in k:
            k["dim"] = k.pop("dimension")
        return _narrow(self, *a, **k)
    torch.Tensor.narrow = narrow
    for m in ("SAModel", "sub_modules", "CaptionModel", "data_io", "myopts"):
        sys.modules.pop(m, None)
    sys.path.insert(0, REF_DIR)
    argv, sys.argv = sys.argv, ["x"]                                      # myopts.parse_opt() reads sys.argv
    try:
        import SAModel as RS
    finally:
        sys.argv = argv
        sys.path.remove(REF_DIR)
    return RS


def build_model(RS, dims, T, drop, P):
    opt = argparse.Namespace(vocab_size=dims["V"], category_size=dims["C"], input_encoding_size=dims["E"],
                             rnn_size=dims["H"], num_layers=1, drop_prob_lm=drop, seq_length=T, seed=1024,
                             feat_size=dims["R"], feat_size2=dims["F"], att_size=dims["A"], fusion_activity="ReLU")
    model = RS.SAModel(opt)
    sd = model.state_dict()
    model.load_state_dict({k: (P[k].clone() if k in P else v) for k, v in sd.items()}, strict=True)
    return model


def greedy(model, batch):
    """SAModel.sample (SAModel.py:163-219), eval mode, greedy"""
    import torch
    model.eval()
    with torch.no_grad():
        return model.sample(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], {"sample_max": 1, "beam_size": 1})


class legacy_scalar_indexing:
    """PyTorch 0.3.1 returned a Python number (a COPY) from integer indexing that yields one element; torch >= 0.4
    returns a 0-dim VIEW, which CaptionModel.py:114-118 then overwrites with -1000.  Restores the 0.3.1 behaviour the
    reference was written for while its beam search runs (same shim as tests/golden/make_golden.py)."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.__getitem__
        orig = self._orig

        def getitem(t, idx):
            r = orig(t, idx)
            if isinstance(r, torch.Tensor) and r.dim() == 0 and not r.requires_grad:
                return r.item()
            return r
        torch.Tensor.__getitem__ = getitem

    def __exit__(self, *a):
        import torch
        torch.Tensor.__getitem__ = self._orig


def beam(model, batch, beam_size, rows=None):
    """SAModel.sample with beam_size > 1 (-> sample_beam + CaptionModel.beam_search), on the first `rows` videos"""
    import torch
    model.eval()
    sl = slice(0, rows)
    with torch.no_grad(), legacy_scalar_indexing():
        return model.sample(batch["rgb"][sl], batch["opfl"][sl], batch["feat_mask"][sl], batch["pos"][sl], {"beam_size": beam_size})


def train_step(model, RS, batch):
    """forward + LanguageModelCriterion + backward (starttrain.py:125-134)"""
    model.train()
    model.zero_grad()
    logp, _ = model(batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], batch["seq"], batch["seq_mask"])
    Lp = logp.shape[1]
    loss = RS.LanguageModelCriterion()(logp, batch["seq"][:, :Lp], batch["seq_mask"][:, :Lp])
    loss.backward()
    return float(loss)
