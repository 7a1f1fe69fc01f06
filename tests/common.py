"""Shared helpers for the test-suite: golden loading, config table, tolerances."""
import contextlib
import os

import numpy as np
import torch

from oracle import xgating_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must stay in sync with tests/golden/make_golden.py:CONFIGS
CONFIGS = {
    "c1": dict(dims=dict(R=1536, F=1024, H=512, E=468, A=1536, V=1000, C=14), B=2, K=28, T=20, pseed=1024, dseed=0),
    "tiny": dict(dims=dict(R=24, F=20, H=16, E=12, A=20, V=40, C=5), B=5, K=6, T=7, pseed=7, dseed=3),
    "mid": dict(dims=dict(R=96, F=64, H=64, E=36, A=80, V=300, C=14), B=8, K=12, T=10, pseed=11, dseed=5),
}

RTOL = 1e-3     # BASELINE.json north_star: "within 1e-3 relative fp32"


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def make_case(name, dtype=torch.float32):
    cfg = CONFIGS[name]
    P = O.synth_params(cfg["dims"], cfg["pseed"], dtype)
    if name == "tiny":
        P["logit.bias"][0] = 0.02
    batch = O.synth_inputs(cfg["dims"], cfg["B"], cfg["K"], cfg["T"], cfg["dseed"], dtype)
    if name == "tiny":
        B = cfg["B"]
        batch["seq"] = torch.cat([batch["seq"], torch.zeros(B, 2, dtype=torch.long)], 1)
        batch["seq_mask"] = torch.cat([batch["seq_mask"], torch.zeros(B, 2, dtype=dtype)], 1)
    return cfg, P, batch


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, eps): the 1e-3 'relative fp32' gate, scale taken per tensor."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-12))


def fro_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# kernels that only the per-step (unfused) launches use: none of them may appear when a persistent kernel ran the loop
UNFUSED_KERNELS = {"greedy_pick", "att_fwd", "dec_cell", "enc_cell", "att_bwd", "dec_cell_bwd", "enc_cell_bwd",
                   "beam_topk", "beam_gather_state"}


@contextlib.contextmanager
def fused_path(model, expect):
    """Run the body on a STRICT handle (a loop that cannot run on its persistent kernel raises, xg_set_strict) with
    per-launch profiling on; afterwards assert that every kernel named in `expect` launched and that no per-step
    kernel did.  `expect` uses the profile names: decode_persistent, train_decode_persistent, decode_step_persistent,
    decode_bwd_persistent, encode_persistent, encode_bwd_persistent."""
    eng = model._engine
    eng.set_strict(True)
    eng.profile(True)
    fused0, unfused0 = eng.path_counters()
    try:
        yield
    finally:
        rep = eng.profile_report()
        eng.profile(False)
        eng.set_strict(False)
    names = {r["name"] for r in rep}
    for e in expect:
        assert e in names, "fused kernel %s did not launch; launched: %s" % (e, sorted(names))
    assert not (names & UNFUSED_KERNELS), "per-step kernels launched: %s" % sorted(names & UNFUSED_KERNELS)
    fused1, unfused1 = eng.path_counters()
    assert unfused1 == unfused0 and fused1 >= fused0 + len(expect), (fused0, fused1, unfused0, unfused1)
