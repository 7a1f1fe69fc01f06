"""Per-tensor gradient errors of the full-size training step (config 3, dropout 0) against the oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import xgating_oracle as O
from tests.common import fro_err
from tests.test_gpu_parity import _full_case, build_model, dev, _xg
X = _xg()
cfg, P, b = _full_case(64); d = dev(b)
m = build_model(cfg, P, drop=0.0).train()
logp, cat = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
loss = X.LanguageModelCriterion()(logp, d["seq"], d["seq_mask"])
loss.backward()
loss_o, grads_o = O.train_step_grads(P, b, train=True)
print("loss", float(loss), float(loss_o))
named = dict(m.named_parameters())
rows = []
for n, ref in grads_o.items():
    ref = np.asarray(ref.numpy(), dtype=np.float64)
    if np.linalg.norm(ref) < 1e-7:
        continue
    rows.append((fro_err(named[n].grad.cpu().numpy(), ref), n, float(np.linalg.norm(ref))))
rows.sort(reverse=True)
for e, n, rn in rows[:8]:
    print("%.3e  %-50s |ref| %.3e" % (e, n, rn))
