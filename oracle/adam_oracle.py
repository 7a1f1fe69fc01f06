"""CPU restatement (numpy) of the optimizer step of the reference training loop — TEST INFRASTRUCTURE ONLY.

Only tests/ (and __graft_entry__.smoke / bench.py's cpu_baseline leg) may import this module; the product path
is the CUDA kernel behind xg_adam_step (controllable_xgating_b200/csrc/xg_optim.cuh).

Follows, in order:
  * myutils.clip_gradient (/root/reference/caption_src/myutils.py:79-85): param.grad.data.clamp_(-c, c)
  * optim.Adam(model.parameters(), lr=opt.learning_rate, weight_decay=opt.weight_decay)
    (/root/reference/caption_src/starttrain.py:76,137) — the algorithm lives in PyTorch, which the reference
    pins at 0.3.1.post3 (README.md:8-11) and which is not vendored; restated here from its published form
    (torch/optim/adam.py, v0.3.1):
        grad = grad + weight_decay * p
        exp_avg    = beta1 * exp_avg    + (1 - beta1) * grad
        exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad * grad
        denom      = sqrt(exp_avg_sq) + eps
        step_size  = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
        p          = p - step_size * exp_avg / denom
    eps_mode=1 restates the PyTorch >= 1.0 form (denom = sqrt(exp_avg_sq)/sqrt(1-beta2^t) + eps,
    step_size = lr / (1 - beta1^t)), which is what torch.optim.Adam of this container computes.
Pinning: tests/test_optim.py checks eps_mode=1 against torch.optim.Adam (CPU, float64 and float32) and eps_mode=0
against the algebraic identity with mode 1 when eps = 0.
"""
import math

import numpy as np


def clip_gradient(grads, grad_clip):
    """myutils.py:79-85"""
    return [np.clip(g, -grad_clip, grad_clip) for g in grads]


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=4e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              grad_clip=0.0, eps_mode=0):
    """One optimizer step on lists of numpy arrays (any float dtype); returns new (params, exp_avg, exp_avg_sq)."""
    if grad_clip and grad_clip > 0:
        grads = clip_gradient(grads, grad_clip)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    out_p, out_m, out_v = [], [], []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        dt = p.dtype.type
        if weight_decay != 0:
            g = g + dt(weight_decay) * p
        m = dt(beta1) * m + dt(1 - beta1) * g
        v = dt(beta2) * v + dt(1 - beta2) * g * g
        if eps_mode == 0:
            denom = np.sqrt(v) + dt(eps)
            step_size = lr * math.sqrt(bc2) / bc1
        else:
            denom = np.sqrt(v) / dt(math.sqrt(bc2)) + dt(eps)
            step_size = lr / bc1
        out_p.append(p - dt(step_size) * (m / denom))
        out_m.append(m)
        out_v.append(v)
    return out_p, out_m, out_v
