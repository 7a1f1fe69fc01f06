"""Fused clamp + Adam step for the training loop of the reference (starttrain.py:134-137):

    loss.backward(); myutils.clip_gradient(optimizer, opt.grad_clip); optimizer.step()

`FusedAdam` is a torch.optim.Optimizer (so `myutils.set_lr`, `param_groups`, `zero_grad`, `state_dict` keep
working) whose `step()` is ONE launch of `xg_adam_step` over every parameter tensor: gradient clamp
(`grad_clip`, myutils.py:79-85), L2 weight decay and Adam.  It replaces `optim.Adam(model.parameters(),
lr=opt.learning_rate, weight_decay=opt.weight_decay)` (starttrain.py:76); calling the reference's own
`clip_gradient` before `step()` is harmless (the clamp is idempotent).  CUDA only: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clip=0.0,
                 reference_eps=True):
        """reference_eps=True: the PyTorch 0.3.1 formula the reference pins (denom = sqrt(v) + eps);
        False: the formula of torch >= 1.0."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_clip=grad_clip)
        super().__init__(params, defaults)
        self.reference_eps = bool(reference_eps)
        self._lib = L.load()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: parameters must live on a CUDA device (there is no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            step = self.state[plist[0]]["step"] + 1
            stream = torch.cuda.current_stream(plist[0].device).cuda_stream
            b1, b2 = group["betas"]
            for i in range(0, len(plist), 64):
                chunk = plist[i:i + 64]
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                table = (L.XgAdamTensor * len(chunk))()
                for k, (p, g) in enumerate(zip(chunk, grads)):
                    st = self.state[p]
                    table[k] = L.XgAdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                              p.numel())
                L.check(self._lib.xg_adam_step(table, len(chunk), step, float(group["lr"]), float(b1), float(b2),
                                               float(group["eps"]), float(group["weight_decay"]), float(group["grad_clip"]),
                                               0 if self.reference_eps else 1, 1, ctypes.c_void_p(stream)), "xg_adam_step")
            for p in plist:
                self.state[p]["step"] = step
        # the kernel rewrites parameter storage behind autograd's version counters: tell every engine to drop
        # the tables it derived from the parameters (tf32 splits, POS-gate token table)
        from .engine import Engine
        Engine.notify_params_changed()
        return loss
