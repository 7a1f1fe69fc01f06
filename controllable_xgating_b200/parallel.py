"""Batch-axis data parallelism for the training path (the reference is single-GPU).

One process per GPU.  Captions are independent in forward / decode / beam search, so those need no
communication: each rank runs its own contiguous batch shard.  Training needs exactly one exchange:
a sum-allreduce of the flat fp32 gradient buffer the backward pass fills (26.3 M floats at V=10k),
issued from inside the backward hook before autograd hands the views to `.grad`.

Loss normalisation: `LanguageModelCriterion` divides by the LOCAL sum(mask) (SAModel.py:233).  With
`exact=True` each replica's gradient is rescaled by local_mask_sum / global_mask_sum before the sum, so
the result equals the single-device gradient of the loss over the concatenated batch (up to BatchNorm,
whose train-mode statistics stay per replica).  With `exact=False` gradients are averaged (DDP style).
Elementwise clamping (`myutils.clip_gradient`) happens after the reduction, as in the reference.
The `exact` rescale knows two normalisers: sum(seq_mask) of forward() (XE loss) and the sampled-sequence mask of
sample() (RewardCriterion).  A loss that mixes in ClassiferCriterion with weight_class > 0 has a second normaliser,
sum(mask * class_mask), for its own term; pass `exact=False` (plain averaging) or call `hook.set_mask()` yourself then.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, world: int, rank: int) -> dict:
    n = next(v for v in batch.values() if isinstance(v, torch.Tensor)).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return {k: (v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


class GradAllReduce:
    """Callable installed as `model._grad_hook`: reduces the flat gradient buffer in place."""

    def __init__(self, group=None, exact: bool = True, overlap: bool = True):
        self.group = group
        self.exact = exact
        self.overlap = overlap          # all-reduce the decoder-side tail of the buffer under the encoder backward
        self._side = None
        self.local_mask_sum: Optional[torch.Tensor] = None
        self.calls = 0
        self.bytes = 0

    def set_mask(self, seq_mask: torch.Tensor):
        self.local_mask_sum = seq_mask.sum().reshape(1).to(torch.float32)
        self._mask_event = None
        if self.local_mask_sum.is_cuda:
            self._mask_event = torch.cuda.Event()
            self._mask_event.record()                     # the side stream waits for this sum, not for the backward pass

    def __call__(self, flat: torch.Tensor, split=None):
        """`split` = (event, n_head) from the engine: the event fires inside the backward pass once flat[n_head:] (the
        decoder-side gradients, 3/4 of the buffer) is final; that tail is then scaled and all-reduced on a side stream
        while the encoder backward still runs, and only the head waits for the end of the pass.  Still ONE logical
        all-reduce of the flat buffer per step (two NCCL calls on disjoint halves)."""
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        if self.exact:
            if self.local_mask_sum is None:
                raise RuntimeError("GradAllReduce(exact=True): call set_mask(seq_mask) before backward")
            loc = self.local_mask_sum.to(flat.device)
            tot = loc.clone()
            # the scalar all-reduce must not queue behind the backward kernels: it runs on the side stream too
        else:
            loc = tot = None
        main = torch.cuda.current_stream(flat.device) if flat.is_cuda else None
        if split is not None and self.overlap and flat.is_cuda:
            ev, n_head = split
            if self._side is None:
                self._side = torch.cuda.Stream(flat.device)
            side = self._side
            head, tail = flat[:n_head], flat[n_head:]
            with torch.cuda.stream(side):
                scale, scale_ready = None, None
                if self.exact:
                    if getattr(self, "_mask_event", None) is not None:
                        side.wait_event(self._mask_event)
                    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
                    scale = loc / tot
                    scale_ready = torch.cuda.Event()
                    scale_ready.record(side)
                side.wait_event(ev)                       # decoder-side gradients are final
                if self.exact:
                    tail.mul_(scale)
                else:
                    tail.div_(world)
                w_tail = dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            if self.exact:
                main.wait_event(scale_ready)              # `scale` was produced on the side stream
                head.mul_(scale)
            else:
                head.div_(world)
            w_head = dist.all_reduce(head, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            w_tail.wait(); w_head.wait()
            main.wait_stream(side)
            for t in (flat, tot, loc) if self.exact else (flat,):
                t.record_stream(side)
        else:
            if self.exact:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
                flat.mul_(loc / tot)
            else:
                flat.div_(world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += flat.numel() * 4


class DataParallelSAModel(torch.nn.Module):
    """Thin wrapper: same forward()/sample() surface; forward() works on this rank's shard and the
    backward hook all-reduces gradients.  Parameters are broadcast from rank 0 at construction, and
    BatchNorm running statistics can be re-synchronised with sync_buffers()."""

    def __init__(self, model, group=None, exact: bool = True, broadcast: bool = True, overlap: bool = True):
        super().__init__()
        self.module = model
        self.group = group
        self.hook = GradAllReduce(group, exact, overlap)
        object.__setattr__(model, "_grad_hook", self.hook)
        if broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
            with torch.no_grad():
                for p in model.parameters():
                    dist.broadcast(p, src=0, group=group)      # through the parameter: bumps Tensor._version
            model.params_changed()                             # ... and tell the engine explicitly as well
            self.sync_buffers()

    def sync_buffers(self):
        for b in self.module.buffers():
            if b.dtype.is_floating_point:
                dist.broadcast(b, src=0, group=self.group)

    def forward(self, feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask):
        self.hook.set_mask(seq_mask)
        return self.module(feats_rgb, feats_opfl, feat_mask, pos_feats, seq, seq_mask)

    def sample(self, feats_rgb, feats_opfl, feat_mask, pos_feats, opt={}):
        """Under model.train() this is the self-critical path (starttrain.py:131): the log-probs carry gradients and
        RewardCriterion normalises by the mask of the SAMPLED sequence, sum(cat[1, (seq > 0)[:, :-1]]) (SAModel.py:262-266);
        the `exact` rescale of the gradient all-reduce has to use that normaliser, not the last XE batch's."""
        seq, lps = self.module.sample(feats_rgb, feats_opfl, feat_mask, pos_feats, opt)
        if self.module.training and torch.is_grad_enabled() and lps.requires_grad:
            m = torch.cat([torch.ones_like(seq[:, :1]), (seq[:, :-1] > 0).to(seq.dtype)], 1)
            self.hook.set_mask(m)
        return seq, lps
