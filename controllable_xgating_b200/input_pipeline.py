"""Input side of the hot path: `collate_fn` with the semantics of the reference's
data_io.collate_fn (/root/reference/caption_src/data_io.py:330-374) and a pinned-memory stager that moves the five
tensors the model consumes to the device (18.5 MB per batch of 64) with asynchronous copies.

Out of scope (unchanged): reading the pre-extracted feature files (h5py), vocabulary building, the Dataset class.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch


def collate_fn(batch):
    """Same outputs, order and dtypes as data_io.collate_fn: items are tuples
    (data, cap, cap_class, class_mask, feat1 (K,R), feat2 (K,F), feat_mask (1,K), pos_feat (H,), gts).

    Returns (data, caps (m,max_len+1) int64 with column 0 = 0, caps_mask (m,max_len+1) with len+1 ones,
    cap_classes, class_masks, feats1 (m,K,R), feats2 (m,K,F), feat_mask (m,K), pos_feat (m,H), lens, gts, image_id),
    the batch sorted by caption length, longest first (data_io.py:331)."""
    batch = sorted(batch, key=lambda x: len(x[1]), reverse=True)      # stable, like list.sort
    data, cap, cap_class, class_mask, feat1, feat2, feat_mask, pos_feat, gts = zip(*batch)
    m = len(cap)
    max_len = len(cap[0])
    feats1 = torch.stack(feat1, dim=0)
    feats2 = torch.stack(feat2, dim=0)
    fmask = torch.cat(feat_mask, dim=0)
    pos = torch.stack(pos_feat, dim=0)
    caps = torch.zeros(m, max_len + 1, dtype=torch.int64)
    caps_mask = torch.zeros(m, max_len + 1)
    lens = []
    for i, c in enumerate(cap):
        caps[i, 1:len(c) + 1] = torch.as_tensor(list(c), dtype=torch.int64)
        caps_mask[i, :len(c) + 1] = 1
        lens.append(len(c))
    cap_classes = torch.zeros(m, max_len + 1, dtype=torch.int64)
    class_masks = torch.zeros(m, max_len + 1)
    for i in range(m):
        cc, cm = list(cap_class[i]), list(class_mask[i])
        cap_classes[i, :len(cc)] = torch.as_tensor(cc, dtype=torch.int64)
        class_masks[i, :len(cm)] = torch.as_tensor(cm, dtype=torch.float32)
        class_masks[i, len(cm)] = 1
    gts = [torch.from_numpy(np.asarray(x)).long() for x in gts]
    image_id = [i.split("_")[0] for i in data]
    return data, caps, caps_mask, cap_classes, class_masks, feats1, feats2, fmask, pos, lens, gts, image_id


class DeviceStager:
    """Double-buffered pinned staging of the tensors SAModel.forward / sample consume.

        stager = DeviceStager(device)
        dev = stager.put(feats1=..., feats2=..., feat_mask=..., pos_feat=..., caps=..., caps_mask=...)   # async H2D
        stager.wait()                                  # the compute stream waits for the copies (no host sync)

    Host tensors are copied into reusable pinned buffers (two sets, so the next batch can be staged while the
    previous one is still in flight) and moved on a dedicated copy stream."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceStager needs a CUDA device (there is no CPU fallback of the hot path)")
        self.depth = depth
        self.stream = torch.cuda.Stream(self.device)
        self._pinned = [dict() for _ in range(depth)]
        self._events = [None] * depth
        self._slot = 0
        self._last = None

    def _pin(self, slot: int, name: str, t: torch.Tensor) -> torch.Tensor:
        buf = self._pinned[slot].get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[slot][name] = buf
        buf.copy_(t)
        return buf

    def put(self, **host: torch.Tensor) -> "StagedBatch":
        slot = self._slot
        self._slot = (slot + 1) % self.depth
        if self._events[slot] is not None:
            self._events[slot].synchronize()           # the pinned buffers of this slot are free again
        out = StagedBatch()
        with torch.cuda.stream(self.stream):
            for name, t in host.items():
                src = t if t.is_pinned() else self._pin(slot, name, t)
                out[name] = src.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        out.ready = ev
        self._events[slot] = ev
        self._last = out
        return out

    def wait(self, batch: Optional["StagedBatch"] = None, stream: Optional[torch.cuda.Stream] = None):
        """Make `stream` (default: the current stream) wait for the copies of `batch` (default: the last put()).
        Passing the batch lets the NEXT batch be staged before the current one is consumed: put(next) ->
        wait(current) -> compute(current), the copy of `next` overlapping the compute."""
        batch = batch if batch is not None else self._last
        if batch is None:
            return
        stream = stream or torch.cuda.current_stream(self.device)
        stream.wait_event(batch.ready)
        for t in batch.values():
            t.record_stream(stream)                    # allocated on the copy stream, consumed on this one: the caching
                                                       # allocator must not hand the block to a later put() too early


class StagedBatch(dict):
    """name -> device tensor of one put(); `.ready` is the event the copies complete at."""
    ready: Optional[torch.cuda.Event] = None
