"""Golden for input_pipeline.collate_fn: runs the REFERENCE's data_io.collate_fn (/root/reference/caption_src/
data_io.py:330-374, imported with an h5py stub) on a seeded synthetic list of items and stores its outputs.
Run in the build container only (the reference does not exist on the GPU box)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def make_items(seed=0, n=7, K=5, R=12, F=8, H=6):
    rng = np.random.default_rng(seed)
    items = []
    for i in range(n):
        ln = int(rng.integers(2, 9))
        cap = [int(x) for x in rng.integers(2, 50, ln)]
        cls = [int(x) for x in rng.integers(0, 14, ln)]
        cm = [int(x) for x in rng.integers(0, 2, ln)]
        f1 = torch.from_numpy(rng.random((K, R), dtype=np.float32))
        f2 = torch.from_numpy(rng.random((K, F), dtype=np.float32))
        if i % 3 == 0:
            f1[-2:] = 0; f2[-2:] = 0
        fm = (torch.sum(f1.view(K, -1), dim=1, keepdim=True) != 0).float().transpose(1, 0)
        pos = torch.from_numpy(rng.standard_normal(H).astype(np.float32))
        gts = rng.integers(0, 50, (3, 6)).astype(np.int64)
        items.append(("video%d_%d" % (i, i % 2), cap, cls, cm, f1, f2, fm, pos, gts))
    return items


if __name__ == "__main__":
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, "/root/reference/caption_src")
    sys.argv = ["x"]
    import data_io as ref   # noqa: E402
    out = ref.collate_fn(make_items())
    data, caps, caps_mask, cap_classes, class_masks, feats1, feats2, feat_mask, pos_feat, lens, gts, image_id = out
    np.savez(os.path.join(HERE, "collate.npz"), data=np.array(data), caps=caps.numpy(), caps_mask=caps_mask.numpy(),
             cap_classes=cap_classes.numpy(), class_masks=class_masks.numpy(), feats1=feats1.numpy(), feats2=feats2.numpy(),
             feat_mask=feat_mask.numpy(), pos_feat=pos_feat.numpy(), lens=np.array(lens), image_id=np.array(image_id),
             gts=np.stack([g.numpy() for g in gts]))
    print("wrote collate.npz")
