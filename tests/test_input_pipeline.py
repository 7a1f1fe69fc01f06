"""input_pipeline.collate_fn against the golden produced by the reference's own data_io.collate_fn
(tests/golden/make_collate_golden.py), and the pinned stager on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_collate_golden import make_items  # noqa: E402


def _collate():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("xg_input_pipeline", os.path.join(root, "controllable_xgating_b200", "input_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_collate_fn_matches_reference_golden():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate.npz"))
    out = _collate().collate_fn(make_items())
    data, caps, caps_mask, cap_classes, class_masks, feats1, feats2, feat_mask, pos_feat, lens, gts, image_id = out
    assert list(data) == list(g["data"]) and list(image_id) == list(g["image_id"]) and list(lens) == list(g["lens"])
    assert caps.dtype == torch.int64 and np.array_equal(caps.numpy(), g["caps"])
    assert np.array_equal(caps_mask.numpy(), g["caps_mask"])
    assert np.array_equal(cap_classes.numpy(), g["cap_classes"]) and np.array_equal(class_masks.numpy(), g["class_masks"])
    assert np.array_equal(feats1.numpy(), g["feats1"]) and np.array_equal(feats2.numpy(), g["feats2"])
    assert np.array_equal(feat_mask.numpy(), g["feat_mask"]) and np.array_equal(pos_feat.numpy(), g["pos_feat"])
    assert np.array_equal(np.stack([x.numpy() for x in gts]), g["gts"])
    assert (caps[:, 0] == 0).all() and sorted(lens, reverse=True) == list(lens)          # BOS column, longest first


@pytest.mark.gpu
def test_device_stager_round_trip():
    mod = _collate()
    st = mod.DeviceStager("cuda:0")
    for k in range(3):        # more batches than pinned slots
        host = {"a": torch.randn(64, 28, 32) + k, "b": torch.randint(0, 9, (64, 7))}
        dev = st.put(**host)
        st.wait()
        torch.cuda.current_stream().synchronize()
        assert all(torch.equal(dev[n].cpu(), host[n]) for n in host)


@pytest.mark.gpu
def test_device_stager_prefetch_order():
    """put(next) before wait(current): each batch carries its own event, the values stay those of its put()."""
    mod = _collate()
    st = mod.DeviceStager("cuda:0")
    hosts = [{"a": torch.full((256, 1024), float(k)).pin_memory()} for k in range(5)]
    nxt = st.put(**hosts[0])
    for k in range(5):
        cur = nxt
        if k + 1 < 5:
            nxt = st.put(**hosts[k + 1])
        st.wait(cur)
        assert float(cur["a"].sum()) == float(k) * 256 * 1024
