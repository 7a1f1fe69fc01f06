"""Host issue time of one greedy sample() call against its GPU time (is the greedy path host bound?).
Run through gpurun: python scripts/host_issue_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
import controllable_xgating_b200 as X
import controllable_xgating_b200.SAModel as XS
from oracle import xgating_oracle as O

XS.VERBOSE = False
dev = torch.device("cuda", 0)
P = O.synth_params(B.DIMS, 1024)
P["logit.bias"][0] = -1e4
batch = O.synth_inputs(B.DIMS, B.BATCH, B.K_FRAMES, B.T_SEQ, seed=0)
model = X.SAModel(B.make_opt(0.5))
model.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
model.cuda().eval()
d = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
gopt = {"sample_max": 1, "beam_size": 1}
for _ in range(5):
    model.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], gopt)
torch.cuda.synchronize()
for rep in range(3):
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        model.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], gopt)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("host issue %.3f ms/call   gpu %.3f ms/call" % ((t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n))
# the asynchronous call (no step count read back): pure host issue cost
for rep in range(3):
    n = 50
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        model.sample_async(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], gopt)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("async: host issue %.3f ms/call   gpu %.3f ms/call" % ((t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n))
# short bursts (the launch queue never fills): what the host needs for one call
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        model.sample_async(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], gopt)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("burst of 4: host issue %.3f ms/call, drained after %.3f ms" % ((t1 - t0) / 4 * 1e3, (t2 - t0) * 1e3))
eng = model._engine
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    enc = model._encode(d["rgb"], d["opfl"], d["feat_mask"])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("encode: host %.3f ms" % ((t1 - t0) * 1e3))
