"""Per-kernel CUDA-event table (xg_profile_enable/_report) of one workload: greedy | train | beam.

    python scripts/profile_path.py train [steps]
"""
import ctypes, json, sys, torch
sys.path.insert(0, ".")
import bench
import controllable_xgating_b200 as X
import controllable_xgating_b200.SAModel as XS
from controllable_xgating_b200 import _lib as XL
from oracle import xgating_oracle as O
XS.VERBOSE = False
what = sys.argv[1] if len(sys.argv) > 1 else "train"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
P = O.synth_params(bench.DIMS, 1024); P["logit.bias"][0] = -1e4
b = O.synth_inputs(bench.DIMS, B, 28, 30, 0, full_length=(what != "train"))
m = X.SAModel(bench.make_opt(0.5)); m.load_state_dict({k: v.clone() for k, v in P.items()}); m.cuda()
d = {k: v.cuda() for k, v in b.items() if isinstance(v, torch.Tensor)}
crit = X.LanguageModelCriterion()


def step():
    if what == "train":
        m.train()
        for p in m.parameters():
            p.grad = None
        logp, _ = m(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        crit(logp, d["seq"], d["seq_mask"]).backward()
    else:
        m.eval()
        m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 5 if what == "beam" else 1})


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    step()
e1.record(); torch.cuda.synchronize()
print("%s: %.3f ms/step unprofiled (B=%d)" % (what, e0.elapsed_time(e1) / n, B))
eng, lib = m._engine, XL.load()
XL.check(lib.xg_profile_enable(eng.handle, 1), "xg_profile_enable", eng.handle)
for _ in range(n):
    step()
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 20)
XL.check(lib.xg_profile_report(eng.handle, buf, len(buf)), "xg_profile_report", eng.handle)
XL.check(lib.xg_profile_enable(eng.handle, 0), "xg_profile_enable", eng.handle)
prof = json.loads(buf.value.decode())
prof.sort(key=lambda p: -p["ms"])
tot = sum(p["ms"] for p in prof)
print("sum of kernel time %.3f ms/step over %d launches/step" % (tot / n, sum(p["launches"] for p in prof) / n))
for p in prof[:40]:
    print("%-44s %6.1f launches/step %9.1f us/launch %7.3f ms/step %5.1f%%" % (
        p["name"], p["launches"] / n, p["ms"] / p["launches"] * 1e3, p["ms"] / n, 100 * p["ms"] / tot))
