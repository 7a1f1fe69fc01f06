"""Run config-2 greedy decoding a few times (for traces / ncu)."""
import argparse, sys, torch
sys.path.insert(0, ".")
import bench
import controllable_xgating_b200 as X
from oracle import xgating_oracle as O
X.SAModel.VERBOSE = False
P = O.synth_params(bench.DIMS, 1024); P["logit.bias"][0] = -1e4
b = O.synth_inputs(bench.DIMS, 64, 28, 30, 0)
m = X.SAModel(bench.make_opt(0.5)); m.load_state_dict({k: v.clone() for k, v in P.items()}); m.cuda().eval()
d = {k: v.cuda() for k, v in b.items() if isinstance(v, torch.Tensor)}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    seq, _ = m.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], {"sample_max": 1, "beam_size": 1})
torch.cuda.synchronize()
print("ok", tuple(seq.shape))
