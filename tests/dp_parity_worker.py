"""Worker of tests/test_gpu_dp.py (one process per GPU, NCCL; launched with torch.distributed.run).

Checks, on real hardware, that N ranks of the CUDA path + the gradient all-reduce hook produce
  (1) the single-GPU gradient of the loss over the CONCATENATED batch (eval-mode BatchNorm so that per-replica batch
      statistics do not enter; exact=True mask-sum rescale) to <= 1e-5 Frobenius, and
  (2) in training mode, the per-shard oracle gradients summed with the same weights (DP parity is
      "the oracle run per shard, grads summed") to <= 1e-3 Frobenius,
at config 4's per-GPU shard size (XG_DP_SHARD, default 256).  Rank 0 writes one JSON line to XG_DP_OUT.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import controllable_xgating_b200 as X
    import controllable_xgating_b200.SAModel as XS
    from controllable_xgating_b200.parallel import DataParallelSAModel, shard_batch
    from oracle import xgating_oracle as O
    from tests.test_gpu_parity import FULL, make_opt
    XS.VERBOSE = False
    shard = int(os.environ.get("XG_DP_SHARD", "256"))
    dims, K, T = FULL["dims"], FULL["K"], FULL["T"]
    P = O.synth_params(dims, 1024)
    full = O.synth_inputs(dims, shard * world, K, T, seed=77, full_length=False)
    full = {k: v for k, v in full.items() if isinstance(v, torch.Tensor)}
    mine = shard_batch(full, world, rank)
    crit = X.LanguageModelCriterion()

    def model(train):
        m = X.SAModel(make_opt(dims, T, drop=0.0))
        m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
        m.cuda()
        m.train() if train else m.eval()
        m._engine.set_strict(True)
        return m

    def step(fwd, m, b):
        d = {k: v.cuda() for k, v in b.items()}
        for p in m.parameters():
            p.grad = None
        logp, _ = fwd(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], d["seq"], d["seq_mask"])
        loss = crit(logp, d["seq"], d["seq_mask"])
        loss.backward()
        return {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters()}

    out = {"world": world, "shard": shard}
    for train in (False, True):
        m = model(train)
        dp = DataParallelSAModel(m, exact=True)
        g_dp = step(dp, m, mine)
        assert dp.hook.calls == 1 and dp.hook.bytes == 4 * sum(p.numel() for p in m.parameters())
        # every rank must hold the same reduced gradient
        chk = torch.tensor([float(sum(g.double().sum() for g in g_dp.values()))], device="cuda", dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(hi - lo) == 0.0, "ranks disagree on the reduced gradient"
        if rank == 0:
            worst, worst_n = 0.0, ""
            if not train:
                ms = model(False)
                g_ref = step(ms, ms, full)
                tol = 1e-5
            else:
                tot = float(full["seq_mask"].sum())
                g_ref = None
                for r in range(world):
                    sh = shard_batch(full, world, r)
                    _, g = O.train_step_grads(P, sh, train=True)
                    w = float(sh["seq_mask"].sum()) / tot
                    g_ref = {n: g[n] * w for n in g} if g_ref is None else {n: g_ref[n] + g[n] * w for n in g}
                tol = 1e-3
            for n, ref in g_ref.items():
                den = float(ref.double().norm())
                if den < 1e-6:        # Linear biases in front of BatchNorm: mathematically zero gradient
                    continue
                e = float((g_dp[n].double() - ref.double()).norm()) / den
                if e > worst:
                    worst, worst_n = e, n
            out["train_mode" if train else "eval_mode"] = {"worst_rel_fro": worst, "tensor": worst_n, "tol": tol,
                                                           "against": "per-shard oracle gradients, mask-weighted sum" if train
                                                           else "single-GPU CUDA gradient of the concatenated batch"}
            assert worst < tol, (train, worst, worst_n)
        dist.barrier()
    if rank == 0:
        with open(os.environ["XG_DP_OUT"], "w") as f:
            f.write(json.dumps(out) + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
