"""Time of the gradient all-reduce of a training step (26.3 M fp32 = 105 MB, one flat buffer) under the NCCL settings in
the environment: torchrun --nproc-per-node N scripts/microbench/allreduce_bench.py [floats]"""
import os, sys, torch, torch.distributed as dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26312159
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
x = torch.ones(n, device="cuda")
for _ in range(5):
    dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    w = dist.get_world_size()
    print("world %d  %d floats  %.3f ms per all-reduce  algbw %.0f GB/s  busbw %.0f GB/s   [%s]" % (
        w, n, t.item(), n * 4 / t.item() / 1e6, n * 4 / t.item() / 1e6 * 2 * (w - 1) / w,
        " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("NCCL_"))), flush=True)
dist.destroy_process_group()
