"""Accuracy of the GEMM engines against fp64 over the layouts / shapes of a training step.
XG_TC_RAW=0/1 selects the tcgen05 kernel (pre-split operands / raw operands split in flight)."""
import sys
import torch
sys.path.insert(0, ".")
from controllable_xgating_b200.engine import debug_gemm

torch.manual_seed(0)
cases = [(0, 1792, 2048, 512), (0, 1984, 10000, 512), (0, 1984, 512, 468), (1, 1984, 512, 10000), (1, 1792, 512, 2048),
         (2, 512, 1024, 1792), (2, 512, 1536, 1792), (2, 10000, 512, 1984), (2, 2048, 512, 1728), (2, 1536, 1024, 1984)]
for (layout, M, N, K) in cases:
    A = (torch.rand(M, K, device="cuda") - 0.3) if layout != 2 else (torch.rand(K, M, device="cuda") - 0.3)
    B = (torch.rand(N, K, device="cuda") - 0.3) if layout == 0 else (torch.rand(K, N, device="cuda") - 0.3)
    a = A.double() if layout != 2 else A.double().t()
    b = B.double().t() if layout == 0 else B.double()
    ref = a @ b
    for eng in (1, 2, 4):
        C = debug_gemm(layout, eng, A, B, M, N, K)
        d = (C.double() - ref).abs()
        print("layout %d %5dx%5dx%5d engine %d: max rel err %.3e  fro %.3e  bad(>1e-4) %d" % (
            layout, M, N, K, eng, (d.max() / ref.abs().max()).item(), (d.norm() / ref.norm()).item(),
            int((d > 1e-4 * ref.abs().max()).sum())), flush=True)
