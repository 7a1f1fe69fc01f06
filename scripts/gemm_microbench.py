"""Time the GEMM engines on one shape with CUDA events (engine 1 = SIMT fp32, 2 = tcgen05 3xTF32,
3 = SIMT fp32 with split-K for skinny shapes).  argv: MxNxK [reps] [layout]"""
import sys
import torch
sys.path.insert(0, ".")
from controllable_xgating_b200.engine import debug_gemm

shapes = [(1792, 2048, 512), (1984, 10000, 512), (64, 2048, 512), (64, 10000, 512), (64, 512, 2048), (64, 1024, 1536)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in sys.argv[1].split("x"))]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
layout = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for (M, N, K) in shapes:
    A = torch.rand(M, K, device="cuda") - 0.5
    B = (torch.rand(N, K, device="cuda") if layout == 0 else torch.rand(K, N, device="cuda")) - 0.5
    if layout == 2:
        A = torch.rand(K, M, device="cuda") - 0.5
    for eng in (1, 2, 3):
        for _ in range(3):
            debug_gemm(layout, eng, A, B, M, N, K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            debug_gemm(layout, eng, A, B, M, N, K)
        e1.record(); e1.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        print("layout %d shape %dx%dx%d engine %d: %.1f us/call  %.1f TFLOP/s (incl. operand split for engine 2)" % (layout, M, N, K, eng, us, 2.0 * M * N * K / us / 1e6), flush=True)
