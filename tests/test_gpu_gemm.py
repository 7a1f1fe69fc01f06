"""The GEMM engines behind the batched products of the path (csrc/xg_gemm.cuh SIMT fp32, csrc/xg_gemm_tc.cuh tcgen05
3xTF32 / 3xFP16 operand pairs) against an fp64 product, over the layouts and odd extents a training step uses (K = 468 is not a multiple of
the 32-wide k-block, N = 10000 is not a multiple of the 128-row tile, nn / tn operands go through the transposing
split).  north_star tolerance for the path is 1e-3 relative; the engines themselves are held to 2e-6 of the largest
output, the class of an fp32 FFMA loop, because greedy token ids must survive them bit-exactly."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # (layout, M, N, K)   layout 0: A (M,K) . B (N,K)^T   1: A (M,K) . B (K,N)   2: A (K,M)^T . B (K,N)
    (0, 1792, 2048, 512), (0, 1984, 10000, 512), (0, 1984, 512, 468), (0, 1792, 512, 1024),
    (1, 1984, 512, 10000), (1, 1792, 512, 2048),
    (2, 512, 1024, 1792), (2, 10000, 512, 1984), (2, 2048, 512, 1728),
    (2, 500, 1000, 1000),      # 64 x 64 transposing split with partial tiles in both directions and a padded k-block
    (2, 501, 1002, 1000),      # leading dimensions that are not multiples of 4: the scalar transposing split
]


@pytest.mark.parametrize("layout,M,N,K", CASES)
@pytest.mark.parametrize("engine", [1, 2, 4])      # 1 SIMT fp32, 2 tcgen05 3xTF32, 4 tcgen05 3xFP16 pairs
def test_gemm_engine_vs_fp64(layout, M, N, K, engine):
    from controllable_xgating_b200.engine import debug_gemm
    g = torch.Generator(device="cuda").manual_seed(1000 * layout + M + N + K)
    A = torch.rand((M, K) if layout != 2 else (K, M), device="cuda", generator=g) - 0.3
    B = torch.rand((N, K) if layout == 0 else (K, N), device="cuda", generator=g) - 0.3
    a = A.double() if layout != 2 else A.double().t()
    b = B.double().t() if layout == 0 else B.double()
    ref = a @ b
    C = debug_gemm(layout, engine, A, B, M, N, K)
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < (2e-6 if engine in (2, 4) else 1e-5), err      # SIMT fp32 accumulates K = 10000 in one chain: 7e-6 measured
