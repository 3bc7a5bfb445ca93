"""-m gpu: tcgen05 building blocks (bf16 split-plane GEMM tile, SS and TS operand modes) against torch fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ts", [0, 1])
@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("K,N", [(32, 16), (96, 256), (256, 256), (256, 96), (64, 80)])
def test_tc_gemm_tile(ts, planes, K, N):
    import sdfstudio_b200 as sb

    lib = sb._lib.load()
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K**0.5).cuda()
    Np = (N + 15) // 16 * 16
    D = torch.full((128, Np), float("nan"), device="cuda")
    scratch = torch.zeros((K // 32) * planes * Np * 64, dtype=torch.uint8, device="cuda")
    sb._lib.check(lib.sdfb200_debug_tc_gemm(A.data_ptr(), W.data_ptr(), K, N, ts, planes, D.data_ptr(), scratch.data_ptr(), 0), "debug_tc_gemm")
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).float()
    scale = float((A.abs().double() @ W.abs().double().t()).max())
    err = float((D[:, :N] - ref).abs().max()) / scale
    assert Np == N or float(D[:, N:].abs().max()) == 0.0
    tol = 2e-5 if planes == 2 else 6e-3
    assert err < tol, f"relative-to-|A||W| error {err:.3e} (planes={planes}, ts={ts})"
    if planes == 2:
        # the split must actually buy precision: error well below a single bf16 pass
        assert err < 1e-4
