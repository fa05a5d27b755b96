"""Wide-MLP tensor-core path (wide_tc.cu): the TMA-fed, 128B-swizzled tcgen05 GEMM and the weight-
gradient kernel against fp32 matrix products of the same bf16-rounded operands (torch is only the
checker here), then the whole layerwise step against the reference golden vectors of the 8x256 net."""
import ctypes

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    from loma_nerf_b200 import api
    c = api.Context(0)
    c.set_stream(torch_cuda.cuda.current_stream())
    yield c
    c.close()


@pytest.mark.parametrize("M,N,K", [(128, 16, 64), (1000, 64, 64), (4096 + 77, 256, 256), (300, 32, 128), (20000, 256, 64)])
def test_wide_gemm_matches_fp32_product_of_bf16_operands(ctx, torch_cuda, M, N, K):
    torch = torch_cuda
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda")
    lib = ctx.lib
    lib.lnb_test_wide_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    ctx._check(lib.lnb_test_wide_gemm(ctx.h, A.data_ptr(), B.data_ptr(), M, N, K, bias.data_ptr(), C.data_ptr()))
    ctx.synchronize()
    ref = A.float() @ B.float().t() + bias
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) <= 2e-5


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 256, 256), (4096 + 77, 128, 64), (50000, 256, 256)])
@pytest.mark.parametrize("masked", [False, True])
def test_wide_gemm_bf16_epilogues(ctx, torch_cuda, M, N, K, masked):
    torch = torch_cuda
    g = torch.Generator(device="cuda").manual_seed(M + N + K + int(masked))
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16).contiguous()
    bias = torch.randn(N, device="cuda", generator=g)
    mask = torch.relu(torch.randn(M, N, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    C = torch.full((M, N), float("nan"), device="cuda").to(torch.bfloat16)
    lib = ctx.lib
    lib.lnb_test_wide_gemm_bf16.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
    ctx._check(lib.lnb_test_wide_gemm_bf16(ctx.h, A.data_ptr(), B.data_ptr(), M, N, K, None if masked else bias.data_ptr(),
                                           mask.data_ptr() if masked else None, C.data_ptr()))
    ctx.synchronize()
    acc = A.float() @ B.float().t()
    ref = torch.where(mask > 0, acc, torch.zeros_like(acc)) if masked else torch.relu(acc + bias)
    got = C.float().cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_err(got, ref.to(torch.bfloat16).float().cpu().numpy()) <= 1e-2    # one bf16 rounding of the output
