import sys, ctypes
sys.path.insert(0, '.')
import torch
from loma_nerf_b200 import api
ctx = api.Context(0); ctx.set_stream(torch.cuda.current_stream())
lib = ctx.lib
lib.lnb_test_wide_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
for (M, N, K) in [(786432, 256, 256), (786432, 256, 64), (786432, 16, 256), (262144, 256, 256)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda")
    for _ in range(3): lib.lnb_test_wide_gemm(ctx.h, A.data_ptr(), B.data_ptr(), M, N, K, None, C.data_ptr())
    pr = ctx.profile_dominant(lambda: [lib.lnb_test_wide_gemm(ctx.h, A.data_ptr(), B.data_ptr(), M, N, K, None, C.data_ptr()) for _ in range(10)])
    us = pr["ms_per_launch"] * 1e3
    print(M, N, K, "%.1f us  %.1f TFLOP/s  (fp32 C store %.0f GB/s)" % (us, 2.0 * M * N * K / us / 1e6, M * N * 4 / us / 1e3))

lib.lnb_test_wide_gemm_bf16.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
for (M, N, K, masked) in [(786432, 256, 256, False), (786432, 256, 256, True), (786432, 256, 64, False)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); mask = torch.relu(torch.randn(M, N, device="cuda")).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    f = lambda: lib.lnb_test_wide_gemm_bf16(ctx.h, A.data_ptr(), B.data_ptr(), M, N, K, None if masked else bias.data_ptr(), mask.data_ptr() if masked else None, C.data_ptr())
    for _ in range(3): f()
    pr = ctx.profile_dominant(lambda: [f() for _ in range(10)])
    us = pr["ms_per_launch"] * 1e3
    print("bf16", M, N, K, "masked" if masked else "relu", "%.1f us  %.1f TFLOP/s" % (us, 2.0 * M * N * K / us / 1e6))
