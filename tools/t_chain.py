"""Chain kernel vs one launch per layer on a C5-size batch: prints loss and gradient checksums (they must be
identical: same MMAs, same epilogues).  Usage: python tools/t_chain.py [R]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from loma_nerf_b200 import api, synthetic
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S, E, width, layers = 192, 10, 256, 9
c_in = 3 + 6 * E
dims = synthetic.mlp_dims(c_in, width, layers, 4)
ws, bs = synthetic.init_mlp(np.random.default_rng(216), dims)
g = torch.Generator(device="cuda").manual_seed(5)
X = torch.randn(R * S, c_in, device="cuda", generator=g)
dists = torch.rand(R, S, device="cuda", generator=g) * 0.05
target = torch.rand(R, 3, device="cuda", generator=g)
ctx = api.Context(0); ctx.set_stream(torch.cuda.current_stream())
cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()
for rep in range(3):
    out = ctx.nerf_step(dims, X, cv(ws), cv(bs), dists, target, R=R, S=S, grad=True, seed=1.0, outputs=("color", "loss"), path="tc")
    ctx.synchronize()
    print("loss %.6f color %.6f d_ws %.6e d_bs %.6e |d_ws| %.6e" % (out["loss"].item(), out["color"].double().sum().item(),
          out["d_ws"].double().sum().item(), out["d_bs"].double().sum().item(), out["d_ws"].double().abs().sum().item()), flush=True)
pr = ctx.profile_dominant(lambda: [ctx.nerf_step(dims, X, cv(ws), cv(bs), dists, target, R=R, S=S, grad=True, seed=1.0, outputs=("color", "loss"), path="tc") for _ in range(3)])
print(pr)
pr = ctx.profile_dominant(lambda: [ctx.nerf_step(dims, X, cv(ws), cv(bs), dists, target, R=R, S=S, grad=False, outputs=("color", "loss"), path="tc") for _ in range(3)])
print("forward only:", pr)
if hasattr(ctx.lib, "lnb_test_wide_clk"):
    import ctypes
    buf = (ctypes.c_ulonglong * 8)()
    ctx.lib.lnb_test_wide_clk(buf, 1)
    ctx.nerf_step(dims, X, cv(ws), cv(bs), dists, target, R=R, S=S, grad=False, outputs=("color", "loss"), path="tc"); ctx.synchronize()
    ctx.lib.lnb_test_wide_clk(buf, 0)
    n_cta = 148
    names = ["mma: wait tempty (epilogue frees the accumulator)", "mma: wait weights", "mma: wait A stage", "mma: total", "epi(thread 64): wait tfull", "epi(thread 64): work", "epi(thread 64): store-complete wait"]
    for i, nm in enumerate(names):
        print("%-55s %10.0f cycles per CTA" % (nm, buf[i] / n_cta))
