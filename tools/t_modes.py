import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api
from oracle import oracle as O
ctx = api.Context(0); dev = torch.device('cuda', 0)
ctx.set_stream(torch.cuda.current_stream(dev))
R, S = 4096, 64
c = O.make_nerf_case(100, R, S)
b = {k: torch.as_tensor(np.ascontiguousarray(c[k], np.float32)).cuda() for k in ("X", "dists", "target", "rays_o", "rays_d", "t")}
dims = [int(v) for v in c["dims"]]
ws = torch.as_tensor(c["ws"]).cuda(); bs = torch.as_tensor(c["bs"]).cuda()
out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
def call(mode, grad):
    if mode == "rays":
        ctx.nerf_step_rays(dims, b["rays_o"], b["rays_d"], b["t"], 5, ws, bs, b["target"], grad=grad, seed=1.0, outputs=("loss",), out=out if grad else dict(loss=out["loss"]), path="tc")
    else:
        ctx.nerf_step(dims, b["X"], ws, bs, b["dists"], b["target"], R=R, S=S, grad=grad, seed=1.0, outputs=("loss",), out=out if grad else dict(loss=out["loss"]), path="tc")
for mode in ("feat", "rays"):
    for grad in (True, False):
        for i in range(3): call(mode, grad)
        pr = ctx.profile_dominant(lambda: [call(mode, grad) for i in range(50)])
        print(mode, "grad" if grad else "fwd ", "%.1f us" % (pr["ms_per_launch"] * 1e3))
