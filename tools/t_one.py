import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api
from oracle import oracle as O
mode = sys.argv[1] if len(sys.argv) > 1 else "rays"
ctx = api.Context(0); dev = torch.device('cuda', 0)
ctx.set_stream(torch.cuda.current_stream(dev))
R, S = 4096, 64
c = O.make_nerf_case(100, R, S)
b = {k: torch.as_tensor(np.ascontiguousarray(c[k], np.float32)).cuda() for k in ("X", "dists", "target", "rays_o", "rays_d", "t")}
dims = [int(v) for v in c["dims"]]
ws = torch.as_tensor(c["ws"]).cuda(); bs = torch.as_tensor(c["bs"]).cuda()
out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
for i in range(6):
    if mode == "rays":
        ctx.nerf_step_rays(dims, b["rays_o"], b["rays_d"], b["t"], 5, ws, bs, b["target"], grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
    else:
        ctx.nerf_step(dims, b["X"], ws, bs, b["dists"], b["target"], R=R, S=S, grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
torch.cuda.synchronize()
print("ok", out["loss"].item())
