"""One fused exact (fp32) train step repeated a few times (for ncu): python tools/t_one_f32.py [R]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api, synthetic
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S, E = 64, 5
ctx = api.Context(0); dev = torch.device('cuda', 0)
rng = np.random.default_rng(1)
dims = synthetic.mlp_dims(33, 30, 3, 4)
ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)
ws, bs = torch.as_tensor(ws_np).cuda(), torch.as_tensor(bs_np).cuda()
o, d = synthetic.random_rays(rng, R); t = synthetic.stratified_t(rng, R, S)
od, dd, td = (torch.as_tensor(v).cuda() for v in (o, d, t))
X, dists = ctx.sample_encode(od, dd, td, E)
target = torch.as_tensor(rng.uniform(0, 1, (R, 3)).astype(np.float32)).cuda()
out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
for i in range(6):
    ctx.nerf_step(dims, X, ws, bs, dists, target, R=R, S=S, grad=True, seed=1.0, outputs=("loss",), out=out, path="f32")
torch.cuda.synchronize()
print("ok", out["loss"].item())
