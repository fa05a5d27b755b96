"""fused_mg_kernel launch time against the number of tile rounds per group (7 groups x 148 SMs = 1036 tiles per round)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api, synthetic
ctx = api.Context(0); dev = torch.device('cuda', 0)
E, S = 5, 64
dims = synthetic.mlp_dims(33, 30, 3, 4)
ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)
ws, bs = torch.as_tensor(ws_np).cuda(), torch.as_tensor(bs_np).cuda()
for tiles in (148, 296, 592, 1036, 2072, 3108, 4144, 8288):
    R = tiles * 2
    rng = np.random.default_rng(1)
    o, d = synthetic.random_rays(rng, R); t = synthetic.stratified_t(rng, R, S)
    od, dd, td = (torch.as_tensor(v).cuda() for v in (o, d, t))
    X, dists = ctx.sample_encode(od, dd, td, E)
    target = torch.as_tensor(rng.uniform(0, 1, (R, 3)).astype(np.float32)).cuda()
    out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
    res = []
    for mode in ("features", "rays"):
        def call():
            if mode == "rays":
                ctx.nerf_step_rays(dims, od, dd, td, E, ws, bs, target, grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
            else:
                ctx.nerf_step(dims, X, ws, bs, dists, target, R=R, S=S, grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
        for i in range(5): call()
        torch.cuda.synchronize()
        pr = ctx.profile_dominant(lambda: [call() for i in range(40)])
        res.append("%s %.1f us" % (mode, pr["ms_per_launch"] * 1e3))
    print("tiles %5d (%.2f per SM): " % (tiles, tiles / 148.0) + "  ".join(res))
