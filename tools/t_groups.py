"""Sweep the groups per CTA of fused_mg_kernel (LNB_TC_GROUPS) and compare with fused_v1_kernel (LNB_TC_V1): CUDA-event
time per launch of the fused kernel, C2 batch (4096 x 64) and a batch 8 x larger, features / rays input."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api, synthetic
ctx = api.Context(0); dev = torch.device('cuda', 0)
E, S = 5, 64
dims = synthetic.mlp_dims(33, 30, 3, 4)
ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)
ws, bs = torch.as_tensor(ws_np).cuda(), torch.as_tensor(bs_np).cuda()
for R in [int(a) for a in (sys.argv[1:] or ["4096", "32768"])]:
    rng = np.random.default_rng(1)
    batches = []
    for i in range(3 if R > 8192 else 7):
        o, d = synthetic.random_rays(rng, R); t = synthetic.stratified_t(rng, R, S)
        od, dd, td = (torch.as_tensor(v).cuda() for v in (o, d, t))
        X, dists = ctx.sample_encode(od, dd, td, E)
        batches.append(dict(o=od, d=dd, t=td, X=X, dists=dists, target=torch.as_tensor(rng.uniform(0, 1, (R, 3)).astype(np.float32)).cuda()))
    out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
    def call(i, mode):
        b = batches[i % len(batches)]
        if mode == "rays":
            ctx.nerf_step_rays(dims, b["o"], b["d"], b["t"], E, ws, bs, b["target"], grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
        else:
            ctx.nerf_step(dims, b["X"], ws, bs, b["dists"], b["target"], R=R, S=S, grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
    for mode in ("features", "rays"):
        res = []
        for cfg in ["v1"] + [str(n) for n in range(1, 8)]:
            os.environ.pop("LNB_TC_V1", None); os.environ.pop("LNB_TC_GROUPS", None)
            if cfg == "v1": os.environ["LNB_TC_V1"] = "1"
            else: os.environ["LNB_TC_GROUPS"] = cfg
            for i in range(5): call(i, mode)
            torch.cuda.synchronize()
            pr = ctx.profile_dominant(lambda: [call(i, mode) for i in range(40)])
            res.append("%s: %.1f" % (cfg, pr["ms_per_launch"] * 1e3))
        print("R=%d %s  us/launch  " % (R, mode) + "  ".join(res), "loss", float(out["loss"].item()))
