import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api, synthetic
from oracle import oracle as O
ctx = api.Context(0); dev = torch.device('cuda', 0)
stream = torch.cuda.current_stream(dev); ctx.set_stream(stream)
R, S = 4096, 64
cases = []
for i in range(6):
    c = O.make_nerf_case(100 + i, R, S)
    cases.append({k: torch.as_tensor(np.ascontiguousarray(c[k], np.float32)).cuda() for k in ("X", "dists", "target", "rays_o", "rays_d", "t")})
c0 = O.make_nerf_case(1, 4, 64)
dims = [int(v) for v in c0["dims"]]
ws = torch.as_tensor(c0["ws"]).cuda(); bs = torch.as_tensor(c0["bs"]).cuda()
out = dict(d_ws=torch.zeros_like(ws), d_bs=torch.zeros_like(bs), loss=torch.zeros(1, device=dev))
def call(i, path):
    b = cases[i % 6]
    if path == "tc_rays":
        ctx.nerf_step_rays(dims, b["rays_o"], b["rays_d"], b["t"], 5, ws, bs, b["target"], grad=True, seed=1.0, outputs=("loss",), out=out, path="tc")
        return
    ctx.nerf_step(dims, b["X"], ws, bs, b["dists"], b["target"], R=R, S=S, grad=True, seed=1.0, outputs=("loss",), out=out, path=path)
for path in ("tc", "tc_rays"):
    for i in range(5): call(i, path)
    torch.cuda.synchronize()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(100): call(i, path)
        e1.record(stream)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(path, "eager loop: %.1f us/step gpu, %.1f us/step cpu-issue" % (e0.elapsed_time(e1) * 10, (t1 - t0) * 1e4), "loss", out["loss"].item())
    pr = ctx.profile_dominant(lambda: [call(i, path) for i in range(100)])
    print(path, "profiled:", pr)
# per-kernel wall: sync after each
for path in ("tc",):
    ts = []
    for i in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); call(i, path); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(path, "sync'd wall per call us:", [round(t * 1e6) for t in ts[5:]])
