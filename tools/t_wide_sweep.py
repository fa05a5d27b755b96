"""Randomised shape sweep of the wide tensor-core path against the float64 restatement (oracle/ is the checker)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from loma_nerf_b200 import api
from oracle import oracle as O

def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
ctx = api.Context(0); ctx.set_stream(torch.cuda.current_stream())
cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    E = int(rng.integers(1, 11)); width = int(rng.integers(63, 257)); layers = int(rng.integers(2, 12))
    R = int(rng.integers(1, 400)); S = int(rng.integers(1, 200)); rays = bool(rng.integers(0, 2)); grad = bool(rng.integers(0, 4))
    case = O.make_nerf_case(5000 + it, R, S, E=E, width=width, n_layers=layers)
    dims = [int(v) for v in case["dims"]]
    try:
        if rays:
            r = [torch.as_tensor(np.ascontiguousarray(case[k], np.float64)).cuda() for k in ("rays_o", "rays_d", "t")]
            out = ctx.nerf_step_rays(dims, r[0], r[1], r[2], E, cv(case["ws"]), cv(case["bs"]), cv(case["target"]), grad=grad, seed=1.0,
                                     outputs=("color", "loss"), path="tc")
        else:
            out = ctx.nerf_step(dims, cv(case["X"]), cv(case["ws"]), cv(case["bs"]), cv(case["dists"]), cv(case["target"]), R=R, S=S, grad=grad,
                                seed=1.0, outputs=("color", "loss"), path="tc")
        ctx.synchronize()
    except Exception as e:
        print("FAIL", dict(E=E, width=width, layers=layers, R=R, S=S, rays=rays, grad=grad), str(e)[:200]); continue
    f = O.nerf_f64(case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], R, S, g=1.0)
    errs = dict(loss=rel_err(out["loss"].cpu().numpy()[0], f["loss"]), color=rel_err(out["color"].cpu().numpy(), f["color"]))
    if grad:
        errs.update(d_ws=rel_err(out["d_ws"].cpu().numpy(), f["d_ws"]), d_bs=rel_err(out["d_bs"].cpu().numpy(), f["d_bs"]))
    w = max(errs.values()); worst = max(worst, w)
    flag = "BAD " if (not np.isfinite(w) or w > 6e-2) else "ok  "
    print(flag, dict(E=E, width=width, layers=layers, R=R, S=S, rays=rays, grad=grad), {k: float("%.2g" % v) for k, v in errs.items()}, flush=True)
print("worst", worst)
