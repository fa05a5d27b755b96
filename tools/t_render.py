"""A few camera-mode forward-only frames (800 x 800 x 64) for ncu / timing: python tools/t_render.py [n_frames]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from loma_nerf_b200 import api, synthetic, render
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = api.Context(0); dev = torch.device('cuda', 0)
dims = synthetic.mlp_dims(33, 30, 3, 4)
ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)
ws, bs = torch.as_tensor(ws_np).cuda(), torch.as_tensor(bs_np).cuda()
K = np.array([[synthetic.FOCAL, 0, 0.5], [0, synthetic.FOCAL, 0.5], [0, 0, 1.0]])
color = torch.empty((640000, 3), dtype=torch.float32, device=dev)
u8 = torch.empty((640000, 3), dtype=torch.uint8, device=dev)
for i in range(n):
    render.render_frame_device(ctx, dims, ws, bs, 800, 800, K, render.pose_spherical(30.0 * i, -30.0, 4.0), 64, 5, out_u8=u8, color=color)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    render.render_frame_device(ctx, dims, ws, bs, 800, 800, K, render.pose_spherical(30.0 * i, -30.0, 4.0), 64, 5, out_u8=u8, color=color)
e1.record(); torch.cuda.synchronize()
print("ok %.3f ms per frame, checksum %d" % (e0.elapsed_time(e1) / n, int(u8.sum().item())))
