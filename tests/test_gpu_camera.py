"""Camera mode: rays and sample depths generated on the device from a pose, against the hosts'
own construction -- get_rays (train_nerf.py:23-62, restated in loma_nerf_b200/render.py), depth = linspace(near, far, S)
(train_nerf.py:289), pts = o + d t, dists (train_nerf.py:306-311)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from loma_nerf_b200 import api
    assert torch.cuda.is_available()
    c = api.Context(0)
    yield torch, api, c
    c.close()


def _scene(width=64, theta=35.0):
    from loma_nerf_b200 import render
    focal = 0.5 / np.tan(0.5 * 0.6911)
    K = np.array([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], np.float64)
    return K, render.pose_spherical(theta, -30.0, 4.0)


def test_camera_rays_and_linspace_depths_are_the_hosts_values(env):
    torch, api, ctx = env
    from loma_nerf_b200 import render
    W, S = 64, 30
    K, pose = _scene(W)
    o_ref, d_ref = render.get_rays(W, W, K, pose)
    t_ref = np.linspace(2.0, 6.0, S)
    cam = api.make_camera(pose, K, W, near=2.0, far=6.0)
    o, d, t = (v.cpu().numpy() for v in ctx.camera_rays(cam, W * W, S))
    assert np.array_equal(o, o_ref)
    assert np.array_equal(t, np.broadcast_to(t_ref, (W * W, S)))            # linspace, endpoint included, bit for bit
    # dirs @ R^T: numpy's matmul may fuse multiply-adds differently: at most one float64 ulp, invisible after the float32 cast
    assert np.abs(d - d_ref).max() <= 4 * np.finfo(np.float64).eps * np.abs(d_ref).max()
    # arbitrary pixels and a window
    pix = np.array([0, 63, 64, 4095, 2017], np.int32)
    cam2 = api.make_camera(pose, K, W, pixels=torch.as_tensor(pix).cuda())
    o2, d2, _ = (v.cpu().numpy() for v in ctx.camera_rays(cam2, len(pix), S))
    assert np.abs(d2 - d_ref[pix]).max() <= 4 * np.finfo(np.float64).eps * np.abs(d_ref).max()
    cam3 = api.make_camera(pose, K, W, first_pixel=1000)
    _, d3, _ = (v.cpu().numpy() for v in ctx.camera_rays(cam3, 77, S))
    assert np.array_equal(d3, d[1000:1077])


def test_stratified_depths_follow_the_documented_generator(env):
    torch, api, ctx = env
    W, S, seed = 32, 64, 12345
    K, pose = _scene(W)
    cam = api.make_camera(pose, K, W, stratified=True, seed=seed, first_pixel=5)
    t = ctx.camera_rays(cam, 40, S)[2].cpu().numpy()
    u = np.array([[ctx.lib.lnb_uniform(seed, 5 + r, s) for s in range(S)] for r in range(40)])
    assert np.array_equal(t, 2.0 + (np.arange(S)[None, :] + u) * (6.0 - 2.0) / S)
    assert (np.diff(t, axis=1) > 0).all() and t.min() >= 2.0 and t.max() < 6.0


@pytest.mark.parametrize("path,tol", [("f32", 1e-5), ("tc", 3e-2)])
def test_camera_mode_step_equals_the_step_on_host_built_rays(env, path, tol):
    """The whole train step from a pose alone against the same step fed with the hosts' features: the reference's own
    shape (4 rays x 30 samples would be one chunk; here a 16 x 16 window of a 64 x 64 frame) and a stratified batch."""
    torch, api, ctx = env
    from loma_nerf_b200 import render
    W, S, E = 64, 30, 5
    K, pose = _scene(W)
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(8), dims, 1.0)
    cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    R = 256
    rng = np.random.default_rng(9)
    target = rng.uniform(0, 1, (R, 3)).astype(np.float32)
    pix = rng.choice(W * W, R, replace=False).astype(np.int32)
    o_all, d_all = render.get_rays(W, W, K, pose)
    for stratified in (False, True):
        cam = api.make_camera(pose, K, W, pixels=torch.as_tensor(pix).cuda(), stratified=stratified, seed=77)
        t = ctx.camera_rays(cam, R, S)[2].cpu().numpy() if stratified else np.broadcast_to(np.linspace(2.0, 6.0, S), (R, S))
        pts, dists = O.sample_points(o_all[pix], d_all[pix], t)
        X = O.positional_encoding(pts, E).reshape(R * S, -1)
        ref = O.nerf_f64(X, ws, bs, dims, target, dists, R, S, g=1.0)
        out = ctx.nerf_step_camera(dims, cam, R, S, E, cv(ws), cv(bs), cv(target), grad=True, seed=1.0, outputs=("color", "loss"), path=path)
        ctx.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
        errs = dict(loss=rel_err(got["loss"][0], ref["loss"]), color=rel_err(got["color"], ref["color"]),
                    d_ws=rel_err(got["d_ws"], ref["d_ws"]), d_bs=rel_err(got["d_bs"], ref["d_bs"]))
        assert max(errs.values()) <= tol, (stratified, errs)


def test_camera_mode_features_are_bit_equal_to_the_hosts_on_the_exact_path(env):
    """inter[0] of the exact path is a function of the features only: equal outputs for equal weights <=> the float32
    features generated from the pose equal positional_encoding(o + d t) as the host computes them."""
    torch, api, ctx = env
    from loma_nerf_b200 import render
    W, S, E = 32, 30, 5
    K, pose = _scene(W, theta=110.0)
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(18), dims, 1.0)
    cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    R = W * W
    o, d = render.get_rays(W, W, K, pose)
    t = np.broadcast_to(np.linspace(2.0, 6.0, S), (R, S))
    pts, dists = O.sample_points(o, d, t)
    X = O.positional_encoding(pts, E).reshape(R * S, -1)
    a = ctx.nerf_step(dims, cv(X), cv(ws), cv(bs), cv(dists), None, R=R, S=S, grad=False, outputs=("color", "alpha"), path="f32")
    b = ctx.nerf_step_camera(dims, api.make_camera(pose, K, W), R, S, E, cv(ws), cv(bs), None, grad=False, outputs=("color", "alpha"), path="f32")
    ctx.synchronize()
    mism = (a["color"] != b["color"]).float().mean().item()
    assert mism <= 1e-3, mism          # a last-bit difference of d (see above) can flip a float32 rounding once in a while
    assert rel_err(b["color"].cpu().numpy(), a["color"].cpu().numpy()) <= 1e-6


def test_device_frame_render_returns_uint8_and_matches_the_host_path(env):
    torch, api, ctx = env
    from loma_nerf_b200 import render
    W, S, E = 48, 64, 5
    K, pose = _scene(W, theta=200.0)
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(28), dims, 1.0)
    ref = render.render_frame(ctx, dims, ws, bs, W, W, K, pose, S, E, path="f32")
    cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    for path, tol in (("f32", 1), ("tc", 4)):
        img = render.render_frame_device(ctx, dims, cv(ws), cv(bs), W, W, K, pose, S, E, path=path, rays_per_call=1000)
        ctx.synchronize()
        got = img.cpu().numpy()
        assert got.dtype == np.uint8 and got.shape == (W, W, 3)
        want = np.rint(255 * ref.clip(0, 1)).astype(np.int32)
        assert np.abs(got.astype(np.int32) - want).max() <= tol, path


def test_trainer_steps_from_a_camera_batch_on_host_and_device(env):
    torch, api, ctx = env
    W, S, E, R = 64, 64, 5, 512
    K, pose = _scene(W)
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(38), dims, 1.0)
    rng = np.random.default_rng(39)
    pix = rng.choice(W * W, R, replace=False).astype(np.int32)
    target = rng.uniform(0, 1, (R, 3)).astype(np.float32)
    tr_d = api.Trainer(ctx, dims, ws, bs)
    tr_h = api.Trainer(ctx, dims, ws, bs)
    for it in range(3):
        cam_d = api.make_camera(pose, K, W, pixels=torch.as_tensor(pix).cuda(), stratified=True, seed=100 + it)
        tr_d.step(camera=cam_d, S=S, pe_bands=E, target=torch.as_tensor(target).cuda(), path="tc")
        cam_h = api.make_camera(pose, K, W, pixels=pix, stratified=True, seed=100 + it)       # host pixel list, host targets
        loss_h = tr_h.step_host(camera=cam_h, S=S, pe_bands=E, target=target, path="tc")
    wd, bd, loss_d = tr_d.read()
    wh, bh, _ = tr_h.read()
    # same arithmetic on both routes; the weight gradients of a launch are summed in TMEM in MMA completion order, so the
    # two trajectories agree to fp32 rounding, not bit for bit
    tol = lambda a, b: float(np.abs(a - b).max()) <= 2e-5 * float(np.abs(b).max())  # noqa: E731
    assert tol(wd, wh) and tol(bd, bh) and abs(loss_d - loss_h) <= 1e-5 * abs(loss_h)
    tr_d.close(); tr_h.close()


@pytest.mark.parametrize("W", [37, 100])
def test_frame_of_a_width_that_is_no_power_of_two_or_tile_multiple(env, W):
    """Pixel row / column come from a 64-bit reciprocal multiply in the fused kernels (no integer division per tile): a
    width that divides nothing, rendered in calls that start in the middle of a pixel row."""
    torch, api, ctx = env
    from loma_nerf_b200 import render
    S, E = 24, 5
    K, pose = _scene(W, theta=75.0)
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(48), dims, 1.0)
    ref = render.render_frame(ctx, dims, ws, bs, W, W, K, pose, S, E, path="f32")
    want = np.rint(255 * ref.clip(0, 1)).astype(np.int32)
    cv = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    for path, tol in (("f32", 1), ("tc", 4)):
        img = render.render_frame_device(ctx, dims, cv(ws), cv(bs), W, W, K, pose, S, E, path=path, rays_per_call=W * 3 + 11)
        ctx.synchronize()
        assert np.abs(img.cpu().numpy().astype(np.int32) - want).max() <= tol, (W, path)
