"""Parity tests proper (B200): the CUDA path, called through the C ABI, against
  * the golden vectors recorded from the REAL reference (tests/golden/*.npz),
  * the C oracle (oracle/nerf_oracle.c) on freshly seeded inputs,
  * the float64 restatement at BASELINE.json's full sizes, plus size-independent properties
    (linearity in the seed, additivity over ray shards, += accumulation).
Tolerance: norm-wise relative error <= 1e-5 for the fp32 paths (BASELINE.json north_star).
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden, rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
NERF_GOLDEN = golden_files("nerf_")
FIT_GOLDEN = golden_files("fit_")
F32_PATHS = ["f32_layerwise", "f32"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    from loma_nerf_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def dev(torch, a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda().contiguous()


def host(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


ALL_NERF_OUT = ("inter", "rgba", "alpha", "cumprod", "weights", "color", "loss", "d_ws", "d_bs",
                "d_X", "d_target", "d_dists", "d_color", "d_inter")


def run_nerf(ctx, torch, gd, path, seed, device=True, rows=None, outputs=ALL_NERF_OUT):
    cv = (lambda a: dev(torch, a, torch.float32)) if device else (lambda a: np.ascontiguousarray(a, np.float32))
    R, S = int(gd["R"]), int(gd["S"])
    out = ctx.nerf_step([int(v) for v in gd["dims"]], cv(gd["X"]), cv(gd["ws"]), cv(gd["bs"]),
                        cv(gd["dists"]), cv(gd["target"]), R=R, S=S, grad=True, seed=seed,
                        outputs=outputs, rows=rows, path=path)
    if device:
        ctx.synchronize()
    return {k: host(v) for k, v in out.items()}


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", F32_PATHS)
@pytest.mark.parametrize("gpath", NERF_GOLDEN, ids=os.path.basename)
def test_flat_nerf_matches_reference_golden(ctx, torch_cuda, gpath, path):
    gd = load_golden(gpath)
    N = int(gd["R"]) * int(gd["S"])
    L = len(gd["dims"]) - 1
    o = run_nerf(ctx, torch_cuda, gd, path, "loss")
    assert rel_err(o["loss"][0], gd["loss"]) <= TOL
    for k in ["rgba", "alpha", "cumprod", "weights", "color"]:
        assert rel_err(o[k], gd[k]) <= TOL, k
    for l in range(L):
        w = int(gd["dims"][l + 1])
        assert rel_err(o["inter"][l, :N, :w], gd["inter"][l, :, :w]) <= TOL, ("inter", l)
        assert rel_err(o["d_inter"][l, :N, :w], gd["d_inter"][l, :, :w]) <= TOL, ("d_inter", l)
    for k, kg in [("d_ws", "d_ws"), ("d_bs", "d_bs"), ("d_X", "d_X"), ("d_target", "d_target"),
                  ("d_dists", "d_dists"), ("d_color", "d_acc")]:
        assert rel_err(o[k], gd[kg]) <= TOL, k
    o1 = run_nerf(ctx, torch_cuda, gd, path, 1.0, outputs=("loss", "d_ws", "d_bs", "d_X", "d_dists"))
    for k in ["d_ws", "d_bs", "d_X", "d_dists"]:
        assert rel_err(o1[k], gd[k + "_g1"]) <= TOL, k


@pytest.mark.parametrize("path", F32_PATHS)
@pytest.mark.parametrize("gpath", FIT_GOLDEN, ids=os.path.basename)
def test_flat_fit_matches_reference_golden(ctx, torch_cuda, gpath, path):
    torch = torch_cuda
    gd = load_golden(gpath)
    N, L = int(gd["N"]), len(gd["dims"]) - 1
    cv = lambda a: dev(torch, a, torch.float32)  # noqa: E731
    out = ctx.fit_step([int(v) for v in gd["dims"]], cv(gd["X"]), cv(gd["ws"]), cv(gd["bs"]),
                       cv(gd["target"]), grad=True, seed="loss",
                       outputs=("inter", "loss", "d_ws", "d_bs", "d_X", "d_target", "d_inter"), path=path)
    ctx.synchronize()
    o = {k: host(v) for k, v in out.items()}
    assert rel_err(o["loss"][0], gd["loss"]) <= TOL
    for l in range(L):
        w = int(gd["dims"][l + 1])
        assert rel_err(o["inter"][l, :N, :w], gd["inter"][l, :, :w]) <= TOL
        assert rel_err(o["d_inter"][l, :N, :w], gd["d_inter"][l, :, :w]) <= TOL
    for k in ["d_ws", "d_bs", "d_X", "d_target"]:
        assert rel_err(o[k], gd[k]) <= TOL, k


@pytest.mark.parametrize("seed,R,S,E,width,layers", [
    (11, 5, 30, 5, 30, 3), (12, 3, 64, 5, 30, 3), (13, 1, 1, 2, 8, 2), (14, 33, 7, 3, 16, 4),
    (15, 2, 100, 5, 30, 3), (16, 9, 33, 4, 64, 3), (17, 2, 192, 10, 128, 5)])
def test_host_path_matches_c_oracle_on_fresh_seeds(ctx, torch_cuda, c_oracle, seed, R, S, E, width, layers):
    """lnb_nerf_step_host (host pointers, pageable numpy) with the reference's rows=256 convention."""
    case = O.make_nerf_case(seed, R, S, E=E, width=width, n_layers=layers)
    rows = max(256, R * S)
    args = (case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], R, S)
    f = c_oracle.nerf_forward(*args, rows=rows)
    b = c_oracle.nerf_backward(*args, float(f["loss"]), rows=rows)
    o = run_nerf(ctx, torch_cuda, case, "f32_layerwise", float(f["loss"]), device=False, rows=rows)
    assert rel_err(o["loss"][0], f["loss"]) <= TOL
    assert rel_err(o["inter"], f["inter"]) <= TOL       # includes the rows >= N the reference fills
    for k in ["rgba", "alpha", "cumprod", "weights", "color"]:
        assert rel_err(o[k], f[k]) <= TOL, k
    for k, kb in [("d_ws", "d_ws"), ("d_bs", "d_bs"), ("d_X", "d_X"), ("d_target", "d_target"),
                  ("d_dists", "d_dists"), ("d_color", "d_acc"), ("d_inter", "d_inter")]:
        assert rel_err(o[k], b[kb]) <= TOL, k


# --------------------------------------------------------------------------------------------
# the five reference symbols through the ragged ABI, driven exactly like the reference hosts
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def compat(torch_cuda):
    from loma_nerf_b200 import _lib
    return O.CompatCaller(ctypes.CDLL(_lib.LIB_PATH), big_stack=False)


@pytest.mark.parametrize("gpath", [p for p in NERF_GOLDEN if "c5" not in p], ids=os.path.basename)
def test_compat_nerf_symbols_match_reference_golden(compat, c_oracle, gpath):
    gd = load_golden(gpath)
    R, S = int(gd["R"]), int(gd["S"])
    N, L = R * S, len(gd["dims"]) - 1
    r = compat.nerf(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S, g="loss")
    assert rel_err(r["loss"], gd["loss"]) <= TOL
    for k in ["rgba", "alpha", "cumprod", "weights", "color"]:
        assert rel_err(r[k], gd[k]) <= TOL, k
    mo = gd["ws"].shape[2]
    assert rel_err(r["inter"][:, :N, :mo], gd["inter"]) <= TOL
    # rows >= N: the reference's bias/activation loops run over all 256 declared rows
    f = c_oracle.nerf_forward(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S, rows=256)
    assert rel_err(r["inter"][:, :256, :mo], f["inter"]) <= TOL
    for k, kg in [("d_ws", "d_ws"), ("d_bs", "d_bs"), ("d_X", "d_X"), ("d_target", "d_target"),
                  ("d_dists", "d_dists"), ("d_acc", "d_acc")]:
        assert rel_err(r[k], gd[kg]) <= TOL, k
    assert rel_err(r["d_inter"][:, :N, :mo], gd["d_inter"]) <= TOL
    # what the reference leaves untouched
    for k in ["d_rgba", "d_alpha", "d_cumprod", "d_weights"]:
        assert not r[k].any(), k
    assert all(not v.any() for v in r["primal_after_grad"].values())


def test_compat_paper_size_mlp(compat):
    gd = load_golden([p for p in NERF_GOLDEN if "c5" in p][0])
    R, S = int(gd["R"]), int(gd["S"])
    big = O.CompatCaller(compat.lib, big_stack=False, scratch_rows=192, scratch_cols=256)
    r = big.nerf(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S, g="loss")
    assert rel_err(r["loss"], gd["loss"]) <= TOL
    for k in ["color", "d_ws", "d_bs", "d_X", "d_dists"]:
        assert rel_err(r[k], gd[k]) <= TOL, k


@pytest.mark.parametrize("gpath", FIT_GOLDEN, ids=os.path.basename)
def test_compat_mlp_fit_symbols_match_reference_golden(compat, gpath):
    gd = load_golden(gpath)
    N = int(gd["N"])
    mo = gd["ws"].shape[2]
    r = compat.mlp_fit(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], g="loss")
    assert rel_err(r["loss"], gd["loss"]) <= TOL
    assert rel_err(r["inter"][:, :N, :mo], gd["inter"]) <= TOL
    for k in ["d_ws", "d_bs", "d_X", "d_target"]:
        assert rel_err(r[k], gd[k]) <= TOL, k
    assert rel_err(r["d_inter"][:, :N, :mo], gd["d_inter"]) <= TOL
    assert not r["primal_after_grad"]["inter"].any() and not r["d_out"].any()


def test_compat_mult_a_b_known_answer(compat):
    """The reference's only known-answer test on this path (fit_img.py:363-374)."""
    c = compat.mult_a_b(np.array([[1, 2], [3, 4], [5, 6]], np.float32), np.array([[100], [200]], np.float32))
    assert np.array_equal(c, np.array([[500], [1100], [1700]], np.float32))
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(37, 19)).astype(np.float32), rng.normal(size=(19, 70)).astype(np.float32)
    assert rel_err(compat.mult_a_b(a, b), a.astype(np.float64) @ b) <= TOL


def test_compat_accumulates_into_gradient_buffers_and_nonzero_scratch(compat, c_oracle):
    """d_ buffers are += (reverse_diff.py:144-155); intermediate_outputs and accumulated_color are
    accumulated onto their entry contents (nerf.py:86,113,284-286)."""
    case = O.make_nerf_case(41, 3, 16)
    R, S = 3, 16
    lib = compat.lib
    X, ws, bs = case["X"], case["ws"], case["bs"]
    dims = [int(v) for v in case["dims"]]
    L, N = len(dims) - 1, R * S
    rows, cols = 256, max(dims[1:])
    wsh, bsh, ish = compat._shapes(dims, rows)
    inter = np.full((L, rows, cols), 0.25, np.float32)
    acc = np.full((R, 3), 0.5, np.float32)
    rgba = np.zeros((R, S, 4), np.float32); al = np.zeros((R, S), np.float32)
    cu = np.zeros((R, S), np.float32); wg = np.zeros((R, S), np.float32)
    keep = []
    def t3(a):
        t, k = O.rows3(a); keep.append(k); return t
    loss = lib.nerf_evaluate_and_march(
        O.rows2(X), N, X.shape[1], t3(ws), O.rows2(bs), O.rows2(case["target"]), R, 3, L,
        O.rows2(wsh, O.c_int_p), O.rows2(bsh, O.c_int_p), O.rows2(ish, O.c_int_p), t3(inter), t3(rgba),
        S, O.rows2(case["dists"]), O.rows2(al), O.rows2(cu), O.rows2(wg), O.rows2(acc))
    # oracle with the same entry contents
    o_inter = np.full((L, rows, cols), 0.25, np.float32)
    o_acc = np.full((R, 3), 0.5, np.float32)
    o_rgba = np.zeros((R, S, 4), np.float32); o_al = np.zeros((R, S), np.float32)
    o_cu = np.zeros((R, S), np.float32); o_wg = np.zeros((R, S), np.float32)
    fp = lambda a: a.ctypes.data_as(O.c_float_p)  # noqa: E731
    d32 = np.ascontiguousarray(case["dims"], np.int32)
    o_loss = c_oracle.lib.oracle_nerf_forward(
        fp(X), N, X.shape[1], fp(ws), fp(bs), L, d32.ctypes.data_as(O.c_int_p), ws.shape[1], ws.shape[2],
        fp(case["target"]), R, S, fp(case["dists"]), rows, cols, fp(o_inter), fp(o_rgba), fp(o_al),
        fp(o_cu), fp(o_wg), fp(o_acc))
    assert rel_err(loss, o_loss) <= TOL
    assert rel_err(inter, o_inter) <= TOL and rel_err(acc, o_acc) <= TOL and rel_err(wg, o_wg) <= TOL


# --------------------------------------------------------------------------------------------
# encoders, optimisers
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,E", [(2, 5), (3, 5), (3, 10), (3, 0)])
def test_pos_encoding_matches_reference_formula(ctx, torch_cuda, F, E):
    torch = torch_cuda
    x = np.random.default_rng(5).uniform(-6, 6, (1000, F))
    enc = host(ctx.pos_encoding(dev(torch, x, torch.float64), E))
    ctx.synchronize()
    ref = O.positional_encoding(x, E)
    assert enc.shape == ref.shape
    assert np.array_equal(enc[:, :F], x.astype(np.float32))          # pos_encoding.py:34,68
    assert np.abs(enc - ref).max() <= 2.0 ** -23                      # float64 sin/cos, 1 f32 ulp


def test_sample_encode_matches_host_construction(ctx, torch_cuda):
    torch = torch_cuda
    case = O.make_nerf_case(77, 13, 64)
    X, dists = ctx.sample_encode(dev(torch, case["rays_o"], torch.float64), dev(torch, case["rays_d"], torch.float64),
                                 dev(torch, case["t"], torch.float64), 5)
    ctx.synchronize()
    assert np.abs(host(X) - case["X"]).max() <= 2.0 ** -22
    assert np.array_equal(host(dists), case["dists"])


def test_adam_and_sgd_follow_the_reference_update_rules(ctx, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(9)
    p = rng.normal(size=3060).astype(np.float32)
    m = np.zeros_like(p); v = np.zeros_like(p)
    pd, md, vd = (dev(torch, a.copy(), torch.float32) for a in (p, m, v))
    lr, b1, b2, eps = 5e-4, 0.9, 0.999, 1e-8
    for t in range(1, 6):
        g = rng.normal(size=p.size).astype(np.float32) * 10 ** rng.uniform(-3, 1)
        # train_nerf.py:147-159, numpy float32 arrays with python-float scalars
        lr_t = lr * (np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * (g ** 2)
        p = p - lr_t * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + eps)
        ctx.adam_step(pd, dev(torch, g, torch.float32), md, vd, t, lr, b1, b2, eps)
    ctx.synchronize()
    assert rel_err(host(pd), p) <= 1e-6 and rel_err(host(md), m) <= 1e-6 and rel_err(host(vd), v) <= 1e-6
    g = rng.normal(size=p.size).astype(np.float32)
    q = dev(torch, p.astype(np.float32), torch.float32)
    ctx.sgd_step(q, dev(torch, g, torch.float32), 1e-4)
    ctx.synchronize()
    assert rel_err(host(q), p.astype(np.float32) - np.float32(1e-4) * g) <= 1e-7


# --------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 2 and 5)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", F32_PATHS)
def test_c2_full_batch_against_f64_restatement(ctx, torch_cuda, path):
    """BASELINE config 2: 33->30->30->4, 4096 rays x 64 samples."""
    case = O.make_nerf_case(215, 4096, 64)
    o = run_nerf(ctx, torch_cuda, case, path, 1.0, outputs=("color", "loss", "d_ws", "d_bs", "d_dists"))
    f = O.nerf_f64(case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], 4096, 64, g=1.0)
    assert rel_err(o["loss"][0], f["loss"]) <= TOL
    assert rel_err(o["color"], f["color"]) <= TOL
    for k in ["d_ws", "d_bs", "d_dists"]:
        assert rel_err(o[k], f[k]) <= TOL, k


@pytest.mark.parametrize("path", F32_PATHS)
def test_linearity_additivity_and_accumulation(ctx, torch_cuda, path):
    torch = torch_cuda
    case = O.make_nerf_case(300, 512, 64)
    outs = ("loss", "d_ws", "d_bs")
    a = run_nerf(ctx, torch, case, path, 1.0, outputs=outs)
    b = run_nerf(ctx, torch, case, path, 3.5, outputs=outs)
    assert rel_err(b["d_ws"], 3.5 * a["d_ws"]) <= 1e-6           # gradients are linear in _dreturn
    # additivity over ray shards (what multi-GPU sharding relies on)
    tot_w, tot_l = np.zeros_like(a["d_ws"], np.float64), 0.0
    for r0 in range(0, 512, 128):
        sub = dict(case, X=case["X"][r0 * 64:(r0 + 128) * 64], target=case["target"][r0:r0 + 128],
                   dists=case["dists"][r0:r0 + 128], R=128)
        s = run_nerf(ctx, torch, sub, path, 1.0, outputs=outs)
        tot_w += s["d_ws"]; tot_l += float(s["loss"][0])
    assert rel_err(tot_w, a["d_ws"]) <= TOL and abs(tot_l - float(a["loss"][0])) <= TOL * tot_l
    # += semantics: a second call into the same buffers doubles them
    cv = lambda x: dev(torch, x, torch.float32)  # noqa: E731
    bufs = dict(d_ws=torch.zeros(case["ws"].shape, device="cuda"), d_bs=torch.zeros(case["bs"].shape, device="cuda"))
    for _ in range(2):
        ctx.nerf_step([int(v) for v in case["dims"]], cv(case["X"]), cv(case["ws"]), cv(case["bs"]),
                      cv(case["dists"]), cv(case["target"]), grad=True, seed=1.0, outputs=outs, out=bufs, path=path)
    ctx.synchronize()
    assert rel_err(host(bufs["d_ws"]), 2 * a["d_ws"]) <= 1e-6


@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("gpath", [p for p in NERF_GOLDEN if "c5" not in p], ids=os.path.basename)
def test_rays_mode_exact_path_matches_reference_golden(ctx, torch_cuda, gpath, device):
    """Rays in (float64, as get_rays / linspace make them): the device builds pts, the positional
    encoding and dists in float64 (encode.cu) and then runs the exact fp32 kernels."""
    torch = torch_cuda
    gd = load_golden(gpath)
    R, S, E = int(gd["R"]), int(gd["S"]), int(gd["E"])
    if device:
        cv = lambda a, dt=torch.float32: dev(torch, a, dt)  # noqa: E731
        rays = [cv(gd[k], torch.float64) for k in ("rays_o", "rays_d", "t")]
    else:
        cv = lambda a, dt=None: np.ascontiguousarray(a, np.float32)  # noqa: E731
        rays = [np.ascontiguousarray(gd[k], np.float64) for k in ("rays_o", "rays_d", "t")]
    out = ctx.nerf_step_rays([int(v) for v in gd["dims"]], rays[0], rays[1], rays[2], E, cv(gd["ws"]), cv(gd["bs"]),
                             cv(gd["target"]), grad=True, seed="loss",
                             outputs=("color", "loss", "d_ws", "d_bs", "d_dists", "d_target"), path="f32")
    if device:
        ctx.synchronize()
    o = {k: host(v) for k, v in out.items()}
    assert rel_err(o["loss"][0], gd["loss"]) <= TOL
    for k in ["color", "d_ws", "d_bs", "d_dists", "d_target"]:
        assert rel_err(o[k], gd[k]) <= TOL, k


def test_render_only_and_edge_shapes(ctx, torch_cuda):
    torch = torch_cuda
    case = O.make_nerf_case(55, 64, 64)
    cv = lambda x: dev(torch, x, torch.float32)  # noqa: E731
    dims = [int(v) for v in case["dims"]]
    out = ctx.nerf_step(dims, cv(case["X"]), cv(case["ws"]), cv(case["bs"]), cv(case["dists"]),
                        target=None, grad=False, outputs=("color",))
    ctx.synchronize()
    f = O.nerf_f64(case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], 64, 64)
    assert rel_err(host(out["color"]), f["color"]) <= TOL
    # zero rays: nothing to do, loss 0, gradients untouched
    e = ctx.nerf_step(dims, cv(case["X"][:0]), cv(case["ws"]), cv(case["bs"]), cv(case["dists"][:0]),
                      cv(case["target"][:0]), grad=True, outputs=("loss", "d_ws"))
    ctx.synchronize()
    assert float(host(e["loss"])[0]) == 0.0 and not host(e["d_ws"]).any()
    # saturated densities: transmittance underflows, the reverse sweep must stay finite
    sat = O.make_nerf_case(56, 8, 64, sigma_bias_shift=50.0)
    o = run_nerf(ctx, torch, sat, "f32", 1.0, outputs=("loss", "d_ws", "d_bs", "d_X", "d_dists"))
    assert all(np.isfinite(v).all() for v in o.values())
    f = O.nerf_f64(sat["X"], sat["ws"], sat["bs"], sat["dims"], sat["target"], sat["dists"], 8, 64, g=1.0)
    assert rel_err(o["d_ws"], f["d_ws"]) <= TOL
    # bad arguments are refused with an error, not a crash
    from loma_nerf_b200 import api
    with pytest.raises(api.LnbError):
        ctx.nerf_step([33, 30, 3], cv(case["X"]), cv(case["ws"]), cv(case["bs"]), cv(case["dists"]), cv(case["target"]))


def test_render_frame_matches_reference_per_ray_colours(ctx, torch_cuda):
    """Forward-only frame render (train_nerf.py:581-661): every pixel's accumulated colour, shared
    linspace depths, against the C oracle fed with the host-built features."""
    from loma_nerf_b200 import render
    rng = np.random.default_rng(77)
    E, S, H = 5, 30, 24
    dims = O.mlp_dims(3 + 6 * E, 30, 3, 4)
    ws, bs = O.init_mlp(rng, dims, 1.0)
    K = np.array([[1.2, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1.0]])
    c2w = np.eye(4); c2w[:3, 3] = [0.3, -0.2, 4.0]
    img = render.render_frame(ctx, dims, ws, bs, H, H, K, c2w, S, E, path="f32")
    o, d = render.get_rays(H, H, K, c2w)
    pts, dists = O.sample_points(o, d, np.linspace(2.0, 6.0, S))
    X = O.positional_encoding(pts, E).reshape(H * H * S, -1)
    f = O.nerf_f64(X, ws, bs, dims, np.zeros((H * H, 3), np.float32), dists, H * H, S)
    assert img.shape == (H, H, 3)
    assert rel_err(img.reshape(-1, 3), f["color"]) <= TOL
    img_tc = render.render_frame(ctx, dims, ws, bs, H, H, K, c2w, S, E, path="tc")
    assert rel_err(img_tc.reshape(-1, 3), f["color"]) <= 3e-2
    assert render.compute_psnr(img_tc, img) > 35.0


def _ragged_like_the_reference(arr):
    """float** / float*** built the way the reference hosts build them (mlp_utils.py:33-118): every
    row is its OWN ctypes allocation (not a view of one contiguous block), values copied in."""
    import ctypes
    from ctypes import POINTER, c_float, cast
    a = np.asarray(arr, np.float32)
    fp = POINTER(c_float)
    keep = []
    if a.ndim == 2:
        tab = (fp * a.shape[0])()
        for i in range(a.shape[0]):
            row = (c_float * a.shape[1])(*a[i].tolist())
            keep.append(row)
            tab[i] = cast(row, fp)
        return cast(tab, POINTER(fp)), keep + [tab]
    outer = (POINTER(fp) * a.shape[0])()
    for i in range(a.shape[0]):
        inner, k = _ragged_like_the_reference(a[i])
        keep.append(k)
        outer[i] = inner
    return cast(outer, POINTER(POINTER(fp))), keep + [outer]


def _read_ragged(ptr, shape):
    out = np.empty(shape, np.float32)
    for idx in np.ndindex(*shape[:-1]):
        p = ptr
        for i in idx:
            p = p[i]
        out[idx] = [p[j] for j in range(shape[-1])]
    return out


def test_compat_symbols_with_truly_ragged_rows_as_the_reference_marshaller_builds_them(compat, c_oracle):
    """Same call sequence as train_nerf.py:325-483: per-row ctypes allocations in, results read back
    through the ctypes objects element by element."""
    import ctypes
    case = O.make_nerf_case(215, 4, 30, stratified=False)           # the reference's own chunk shape
    R, S, N = 4, 30, 120
    dims = [int(v) for v in case["dims"]]
    L, rows, cols = len(dims) - 1, 256, 256
    wsh, bsh, ish = compat._shapes(dims, rows)
    lib = compat.lib
    K = []
    def rg(a):
        p, k = _ragged_like_the_reference(a); K.append(k); return p
    X, ws, bs, tg, dd = rg(case["X"]), rg(case["ws"]), rg(case["bs"]), rg(case["target"]), rg(case["dists"])
    inter, rgba = rg(np.zeros((L, rows, cols))), rg(np.zeros((R, S, 4)))
    al, cu, wg, acc = rg(np.zeros((R, S))), rg(np.zeros((R, S))), rg(np.zeros((R, S))), rg(np.zeros((R, 3)))
    ip = lambda a: O.rows2(a, O.c_int_p)  # noqa: E731
    loss = lib.nerf_evaluate_and_march(X, N, dims[0], ws, bs, tg, R, 3, L, ip(wsh), ip(bsh), ip(ish), inter, rgba, S, dd,
                                       al, cu, wg, acc)
    f = c_oracle.nerf_forward(case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], R, S, rows=256)
    assert rel_err(loss, f["loss"]) <= TOL
    assert rel_err(_read_ragged(acc, (R, 3)), f["color"]) <= TOL
    assert rel_err(_read_ragged(wg, (R, S)), f["weights"]) <= TOL
    assert rel_err(_read_ragged(inter, (L, N, 30)), f["inter"][:, :N, :30]) <= TOL
    # the grad call, exactly the 41 positional arguments (train_nerf.py:395-478), _dreturn = loss
    z = lambda shape: rg(np.zeros(shape))  # noqa: E731
    d_X, d_ws, d_bs, d_tg, d_dd = z((N, dims[0])), z(case["ws"].shape), z(case["bs"].shape), z((R, 3)), z((R, S))
    d_inter, d_rgba, d_al, d_cu, d_wg, d_acc = z((L, rows, cols)), z((R, S, 4)), z((R, S)), z((R, S)), z((R, S)), z((R, 3))
    p_inter, p_rgba, p_al, p_cu, p_wg, p_acc = z((L, rows, cols)), z((R, S, 4)), z((R, S)), z((R, S)), z((R, S)), z((R, 3))
    di = [ctypes.c_int(0) for _ in range(7)]
    zi = lambda a: O.rows2(np.zeros_like(a), O.c_int_p)  # noqa: E731
    lib.grad_nerf_evaluate_and_march(
        X, d_X, N, ctypes.byref(di[0]), dims[0], ctypes.byref(di[1]), ws, d_ws, bs, d_bs, tg, d_tg, R, ctypes.byref(di[2]),
        3, ctypes.byref(di[3]), L, ctypes.byref(di[4]), ip(wsh), zi(wsh), ip(bsh), zi(bsh), ip(ish), zi(ish),
        p_inter, d_inter, p_rgba, d_rgba, S, ctypes.byref(di[5]), dd, d_dd, p_al, d_al, p_cu, d_cu, p_wg, d_wg,
        p_acc, d_acc, float(loss))
    b = c_oracle.nerf_backward(case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], R, S,
                               float(f["loss"]), rows=256)
    mi, mo = case["ws"].shape[1], case["ws"].shape[2]
    assert rel_err(_read_ragged(d_ws, (L, mi, mo)), b["d_ws"]) <= TOL
    assert rel_err(_read_ragged(d_bs, (L, mo)), b["d_bs"]) <= TOL
    assert rel_err(_read_ragged(d_X, (N, dims[0])), b["d_X"]) <= TOL
    assert not _read_ragged(p_acc, (R, 3)).any() and not _read_ragged(d_wg, (R, S)).any()
    assert all(v.value == 0 for v in di)
