"""Pins the oracle (oracle/nerf_oracle.c + the float64 numpy restatement) to the reference:
  * the reference's own known-answer test mult_a_b (fit_img.py:363-374),
  * the PE identity-prefix property (pos_encoding.py:34,68),
  * golden vectors recorded from the REAL reference (tests/golden/make_golden.py),
  * live against oracle/_ref/*.so when present (they are built where /root/reference exists).
CPU only.
"""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden, rel_err
from oracle import oracle as O

NERF_GOLDEN = golden_files("nerf_")
FIT_GOLDEN = golden_files("fit_")
# the serial C restatement needs ~minutes for the 9x256 case; the f64 restatement covers it
SMALL_NERF = [p for p in NERF_GOLDEN if "c5" not in p]


def test_golden_present():
    assert len(NERF_GOLDEN) >= 8 and len(FIT_GOLDEN) >= 2


def test_mult_a_b_known_answer(c_oracle):
    c = c_oracle.mult_a_b(np.array([[1, 2], [3, 4], [5, 6]], np.float32),
                          np.array([[100], [200]], np.float32))
    assert np.array_equal(c, np.array([[500], [1100], [1700]], np.float32))


@pytest.mark.parametrize("F,E", [(2, 5), (3, 5), (3, 10), (3, 0)])
def test_pos_encoding_identity_prefix_and_layout(c_oracle, F, E):
    rng = np.random.default_rng(1)
    x = rng.uniform(-6, 6, (7, 5, F))
    enc = O.positional_encoding(x, E)
    assert enc.shape == (7, 5, F * (1 + 2 * E)) and enc.dtype == np.float32
    assert np.isclose(x, enc[..., :F]).all()                       # pos_encoding.py:34,68
    for i in range(E):
        assert np.array_equal(enc[..., (2 * i + 1) * F:(2 * i + 2) * F],
                              np.sin(2.0 ** i * x).astype(np.float32))
        assert np.array_equal(enc[..., (2 * i + 2) * F:(2 * i + 3) * F],
                              np.cos(2.0 ** i * x).astype(np.float32))
    enc_c = c_oracle.pos_encoding(x, E)
    # libm vs numpy sin/cos may differ in the last float64 bit -> at most 1 float32 ulp
    assert np.abs(enc_c - enc).max() <= 2.0 ** -23


@pytest.mark.parametrize("path", SMALL_NERF, ids=os.path.basename)
def test_c_oracle_nerf_matches_reference_golden(c_oracle, path):
    gd = load_golden(path)
    R, S = int(gd["R"]), int(gd["S"])
    N = R * S
    args = (gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S)
    f = c_oracle.nerf_forward(*args, rows=256)
    # forward: bit-exact (same fp32 operation order as the generated C)
    assert np.float32(f["loss"]) == gd["loss"]
    assert np.array_equal(f["inter"][:, :N], gd["inter"])
    for k in ["rgba", "alpha", "cumprod", "weights", "color"]:
        assert np.array_equal(f[k], gd[k]), k
    for seed, sfx in [(float(gd["g"]), ""), (1.0, "_g1")]:
        b = c_oracle.nerf_backward(*args, seed, rows=256)
        for k in ["d_X", "d_ws", "d_bs", "d_target", "d_dists", "d_acc"]:
            assert rel_err(b[k], gd[k + sfx]) <= 1e-6, (k, sfx)      # sigmoid' written as y(1-y)
        if sfx == "":
            assert rel_err(b["d_inter"][:, :N], gd["d_inter"]) <= 1e-6


@pytest.mark.parametrize("path", NERF_GOLDEN, ids=os.path.basename)
def test_f64_restatement_nerf_matches_reference_golden(path):
    gd = load_golden(path)
    R, S = int(gd["R"]), int(gd["S"])
    out = O.nerf_f64(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S,
                     g=float(gd["g"]))
    tol = 1e-5   # the reference itself is fp32: its own rounding noise vs exact math is ~2e-6
    assert rel_err(out["loss"], gd["loss"]) <= tol
    for k_o, k_g in [("color", "color"), ("rgba", "rgba"), ("alpha", "alpha"),
                     ("cumprod", "cumprod"), ("weights", "weights")]:
        assert rel_err(out[k_o], gd[k_g]) <= tol, k_o
    for k in ["d_X", "d_ws", "d_bs", "d_target", "d_dists", "d_acc"]:
        assert rel_err(out[k], gd[k]) <= tol, k
    for l in range(len(gd["dims"]) - 1):
        w = int(gd["dims"][l + 1])
        assert rel_err(out["inter"][l], gd["inter"][l][:, :w]) <= tol
        assert rel_err(out["d_inter"][l], gd["d_inter"][l][:, :w]) <= tol


@pytest.mark.parametrize("path", FIT_GOLDEN, ids=os.path.basename)
def test_oracles_mlp_fit_match_reference_golden(c_oracle, path):
    gd = load_golden(path)
    N = int(gd["N"])
    args = (gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"])
    f = c_oracle.mlp_fit_forward(*args)
    assert np.float32(f["loss"]) == gd["loss"]
    assert np.array_equal(f["inter"][:, :N], gd["inter"])
    b = c_oracle.mlp_fit_backward(*args, float(gd["g"]))
    o64 = O.mlp_fit_f64(*args, g=float(gd["g"]))
    for k in ["d_X", "d_ws", "d_bs", "d_target"]:
        assert rel_err(b[k], gd[k]) <= 1e-6, k
        assert rel_err(o64[k], gd[k]) <= 1e-5, k
    assert rel_err(b["d_inter"][:, :N], gd["d_inter"]) <= 1e-6


def test_gradients_are_linear_in_the_seed():
    """The hosts pass the LOSS as _dreturn (train_nerf.py:477): grads = loss * dloss/dtheta."""
    gd = load_golden(SMALL_NERF[0])
    for k in ["d_ws", "d_bs", "d_X", "d_dists"]:
        assert rel_err(gd[k + "_g1"] * np.float64(gd["g"]), gd[k]) <= 1e-6


def test_compositing_is_the_reference_inclusive_variant():
    """w_0 = alpha_0, w_j = alpha_j * prod_{k<=j} q_k  (not textbook NeRF)."""
    gd = load_golden(SMALL_NERF[0])
    a = gd["alpha"].astype(np.float64)
    q = (1.0 - a) + 1e-10
    C = np.cumprod(q, axis=1)
    C[:, 0] = 1.0
    assert rel_err(a * C, gd["weights"]) <= 1e-6
    assert np.all(gd["cumprod"][:, 0] == 1.0)


@pytest.mark.skipif(not O.have_ref("nerf"), reason="oracle/_ref not built here")
def test_live_reference_agrees_with_c_oracle_on_fresh_seed(c_oracle):
    ref = O.load_ref("nerf")
    for seed, R, S in [(901, 5, 30), (902, 3, 64)]:
        case = O.make_nerf_case(seed, R, S)
        args = (case["X"], case["ws"], case["bs"], case["dims"], case["target"], case["dists"], R, S)
        r = ref.nerf(*args, g="loss")
        f = c_oracle.nerf_forward(*args, rows=256)
        assert f["loss"] == r["loss"] and np.array_equal(f["color"], r["color"])
        b = c_oracle.nerf_backward(*args, float(r["loss"]), rows=256)
        for k in ["d_ws", "d_bs", "d_X", "d_dists", "d_target"]:
            assert rel_err(b[k], r[k]) <= 1e-6
        # by-products the reference leaves behind
        for k in ["d_rgba", "d_alpha", "d_cumprod", "d_weights"]:
            assert not r[k].any()
        assert all(not v.any() for v in r["primal_after_grad"].values())


@pytest.mark.skipif(not O.have_ref("nerf"), reason="oracle/_ref not built here")
def test_reference_chunking_is_additive():
    """Batches beyond the reference's 256-sample capacity are evaluated chunk by chunk; loss and
    weight gradients add (the loss is a plain sum over rays, nerf.py:297-302)."""
    ref = O.load_ref("nerf")
    case = O.make_nerf_case(77, 12, 64)
    tot = O.ref_nerf_chunked(ref, case, g=1.0)
    f64 = O.nerf_f64(case["X"], case["ws"], case["bs"], case["dims"], case["target"],
                     case["dists"], 12, 64, g=1.0)
    assert rel_err(tot["loss"], f64["loss"]) <= 1e-5
    assert rel_err(tot["d_ws"], f64["d_ws"]) <= 1e-5
    assert rel_err(tot["d_bs"], f64["d_bs"]) <= 1e-5


def test_f64_restatement_matches_the_eight_ray_paper_size_golden():
    """BASELINE config 5 (63 -> 8 x 256 -> 4, S = 192) on eight rays: loss and weight gradients summed over eight
    single-ray calls of the paper-size reference build (tests/golden/make_golden_c5_r8.py), unit seed."""
    gd = load_golden(os.path.join(os.path.dirname(NERF_GOLDEN[0]), "wide_c5_r8_s192.npz"))
    R, S = int(gd["R"]), int(gd["S"])
    f = O.nerf_f64(gd["X"], gd["ws"], gd["bs"], gd["dims"], gd["target"], gd["dists"], R, S, g=1.0)
    assert rel_err(f["loss"], gd["loss"]) <= 2e-6
    assert rel_err(f["color"], gd["color"]) <= 2e-6
    assert rel_err(f["d_ws"], gd["d_ws_g1"]) <= 5e-6 and rel_err(f["d_bs"], gd["d_bs_g1"]) <= 5e-6
