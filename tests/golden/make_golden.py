#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref, built by oracle/build_ref.py
from /root/reference's own loma programs with its own compiler).  Run in the build container
(where /root/reference exists); the .npz files are committed so the GPU box and the CPU suite can
check the oracle restatement and the CUDA path without the reference tree.

Each file stores the inputs (seeded, oracle.make_*_case) and every output the reference produced:
forward scratch arrays, loss, and all d_* buffers of the grad call with _dreturn = loss (as the
reference hosts pass it, train_nerf.py:477 / fit_img.py:497) -- plus _dreturn = 1 gradients.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

NERF_CASES = [
    # name, seed, R, S, E, width, layers, stratified, library
    ("nerf_ref_r4_s30", 215, 4, 30, 5, 30, 3, False, "nerf"),   # the reference's own chunk
    ("nerf_r8_s30", 216, 8, 30, 5, 30, 3, True, "nerf"),
    ("nerf_r8_s32", 217, 8, 32, 5, 30, 3, True, "nerf"),
    ("nerf_c2_r4_s64", 218, 4, 64, 5, 30, 3, True, "nerf"),     # BASELINE config 2 shape per call
    ("nerf_r2_s128", 219, 2, 128, 5, 30, 3, True, "nerf"),
    ("nerf_r1_s192", 220, 1, 192, 5, 30, 3, True, "nerf"),
    ("nerf_r3_s7_w16", 221, 3, 7, 2, 16, 2, True, "nerf"),      # ragged: odd S, 2 layers
    ("nerf_c5_r1_s192", 222, 1, 192, 10, 256, 9, True, "nerf_big"),  # paper-size, 63->8x256->4
]
FIT_CASES = [
    ("fit_c1_n256", 230, 256, 5, 16, 3),   # fit_img.py chunk: 22->16->16->3, 256 pixels
    ("fit_n100_w8", 231, 100, 3, 8, 2),
]


def main():
    build_ref.build(verbose=True)
    for name, seed, R, S, E, width, layers, strat, libname in NERF_CASES:
        ref = O.load_ref(libname)
        case = O.make_nerf_case(seed, R, S, E=E, width=width, n_layers=layers, stratified=strat)
        out = ref.nerf(case["X"], case["ws"], case["bs"], case["dims"], case["target"],
                       case["dists"], R, S, g="loss")
        out1 = ref.nerf(case["X"], case["ws"], case["bs"], case["dims"], case["target"],
                        case["dists"], R, S, g=1.0)
        N, mo = R * S, case["ws"].shape[2]
        save = dict(seed=seed, R=R, S=S, E=E, width=width, layers=layers, stratified=strat,
                    X=case["X"], ws=case["ws"], bs=case["bs"], dims=case["dims"],
                    target=case["target"], dists=case["dists"], rays_o=case["rays_o"],
                    rays_d=case["rays_d"], t=case["t"],
                    loss=out["loss"], inter=out["inter"][:, :N, :mo].copy(), rgba=out["rgba"],
                    alpha=out["alpha"], cumprod=out["cumprod"], weights=out["weights"],
                    color=out["color"], g=np.float32(out["g"]))
        for k in ["d_X", "d_ws", "d_bs", "d_target", "d_dists", "d_acc"]:
            save[k] = out[k]
            save[k + "_g1"] = out1[k]
        save["d_inter"] = out["d_inter"][:, :N, :mo].copy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print("wrote", name, "loss", out["loss"])
    ref = O.load_ref("mlp_fit")
    for name, seed, N, E, width, layers in FIT_CASES:
        case = O.make_fit_case(seed, N, E=E, width=width, n_layers=layers)
        out = ref.mlp_fit(case["X"], case["ws"], case["bs"], case["dims"], case["target"], g="loss")
        out1 = ref.mlp_fit(case["X"], case["ws"], case["bs"], case["dims"], case["target"], g=1.0)
        mo = case["ws"].shape[2]
        save = dict(seed=seed, N=N, E=E, width=width, layers=layers, X=case["X"], ws=case["ws"],
                    bs=case["bs"], dims=case["dims"], target=case["target"], xy=case["xy"],
                    loss=out["loss"], inter=out["inter"][:, :N, :mo].copy(), g=np.float32(out["g"]))
        for k in ["d_X", "d_ws", "d_bs", "d_target"]:
            save[k] = out[k]
            save[k + "_g1"] = out1[k]
        save["d_inter"] = out["d_inter"][:, :N, :mo].copy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print("wrote", name, "loss", out["loss"])
    # the reference's only known-answer test on this path (fit_img.py:363-374)
    c = ref.mult_a_b(np.array([[1, 2], [3, 4], [5, 6]], np.float32),
                     np.array([[100], [200]], np.float32))
    assert np.allclose(c, [[500], [1100], [1700]]), c
    print("mult_a_b KAT ok", c.ravel())


if __name__ == "__main__":
    main()
