#!/usr/bin/env python3
"""tests/golden/wide_c5_r8_s192.npz: BASELINE config 5 (63 -> 8 x 256 -> 4, 192 samples per ray) on EIGHT rays from the
real reference (paper-size build of scripts/nerf.py, oracle/build_ref.py), unit seed.  Run under `ulimit -s unlimited`."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import build_ref
from oracle import oracle as O
build_ref.build(verbose=False)
ref = O.load_ref("nerf_big")
R, S, E = 8, 192, 10
case = O.make_nerf_case(223, R, S, E=E, width=256, n_layers=9, stratified=True)
t0 = time.time()
# the paper-size reference build holds one ray per call (1.9 GB of tape on the stack): loop over
# rays, sum loss / gradients in float64 (the loss is a plain sum over rays, nerf.py:297-302)
tot = O.ref_nerf_chunked(ref, case, g=1.0, rays_per_call=1)
print("took", time.time() - t0, "loss", tot["loss"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wide_c5_r8_s192.npz"), seed=223, R=R, S=S, E=E, width=256, layers=9, stratified=True,
                    X=case["X"], ws=case["ws"], bs=case["bs"], dims=case["dims"], target=case["target"], dists=case["dists"],
                    rays_o=case["rays_o"], rays_d=case["rays_d"], t=case["t"], loss=np.float64(tot["loss"]), color=tot["color"],
                    d_ws_g1=tot["d_ws"].astype(np.float32), d_bs_g1=tot["d_bs"].astype(np.float32), g=np.float32(1.0))
