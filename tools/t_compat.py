"""Latency of the five drop-in symbols on the hosts' own chunk shape (4 rays x 30 samples, train_nerf.py:275-499) against the
real reference library (oracle/_ref/nerf.so) on one host core: python tools/t_compat.py"""
import ctypes, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as O
from loma_nerf_b200 import _lib
for R, S in ((4, 30), (4, 64)):
    case = O.make_nerf_case(215, R, S, stratified=False)
    args = (case["X"], case["ws"], case["bs"], [int(v) for v in case["dims"]], case["target"], case["dists"], R, S)
    ours = O.CompatCaller(ctypes.CDLL(_lib.LIB_PATH), big_stack=False)
    for name, caller in (("ours (GPU, compat symbols)", ours),) + ((("reference nerf.so (1 core)", O.load_ref("nerf")),) if O.have_ref("nerf") else ()):
        def body():
            for _ in range(3):
                caller.nerf(*args, g="loss")
            t0 = time.perf_counter(); n = 30
            for _ in range(n):
                caller.nerf(*args, g="loss")
            return (time.perf_counter() - t0) / n
        if "reference" in name:
            box = {}
            O.run_big_stack(lambda: box.setdefault("t", body()), 256 << 20)
            t = box["t"]
        else:
            t = body()
        print("R=%d S=%d %-32s %.3f ms per forward + grad call pair (marshalling included) = %.0f samples/s" % (R, S, name, t * 1e3, R * S / t))
