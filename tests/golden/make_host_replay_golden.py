#!/usr/bin/env python3
"""Generate tests/golden/host_replay.npz: the reference hosts' call sequence (tests/host_replay.py,
restating train_nerf.py:209-499 and fit_img.py:355-532) driven against the REAL reference libraries
oracle/_ref/nerf.so / mlp_fit.so (the reference's own loma programs compiled by its own compiler,
oracle/build_ref.py).  Run in the build container; the .npz is committed."""
import ctypes
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import host_replay  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import oracle as O  # noqa: E402


def ref_compiler(name):
    """A `compiler` whose compile() hands back the real reference library with the argtypes
    loma_public/compiler.py:262-276 sets."""
    mod = types.SimpleNamespace()
    mod.compile = lambda code, target="c", output_filename=None: ({}, O.set_compat_argtypes(
        ctypes.CDLL(os.path.join(O.REF_DIR, name + ".so"))))
    return mod


def main():
    build_ref.build(verbose=False)
    big = lambda fn, *a: O.run_big_stack(lambda: fn(*a))  # noqa: E731  (16 MB of tape on the stack)
    nerf = host_replay.run_nerf_host(ref_compiler("nerf"), call=big, n_chunks=3)
    fit = host_replay.run_fit_host(ref_compiler("mlp_fit"), call=big, n_chunks=3)
    save = {"nerf_" + k: v for k, v in nerf.items()}
    save.update({"fit_" + k: v for k, v in fit.items()})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_replay.npz"), **save)
    print("nerf losses", nerf["loss"], "fit losses", fit["loss"])


if __name__ == "__main__":
    main()
