#!/usr/bin/env python3
"""Build the REAL reference implementation of the hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may load what this script produces.

What it does: imports the reference's own loma compiler *where it lies* under
/root/reference/loma_public (nothing is copied into this repository), feeds it
the reference's own loma programs scripts/nerf.py and scripts/mlp_fit.py with
target="c" exactly as the reference hosts do (train_nerf.py:209-213,
fit_img.py:355-361), and lets it run `gcc -shared -fPIC -O2`
(loma_public/compiler.py:154).  Outputs go ONLY to oracle/_ref/ (git-ignored,
not gpurun-ignored, so the .so files travel to the GPU box):

    oracle/_ref/nerf.so        nerf_evaluate_and_march + grad_ (stock capacity:
                               <=3 layers, <=256 samples per call, widths <=32)
    oracle/_ref/mlp_fit.so     mlp_fit + grad_mlp_fit + mult_a_b
    oracle/_ref/nerf_big.so    same program with ONLY the `max_iter :=` literals
                               enlarged (9 layers, width 256, 192 samples/ray,
                               1 ray per call) for the paper-size config
    oracle/_ref/*_gen.c        the generated C text, kept for reading

Third-party modules the reference compiler imports but this image lacks
(asdl, yapf, gpuctypes) are stubbed; they are not on the C-backend path
(SURVEY.md Appendix A).  /root/reference does not exist on the GPU box: this
script is a no-op there and the prebuilt files are used.
"""
import contextlib
import io
import os
import re
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("LOMA_NERF_REFERENCE", "/root/reference")


def _stub_missing_modules():
    for name in ["asdl", "gpuctypes", "gpuctypes.opencl", "yapf", "yapf.yapflib",
                 "yapf.yapflib.yapf_api"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["yapf.yapflib.yapf_api"].FormatCode = lambda *a, **k: (a[0], False)
    sys.modules["gpuctypes"].opencl = sys.modules["gpuctypes.opencl"]
    stub = types.ModuleType("asdl_gen")
    stub.ADT = lambda *a, **k: None
    sys.modules["asdl_gen"] = stub


def _enlarge_max_iter(src):
    """Paper-size variant: change ONLY `max_iter := N` literals."""
    def repl(m):
        n = int(m.group(1))
        new = {3: 9, 256: 192, 32: 256, 5: 5, 500: 500}[n]
        return "max_iter := %d" % new
    out = re.sub(r"max_iter := (\d+)", repl, src)
    # the per-ray sample loops (`< num_samples`) were 32 -> need 192, not 256;
    # 256 is a safe over-capacity, left as is (capacity only sizes the tapes).
    return out


def build(verbose=False):
    if not os.path.isdir(os.path.join(REF, "loma_public")):
        if verbose:
            print("build_ref: %s absent; using prebuilt oracle/_ref" % REF)
        return False
    os.makedirs(OUT, exist_ok=True)
    _stub_missing_modules()
    sys.path.insert(0, os.path.join(REF, "loma_public"))
    sys.dont_write_bytecode = True
    import ir  # noqa: E402  (reference module)
    ir.generate_asdl_file = lambda: None  # _asdl/loma.py is shipped pre-generated
    import compiler  # noqa: E402  (reference module)

    jobs = [
        ("nerf", open(os.path.join(REF, "scripts", "nerf.py")).read()),
        ("mlp_fit", open(os.path.join(REF, "scripts", "mlp_fit.py")).read()),
    ]
    if os.environ.get("LOMA_NERF_BUILD_BIG", "1") == "1":
        jobs.append(("nerf_big", _enlarge_max_iter(jobs[0][1])))
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        for name, src in jobs:
            so = os.path.join(OUT, name + ".so")
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                compiler.compile(src, target="c", output_filename=os.path.join(OUT, name))
            text = buf.getvalue()
            k = text.find("Generated C code:")
            if k >= 0:
                with open(os.path.join(OUT, name + "_gen.c"), "w") as f:
                    f.write(text[k + len("Generated C code:"):])
            if not os.path.exists(so):
                raise RuntimeError("reference compiler did not produce " + so)
            if verbose:
                print("build_ref: built", so)
    finally:
        os.chdir(cwd)
    return True


if __name__ == "__main__":
    build(verbose=True)
