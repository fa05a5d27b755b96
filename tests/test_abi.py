"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/loma_nerf_b200.h declares, the Python struct mirror matches the C layout, the
`compiler.compile` shim binds the reference's argtypes, and nothing falls back to the CPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "loma_nerf_b200.h")


@pytest.fixture(scope="module")
def lib():
    from loma_nerf_b200 import build, _lib
    build.build_library()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"LNB_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_header_declares_the_reference_symbols():
    syms = declared_symbols()
    for s in ["nerf_evaluate_and_march", "grad_nerf_evaluate_and_march", "mlp_fit", "grad_mlp_fit",
              "mult_a_b", "lnb_nerf_step", "lnb_fit_step", "lnb_nerf_step_host"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_make_dfloat_pair_constructor(lib):
    """autodiff.py:199-208: the (value, tangent) struct constructor every loma-generated library carries."""
    class DFloat(ctypes.Structure):
        _fields_ = [("val", ctypes.c_float), ("dval", ctypes.c_float)]
    assert "make__dfloat" in declared_symbols()
    lib.make__dfloat.restype = DFloat
    lib.make__dfloat.argtypes = [ctypes.c_float, ctypes.c_float]
    r = lib.make__dfloat(1.5, -2.25)
    assert (r.val, r.dval) == (1.5, -2.25)


def test_struct_mirror_matches_c_layout(lib):
    from loma_nerf_b200 import _lib as L
    out = (ctypes.c_int * 16)()
    n = lib.lnb_struct_layout(out, 16)
    A = L.LnbStepArgs
    mine = [ctypes.sizeof(L.LnbMlp), ctypes.sizeof(A), A.X.offset, A.inter.offset, A.rgba.offset,
            A.loss.offset, A.want_grad.offset, A.d_ws.offset, A.path.offset, A.rays_o.offset,
            A.pe_bands.offset, A.cam.offset, ctypes.sizeof(L.LnbCamera), L.LnbCamera.pixels.offset]
    assert n == len(mine)
    assert list(out[:n]) == mine
    assert lib.lnb_abi_version() == 4


def test_counter_based_jitter_generator_matches_its_restatement(lib):
    """lnb_uniform (host-callable, the generator camera mode's stratified depths use): SplitMix64 finaliser over
    (seed, pixel, sample), 24 bits -- restated here in numpy uint64 arithmetic."""
    import numpy as np

    def mix(seed, q, s):
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.uint64(q) * np.uint64(4096) + np.uint64(s) + np.uint64(1))
            z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
            z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
        return float(int(z) >> 40) / 16777216.0
    rng = np.random.default_rng(3)
    us = []
    for _ in range(200):
        seed, q, s = int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 640000)), int(rng.integers(0, 192))
        u = lib.lnb_uniform(seed, q, s)
        assert u == mix(seed, q, s) and 0.0 <= u < 1.0
        us.append(u)
    assert 0.4 < np.mean(us) < 0.6


def test_compile_shim_binds_reference_argtypes(lib):
    from loma_nerf_b200 import compiler
    src = "def nerf_evaluate_and_march(a):\n    pass\ngrad_nerf_evaluate_and_march = rev_diff(nerf_evaluate_and_march)\n"
    structs, l2 = compiler.compile(src, target="c", output_filename="_code/nerf")
    assert structs == {}
    assert len(l2.nerf_evaluate_and_march.argtypes) == 20          # scripts/nerf.py:1-22
    assert len(l2.grad_nerf_evaluate_and_march.argtypes) == 41     # reverse_diff.py:504-517
    assert len(l2.mlp_fit.argtypes) == 14 and len(l2.grad_mlp_fit.argtypes) == 29
    assert l2.nerf_evaluate_and_march.restype is ctypes.c_float
    assert l2.grad_mlp_fit.restype is None
    with pytest.raises(NotImplementedError):
        compiler.compile("def something_else(x):\n    pass\n")


def test_reference_sources_are_accepted_by_the_shim(lib):
    """The function names of the reference's two loma programs (scripts/nerf.py:1,306;
    scripts/mlp_fit.py:1,150,174), as the hosts pass them."""
    from loma_nerf_b200 import compiler
    nerf = "def nerf_evaluate_and_march(x):\n    pass\n\ngrad_nerf_evaluate_and_march = rev_diff(nerf_evaluate_and_march)\n"
    fit = ("def mlp_fit(x):\n    pass\n\ndef mult_a_b(a):\n    pass\n\n"
           "grad_mlp_fit = rev_diff(mlp_fit)\n# fwd_mlp_fit = fwd_diff(mlp_fit)\n")
    compiler.compile(nerf)
    compiler.compile(fit)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must refuse to run, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from loma_nerf_b200 import api
    assert lib.lnb_device_count() == 0
    with pytest.raises(api.LnbError):
        api.Context()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "loma_nerf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "nerf_oracle" not in txt and "liboracle" not in txt, f


def test_marshal_views_are_zero_copy():
    import numpy as np
    from loma_nerf_b200 import marshal
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    r = marshal.as_ragged(a)
    assert r.ptr[1][2][3] == 23.0
    r.ptr[1][2][3] = -1.0
    assert r.array[1, 2, 3] == -1.0
    b = marshal.as_ragged(np.arange(6, dtype=np.int32).reshape(3, 2))
    assert b.ptr[2][1] == 5
    assert np.array_equal(marshal.ragged_to_numpy(r.ptr, (2, 3, 4)), r.array)


def test_compat_symbols_fail_loudly_without_a_gpu(lib):
    """On a box without CUDA the drop-in symbols must signal failure the way the reference hosts
    can see it (NaN loss / NaN outputs, train_nerf.py:486-489), never compute on the host."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from loma_nerf_b200 import compiler, marshal
    l2 = compiler.bind()
    a = marshal.as_ragged(np.array([[1, 2], [3, 4], [5, 6]], np.float32))
    b = marshal.as_ragged(np.array([[100], [200]], np.float32))
    c = marshal.as_ragged(np.zeros((3, 1), np.float32))
    l2.mult_a_b(a.ptr, 3, 2, b.ptr, 2, 1, c.ptr)
    assert np.isnan(c.array).all()
