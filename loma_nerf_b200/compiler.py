"""Drop-in for the reference's `compiler.compile` on the NeRF / MLP-fit path.

The reference hosts do (train_nerf.py:209-213, fit_img.py:355-361)

    _, lib = compiler.compile(open("scripts/nerf.py").read(), target="c", output_filename="_code/nerf")
    nerf_evaluate_and_march = lib.nerf_evaluate_and_march
    grad_nerf_evaluate_and_march = lib.grad_nerf_evaluate_and_march

and every run regenerates and overwrites the `.so`.  The hosts find the module as a TOP-LEVEL
`compiler` (they append loma_public/ to sys.path and `import compiler`, train_nerf.py:4-12,
fit_img.py:4-9), so this file works both ways: as `loma_nerf_b200.compiler` and as a plain
`compiler` module found through PYTHONPATH=<repo>/loma_nerf_b200/dropin (a directory holding only
a re-export of this file) or PYTHONPATH=<repo>/loma_nerf_b200.  Either way `compile()` returns
`({}, libloma_nerf_b200.so)` with the same argtypes / restype that
/root/reference/loma_public/compiler.py:262-276 would have set, so the hosts run unmodified on the
GPU.  Nothing is compiled from the loma source: only the program's function names are looked at, to
refuse programs this library does not implement.
"""
import ctypes
import os
import re
import sys
from ctypes import POINTER, c_float, c_int

if __package__:
    from . import _lib
else:
    # imported as a top-level module (`import compiler` with this directory on sys.path): make the
    # package importable by its real name and use its loader, so there is one library handle
    _repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if _repo not in sys.path:
        sys.path.append(_repo)
    from loma_nerf_b200 import _lib

c_float_p = POINTER(c_float)
c_float_pp = POINTER(c_float_p)
c_float_ppp = POINTER(c_float_pp)
c_int_p = POINTER(c_int)
c_int_pp = POINTER(c_int_p)

# scripts/nerf.py:1-22 -- In[Array[Array[float]]] -> float**, In[int] -> int, ...
NERF_ARGTYPES = [c_float_pp, c_int, c_int, c_float_ppp, c_float_pp, c_float_pp, c_int, c_int, c_int,
                 c_int_pp, c_int_pp, c_int_pp, c_float_ppp, c_float_ppp, c_int, c_float_pp,
                 c_float_pp, c_float_pp, c_float_pp, c_float_pp]
# scripts/mlp_fit.py:1-16
FIT_ARGTYPES = [c_float_pp, c_int, c_int, c_float_pp, c_float_ppp, c_float_pp, c_float_pp, c_int,
                c_int, c_int, c_int_pp, c_int_pp, c_int_pp, c_float_ppp]
MULT_ARGTYPES = [c_float_pp, c_int, c_int, c_float_pp, c_int, c_int, c_float_pp]


def grad_argtypes(argtypes):
    """loma_public/reverse_diff.py:504-517: each In argument is followed by its adjoint
    (int -> int*), and a trailing float _dreturn closes the list."""
    out = []
    for t in argtypes:
        out += [t, c_int_p if t is c_int else t]
    return out + [c_float]


SUPPORTED = {
    "nerf_evaluate_and_march": (NERF_ARGTYPES, c_float),
    "grad_nerf_evaluate_and_march": (grad_argtypes(NERF_ARGTYPES), None),
    "mlp_fit": (FIT_ARGTYPES, c_float),
    "grad_mlp_fit": (grad_argtypes(FIT_ARGTYPES), None),
    "mult_a_b": (MULT_ARGTYPES, None),
}


def bind(lib=None):
    """Set the reference's argtypes on the five compat symbols of the (loaded) library."""
    lib = lib or ctypes.CDLL(_lib.LIB_PATH)
    for name, (argtypes, restype) in SUPPORTED.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


def compile(loma_code, target="c", output_filename=None, opencl_context=None, opencl_device=None,  # noqa: A001
            opencl_command_queue=None, print_error=True):
    """Same signature as the reference's compile(); returns (ctypes_structs, lib)."""
    if target not in ("c", "ispc"):
        raise ValueError("loma_nerf_b200.compiler: only the reference's C-ABI targets are mirrored, got %r" % target)
    _lib.load()  # raises LibraryMissing when the CUDA library was not built: no CPU fallback
    names = set(re.findall(r"^def\s+(\w+)\s*\(", loma_code, flags=re.M))
    names |= set(re.findall(r"^(\w+)\s*=\s*rev_diff\(", loma_code, flags=re.M))
    unknown = sorted(n for n in names if n not in SUPPORTED)
    if unknown:
        raise NotImplementedError(
            "loma_nerf_b200 implements nerf_evaluate_and_march / mlp_fit / mult_a_b and their "
            "rev_diff gradients only; the program also defines %s" % ", ".join(unknown))
    return {}, bind()
