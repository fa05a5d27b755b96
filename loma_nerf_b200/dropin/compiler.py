"""`import compiler` for the unmodified reference hosts (train_nerf.py:12, fit_img.py:9).

    PYTHONPATH=<repo>/loma_nerf_b200/dropin python train_nerf.py

This directory holds nothing else, so no other top-level module name of the hosts is shadowed.
The hosts append loma_public/ to sys.path (train_nerf.py:4-6), PYTHONPATH entries come first, so
this module wins and `compiler.compile(...)` hands them libloma_nerf_b200.so with the reference's
argtypes (see loma_nerf_b200/compiler.py).
"""
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo not in sys.path:
    sys.path.append(_repo)

from loma_nerf_b200.compiler import *  # noqa: E402,F401,F403
from loma_nerf_b200.compiler import bind, compile, grad_argtypes  # noqa: E402,F401,A004
