"""loma_nerf_b200 -- B200-native (sm_100a) implementation of loma-nerf's NeRF / coordinate-MLP hot path.

The product is the CUDA shared library `libloma_nerf_b200.so` (built from `csrc/`, C ABI in
`include/loma_nerf_b200.h`); this package is the thin host-side mirror of the reference's Python
interface to it:

* `compiler.compile(...)`  -- drop-in for /root/reference/loma_public/compiler.py:70-278 that
  hands the reference hosts (train_nerf.py, fit_img.py) our library with the same argtypes;
* `api.Context`            -- the flat (contiguous-buffer) API used for training / rendering;
* `marshal`                -- zero-copy replacements for mlp_utils.convert_ndim_array_to_ndim_ctypes;
* `sharding`               -- ray sharding + gradient all-reduce across the GPUs of one box.

There is no CPU fallback anywhere in this package.
"""
from ._lib import LIB_PATH, LibraryMissing, load  # noqa: F401

__all__ = ["LIB_PATH", "LibraryMissing", "load"]
