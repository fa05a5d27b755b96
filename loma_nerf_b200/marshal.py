"""Zero-copy marshalling for the reference's ragged ctypes ABI.

The reference builds every float** / float*** argument with
mlp_utils.convert_ndim_array_to_ndim_ctypes (/root/reference/mlp_utils.py:33-118): numpy ->
.tolist() -> one ctypes array per row, ~34 ms per (3,256,256) array, and reads
results back element by element (mlp_utils.py:120-164).  These helpers produce the same pointer
shapes as *views* of contiguous numpy arrays: the library writes straight into the numpy memory,
so there is nothing to convert back.
"""
import ctypes
from ctypes import POINTER, c_float, c_int

import numpy as np

c_float_p = POINTER(c_float)
c_float_pp = POINTER(c_float_p)
c_int_p = POINTER(c_int)


class Ragged:
    """Keeps the numpy array and every pointer table alive; `.ptr` is what to pass to ctypes."""

    def __init__(self, arr, ptr, keep):
        self.array, self.ptr, self._keep = arr, ptr, keep


def _rows2(a, elem_p):
    n, stride, base = a.shape[0], a.strides[0], a.ctypes.data
    tab = (elem_p * n)()
    for i in range(n):
        tab[i] = ctypes.cast(base + i * stride, elem_p)
    return tab


def as_ragged(arr):
    """float**/int** (2-D) or float***/int*** (3-D) view of `arr` (float32 / int32, made
    contiguous if needed).  float64 input is converted to float32 like the reference's c_float."""
    a = np.asarray(arr)
    if a.dtype.kind == "f":
        a = np.ascontiguousarray(a, np.float32)
        ep = c_float_p
    else:
        a = np.ascontiguousarray(a, np.int32)
        ep = c_int_p
    if a.ndim == 2:
        tab = _rows2(a, ep)
        return Ragged(a, tab, [tab])
    if a.ndim == 3:
        inner = [_rows2(a[i], ep) for i in range(a.shape[0])]
        tab = (POINTER(ep) * a.shape[0])()
        for i, t in enumerate(inner):
            tab[i] = ctypes.cast(t, POINTER(ep))
        return Ragged(a, tab, inner + [tab])
    raise ValueError("as_ragged: 2-D or 3-D arrays only (got %d-D)" % a.ndim)


def ragged_to_numpy(ptr, shape):
    """Fast replacement for lp_lp_c_float_to_numpy / lp_lp_lp_c_float_to_numpy
    (mlp_utils.py:120-164) for pointers made by the REFERENCE's marshaller: one memmove per row."""
    out = np.empty(shape, np.float32)
    if len(shape) == 2:
        for i in range(shape[0]):
            ctypes.memmove(out[i].ctypes.data, ptr[i], shape[1] * 4)
    elif len(shape) == 3:
        for i in range(shape[0]):
            for j in range(shape[1]):
                ctypes.memmove(out[i, j].ctypes.data, ptr[i][j], shape[2] * 4)
    else:
        raise ValueError("2-D or 3-D only")
    return out
