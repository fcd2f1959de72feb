"""Host-side mirror of DoublePIR's packed matvec (lib/doublepir/src/matrix/kernels.rs:118-178)."""
import ctypes as C

import numpy as np

from ._lib import LIB, check


class PackedMatrix:
    """A squished database matrix (`Matrix` of u32, 3 x 10-bit per word; squish.rs:53-70) resident in HBM."""

    def __init__(self, a=None, rows=None, cols=None, device=0, synthetic_seed=None):
        h = C.c_void_p()
        if a is not None:
            if a.dtype != np.uint32 or not a.flags["C_CONTIGUOUS"] or a.size != rows * cols:
                raise TypeError("a must be a C-contiguous uint32 array of rows*cols words")
            check(LIB.b200pir_dpir_create(device, a.ctypes.data, rows, cols, C.byref(h)))
        else:
            check(LIB.b200pir_dpir_create_synthetic(device, rows, cols, int(synthetic_seed), C.byref(h)))
        self._h, self.rows, self.cols = h, rows, cols

    def close(self):
        if getattr(self, "_h", None):
            LIB.b200pir_dpir_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def matrix_mul_vec_packed(a, b, basis=10, compression=3):
    """kernels.rs:118-178: asserts a.cols * compression == b.rows, basis == 10, compression == 3."""
    assert basis == 10 and compression == 3
    if b.dtype != np.uint32 or b.size != a.cols * compression:
        raise ValueError("a.cols %d compression %d b.rows %d" % (a.cols, compression, b.size))
    out = np.zeros(a.rows, dtype=np.uint32)
    check(LIB.b200pir_dpir_matvec_packed(a._h, b.ctypes.data, out.ctypes.data))
    return out
