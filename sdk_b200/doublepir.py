"""Host-side mirror of DoublePIR's packed matvec (lib/doublepir/src/matrix/kernels.rs:118-178)."""
import ctypes as C

import numpy as np

from ._lib import LIB, check, B200PirError  # noqa: F401


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


def matrix_mul_vec_packed_rows(a, row_begin, row_count, b):
    """matrix_mul_vec_packed(db.rows(start, n), q) (doublepir.rs:301)."""
    out = np.zeros(row_count, dtype=np.uint32)
    check(LIB.b200pir_dpir_matvec_packed_rows(a._h, row_begin, row_count, b.ctypes.data, out.ctypes.data))
    return out


def matrix_mul_transposed_packed(a, a_rows, a_cols, b, b_rows, b_cols, basis=10, compression=3, device=0):
    """kernels.rs:256-278."""
    assert basis == 10 and compression == 3
    out = np.zeros(a_rows * b_rows, dtype=np.uint32)
    check(LIB.b200pir_dpir_matrix_mul_transposed_packed(device, a.ctypes.data, a_rows, a_cols, b.ctypes.data, b_rows, b_cols,
                                                        out.ctypes.data))
    return out


def transpose_expand_concat_cols_squish(a, rows, cols, modulus, delta, concat, basis=10, d=3, device=0):
    """matrix/indexing.rs:117-143 -> (out, out_rows, out_cols)."""
    assert basis == 10 and d == 3
    orows, ocols = cols * delta * concat, (rows // concat + 2) // 3
    out = np.zeros(orows * ocols, dtype=np.uint32)
    r, c = C.c_uint64(), C.c_uint64()
    check(LIB.b200pir_dpir_transpose_expand_concat_cols_squish(device, a.ctypes.data, rows, cols, modulus, delta, concat,
                                                               out.ctypes.data, C.byref(r), C.byref(c)))
    return out, r.value, c.value


def answer(db, queries, h_1, a_2_transpose, p, delta, x, ne):
    """DoublePIR server answer (doublepir.rs:246-350, raw_data = None, chunk_idx = None).
    db: PackedMatrix; queries: list of [q_1, q_2, ...]; h_1 / a_2_transpose: (array, rows, cols)."""
    nq = len(queries)
    batch = db.rows // nq
    parts, last = [], 0
    for b, q in enumerate(queries):
        bs = db.rows - last if b == nq - 1 else batch
        parts.append(matrix_mul_vec_packed_rows(db, last, bs, q[0]))
        last += bs
    a_1 = np.concatenate(parts)
    a_1, r1, c1 = transpose_expand_concat_cols_squish(a_1, db.rows, 1, p, delta, x)
    a2, a2_rows, a2_cols = a_2_transpose
    msg = [matrix_mul_transposed_packed(a_1, r1, c1, a2, a2_rows, a2_cols)]
    h, h_rows, h_cols = h_1
    hm = PackedMatrix(h, h_rows, h_cols)
    am = PackedMatrix(a_1, r1, c1)
    for q in queries:
        for j in range(ne // x):
            msg.append(matrix_mul_vec_packed(hm, q[1 + j]))
            msg.append(matrix_mul_vec_packed(am, q[1 + j]))
    hm.close()
    am.close()
    return msg


def matmul(a, b, device=0):
    """`&Matrix * &Matrix` (matrix/ops.rs:169-191), wrapping u32, for a left operand with small signed entries (|a| < 2^15):
    exact 8-bit limb products on the tensor cores."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[0]:
        raise ValueError("a.cols %r b.rows %r" % (a.shape, b.shape))
    out = np.zeros((a.shape[0], b.shape[1]), dtype=np.uint32)
    check(LIB.b200pir_dpir_matmul(device, a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[1], out.ctypes.data))
    return out


def setup(db, a1, a2, p, delta, x, device=0):
    """doublepir.rs:76-108 setup(): returns dict(db_squished, h1_squished, a2_t, h2) = (server_state pieces, hint).
    db: l x m with entries centred in [-p/2, p/2) (wrapping u32); a1: m x n; a2: (l/x) x n."""
    db = np.ascontiguousarray(db, dtype=np.uint32)
    a1 = np.ascontiguousarray(a1, dtype=np.uint32)
    a2 = np.ascontiguousarray(a2, dtype=np.uint32)
    l, m = db.shape
    n = a1.shape[1]
    if a1.shape[0] != m or a2.shape != (l // x, n) or l % x:
        raise ValueError("shapes: db (l, m), a1 (m, n), a2 (l/x, n)")
    lx = l // x
    rows1 = n * delta * x
    out = dict(db_squished=np.zeros((l, (m + 2) // 3), dtype=np.uint32), h1_squished=np.zeros((rows1, (lx + 2) // 3), dtype=np.uint32),
               a2_t=np.zeros((n, lx + (3 - lx % 3) % 3), dtype=np.uint32), h2=np.zeros((rows1, n), dtype=np.uint32))
    check(LIB.b200pir_dpir_setup(device, db.ctypes.data, l, m, a1.ctypes.data, n, a2.ctypes.data, p, delta, x,
                                 out["db_squished"].ctypes.data, out["h1_squished"].ctypes.data, out["a2_t"].ctypes.data,
                                 out["h2"].ctypes.data))
    return out
