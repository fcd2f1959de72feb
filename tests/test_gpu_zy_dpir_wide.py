"""DoublePIR's short-and-wide database shape (second case): the reference's own end-to-end test uses 2^24
one-bit entries, which approx_database_dims lays out as l = 29 rows x m = 65536 columns (doublepir.rs:471-483,
database.rs:376-418): 21846 packed words per row, more than fits beside `b` in shared memory, served by the wide-row kernel.
Sorted late: added after the last GPU session of the round."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_parity import _gpu

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("rows,cols", [(29, 21846), (32, 17067), (3, 40001), (1, 17068)])
def test_dpir_matvec_wide_rows_match_oracle(rows, cols):
    _gpu()
    import sdk_b200.doublepir as D
    rng = np.random.default_rng(rows * 7 + cols)
    a = rng.integers(0, 2**30, rows * cols, dtype=np.uint32)
    a[:cols] = 0x3FFFFFFF                                       # first row: every 10-bit field at its maximum
    b = rng.integers(0, 2**32, 3 * cols, dtype=np.uint32)
    b[:7] = 0xFFFFFFFF
    m = D.PackedMatrix(a, rows, cols)
    try:
        assert np.array_equal(D.matrix_mul_vec_packed(m, b), O.dpir_matvec_packed(a, b, rows, cols))
    finally:
        m.close()
