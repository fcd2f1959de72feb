"""BASELINE config #5 at poly_len = 4096: the 512-thread cooperative transform (sdk_b200/csrc/ntt_core4096.cuh) against the
oracle's scalar transforms instantiated at 4096.  The pass logic is already checked thread by thread on the CPU
(tests/cpp/ntt_core4096_emul.cpp, part of the CPU suite)."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_parity import setup_case, Q0, Q1

pytestmark = [pytest.mark.gpu]


def test_ntt4096_forward_inverse_match_oracle():
    S, P, cl, pp, db, G, gdb, gpp = setup_case("T")
    rng = np.random.default_rng(41)
    count = 19
    v = np.empty((count, 2, 4096), dtype=np.uint64)
    v[:, 0, :] = rng.integers(0, Q0, (count, 4096), dtype=np.uint64)
    v[:, 1, :] = rng.integers(0, Q1, (count, 4096), dtype=np.uint64)
    v[0] = 0
    v[1, 0, :] = Q0 - 1
    v[1, 1, :] = Q1 - 1
    v[2, 0, :] = rng.integers(0, 4 * Q0, 4096, dtype=np.uint64)        # lazy-range input
    ref = v.copy().reshape(-1)
    assert O.LIB.orc_ntt4096(O._p64(ref), count, 0) == 0
    got = v.copy().reshape(-1)
    S.ntt4096(G, got)
    assert np.array_equal(got, ref)
    ref2 = ref.copy()
    assert O.LIB.orc_ntt4096(O._p64(ref2), count, 1) == 0
    S.ntt4096(G, got, inverse=True)
    assert np.array_equal(got, ref2)
    back = v.reshape(-1).copy()
    back.reshape(count, 2, 4096)[2, 0, :] %= np.uint64(Q0)
    assert np.array_equal(got, back)                                    # round trip = canonical input
