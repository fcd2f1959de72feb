"""The arithmetic identity behind DoublePIR's setup() GEMM on the INT8 tensor cores (dpir_gemm.cu), checked in numpy without a
GPU: a = a0 + 2^8 a1 (a0 unsigned byte, a1 signed byte, |a| < 2^15), b = four unsigned bytes; modulo 2^32
    a * b = sum over i + j <= 3 of a_i b_j 2^(8 (i + j)),
the seven partial products accumulated over K into four s32 accumulators (one per shift) that WRAP, recombined with shifts.
A wrapped accumulator is still exact modulo 2^32, and the accumulator of shift 8 s is only needed modulo 2^(32 - 8 s)."""
import numpy as np
import pytest


def limb_gemm(a, b):
    """a: (M, K) int64 in [-2^15, 2^15); b: (K, N) int64 in [0, 2^32) -> (M, N) u32, the way the kernel computes it"""
    a0 = a & 255
    a1 = (a - a0) >> 8
    assert a1.min() >= -128 and a1.max() <= 127
    planes = [(b >> (8 * j)) & 255 for j in range(4)]
    wrap = lambda v: ((v + 2**31) % 2**32) - 2**31                           # what an s32 accumulator holds
    acc = [np.zeros((a.shape[0], b.shape[1]), dtype=np.int64) for _ in range(4)]
    for i, ai in ((0, a0), (1, a1)):
        for j in range(4):
            if i + j <= 3:
                # int64 matmul is exact here (|products| < 2^16, K < 2^31); wrapping once at the end equals wrapping after every
                # k-step because wrap() is the reduction modulo 2^32 into the signed range
                acc[i + j] = wrap(acc[i + j] + ai @ planes[j])
    out = np.zeros_like(acc[0])
    for s in range(4):
        out = (out + ((acc[s] % 2**32) << (8 * s))) % 2**32                  # epilogue: shifts and wrapping u32 adds
    return out.astype(np.uint32)


@pytest.mark.parametrize("k", [1, 31, 257, 40000, 100000])
def test_limb_decomposition_is_exact_mod_2_32(k):
    rng = np.random.default_rng(k)
    m, n = 6, 5
    a = rng.integers(-(1 << 15), 1 << 15, (m, k), dtype=np.int64)
    a[0, :] = 255                      # largest unsigned low limb, every k: the shift-0 accumulator overflows s32 for k > 33000
    a[1, :] = -(1 << 15)
    a[2, :] = (1 << 15) - 1
    b = rng.integers(0, 1 << 32, (k, n), dtype=np.int64)
    b[:, 0] = (1 << 32) - 1
    ref = np.zeros((m, n), dtype=np.uint64)
    for c in range(0, k, 4096):                                              # exact reference in chunks that fit u64
        ref = (ref + (((a[:, c:c + 4096] % (1 << 32)).astype(np.uint64) @ b[c:c + 4096].astype(np.uint64)) % (1 << 32))) % (1 << 32)
    got = limb_gemm(a, b)
    assert np.array_equal(got, ref.astype(np.uint32))
    if k >= 40000:                                                           # the case the test is for: an accumulator did wrap
        assert int((a[0] & 255) @ (b[:, 0] & 255)) >= 2**31
