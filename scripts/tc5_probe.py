"""Bring-up probe for the tcgen05 first dimension (database format 2).  Run on a B200 UNDER A TIMEOUT:

    timeout 120 python scripts/tc5_probe.py

It pushes a small deterministic database / query through b200pir_multiply_reg_by_database with db_format 2, asks the
launcher to dump the raw s32 accumulators of the first tiles (B200PIR_TC5_DUMP), and compares them with D[M][N] computed
in numpy from the operand values under (a) the assumed layout and (b) the usual suspects (LBO/SBO swapped, operands
transposed, TMEM rows permuted), so that a wrong hardware assumption is identified in one run.  Uses only the product
library and numpy: the reference values are plain integer sums."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

Q0, Q1 = 268369921, 249561089
N = 2048


def tile_off(midx, k, lbo=128, sbo=256):
    return (midx >> 3) * sbo + (k >> 4) * lbo + (midx & 7) * 16 + (k & 15)


def main():
    dump = os.path.join(tempfile.gettempdir(), "tc5_dump.bin")
    os.environ["B200PIR_TC5_DUMP"] = dump
    import sdk_b200.spiral as S
    kw = dict(n=1, nu_1=6, nu_2=5, p=256, q2_bits=20, t_gsw=8, t_conv=4, t_exp_left=8, t_exp_right=8, instances=1,
              db_item_size=2048, version=0)                          # dim0 = 64 (2 k-steps), 32 rows (1 row tile), 1 slice
    G = S.Params(**kw)
    dim0, rows = 64, 32
    rng = np.random.default_rng(5)
    a0 = rng.integers(0, Q0, (N, rows, dim0), dtype=np.uint64)       # [z][ii][j]
    a1 = rng.integers(0, Q1, (N, rows, dim0), dtype=np.uint64)
    db = (a0 | (a1 << np.uint64(32))).reshape(-1)
    b0 = rng.integers(0, Q0, (N, dim0, 2), dtype=np.uint64)          # [z][j][r]
    b1 = rng.integers(0, Q1, (N, dim0, 2), dtype=np.uint64)
    v = (b0 | (b1 << np.uint64(32))).reshape(-1)
    tdb = S.Database.from_words(G, db, fmt=2)
    got = S.multiply_reg_by_database(G, tdb, 0, v)                    # rows x [2][2][2048]
    # reference product, modulus 0, z = 0
    ref = (a0[0].astype(object) @ b0[0].astype(object)) % Q0          # [ii][r]
    out = got.reshape(rows, 2, 2, N)
    ok_final = all(int(out[ii, r, 0, 0]) == int(ref[ii, r]) for ii in range(rows) for r in range(2))
    print("final residues (n = 0, z = 0):", "MATCH" if ok_final else "MISMATCH")
    D = np.fromfile(dump, dtype=np.int32).reshape(-1, 128, 128)[0]
    # limb matrices of the first tile: A[M = 4 ii + l][j], B[N = 4 c + m][j], c = 2 query + r (one query here)
    A = np.zeros((128, dim0), dtype=np.int64)
    B = np.zeros((128, dim0), dtype=np.int64)
    for ii in range(rows):
        for l in range(4):
            A[4 * ii + l] = (a0[0, ii].astype(np.int64) >> (7 * l)) & 127
    for r in range(2):
        for m in range(4):
            B[4 * r + m] = (b0[0, :, r].astype(np.int64) >> (7 * m)) & 127
    want = A @ B.T
    if np.array_equal(D, want):
        print("raw accumulators: MATCH the assumed layout (M = 4 row + limb in TMEM lanes, N = 4 col + limb in columns)")
        return 0 if ok_final else 2
    print("raw accumulators: MISMATCH; D[0, :8] =", D[0, :8], "want", want[0, :8])
    hyp = {"transposed (D = B A^T)": want.T}
    # operand images re-read with LBO and SBO exchanged
    def reread(Mtx, lbo, sbo):
        img = np.zeros((dim0 // 32, 4096), dtype=np.int64)
        for ks in range(dim0 // 32):
            for mi in range(128):
                for k in range(32):
                    img[ks, tile_off(mi, k)] = Mtx[mi, ks * 32 + k]
        out = np.zeros_like(Mtx)
        for ks in range(dim0 // 32):
            for mi in range(128):
                for k in range(32):
                    out[mi, ks * 32 + k] = img[ks, tile_off(mi, k, lbo, sbo) % 4096]
        return out
    As, Bs = reread(A, 256, 128), reread(B, 256, 128)
    hyp["LBO/SBO exchanged"] = As @ Bs.T
    hyp["rows 0..63 only (M = 64 semantics)"] = np.vstack([want[:64], np.zeros((64, 128), dtype=np.int64)])
    for name, h in hyp.items():
        print("  hypothesis %-40s %s" % (name, "MATCH" if np.array_equal(D, h) else "no"))
    nz = np.argwhere(D != want)
    print("  first differing (M, N):", nz[:5].tolist(), " count", len(nz))
    return 1


if __name__ == "__main__":
    sys.exit(main())
