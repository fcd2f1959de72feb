// ORACLE — TEST INFRASTRUCTURE ONLY (see spiral_oracle.hpp).
// CPU restatement of DoublePIR's packed matvec, lib/doublepir/src/matrix/kernels.rs.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cstring>

namespace dpir {

// kernels.rs:9-12
static const unsigned COMPRESSION = 3;
static const uint32_t BASIS = 10;
static const uint32_t MASK = (1u << BASIS) - 1;

// kernels.rs:14-113.  out[i] += sum_k sum_{m<3} ((a[i][k] >> 10m) & 1023) * b[3k+m]  (wrapping u32).
// The reference unrolls 8 rows; the arithmetic per row is identical.
inline void raw_mat_mul_vec_packed(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t a_rows, size_t a_cols) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < a_rows; i++) {
    uint32_t tmp = 0;
    const uint32_t* row = a + i * a_cols;
    for (size_t k = 0; k < a_cols; k++) {
      uint32_t db = row[k];
      tmp += (db & MASK) * b[3 * k];
      tmp += ((db >> BASIS) & MASK) * b[3 * k + 1];
      tmp += ((db >> (2 * BASIS)) & MASK) * b[3 * k + 2];
    }
    out[i] += tmp;
  }
}

// kernels.rs:118-178: output allocated rows+8, tail rows zero-padded to a block of 8, truncated.
inline void matrix_mul_vec_packed(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t rows, size_t cols) {
  std::vector<uint32_t> o(rows + 8, 0);
  size_t down = (rows / 8) * 8;
  raw_mat_mul_vec_packed(o.data(), a, b, down, cols);
  if (down < rows) {
    size_t diff = rows - down;
    std::vector<uint32_t> tmp(8 * cols, 0);
    std::memcpy(tmp.data(), a + down * cols, diff * cols * 4);
    raw_mat_mul_vec_packed(o.data() + down, tmp.data(), b, 8, cols);
  }
  std::memcpy(out, o.data(), rows * 4);
}

// kernels.rs:180-278 matrix_mul_transposed_packed: out[i][j] = sum_k sum_m ((a[i][k] >> 10m) & 1023) * b[j][3k+m]
// (both loop orders of the reference compute this; wrapping u32)
inline void matrix_mul_transposed_packed(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t a_rows, size_t a_cols,
                                         size_t b_rows, size_t b_cols) {
  for (size_t i = 0; i < a_rows; i++)
    for (size_t j = 0; j < b_rows; j++) {
      uint32_t tmp = 0;
      for (size_t k = 0; k < a_cols; k++) {
        uint32_t db = a[i * a_cols + k];
        for (unsigned m = 0; m < COMPRESSION; m++) tmp += ((db >> (m * BASIS)) & MASK) * b[j * b_cols + k * COMPRESSION + m];
      }
      out[i * b_rows + j] = tmp;
    }
}

// matrix/indexing.rs:117-143 transpose_expand_concat_cols_squish.  out: (cols*delta*concat) x ceil((rows/concat)/d)
inline void transpose_expand_concat_cols_squish(std::vector<uint32_t>& out, size_t& out_rows, size_t& out_cols,
                                                const uint32_t* a, size_t rows, size_t cols, uint64_t modulus, size_t delta,
                                                size_t concat, uint64_t basis, size_t d) {
  out_rows = cols * delta * concat;
  out_cols = (rows / concat + d - 1) / d;
  out.assign(out_rows * out_cols, 0);
  for (size_t j = 0; j < rows; j++)
    for (size_t i = 0; i < cols; i++) {
      uint64_t val = a[i + j * cols];
      for (size_t f = 0; f < delta; f++) {
        uint64_t new_val = val % modulus;
        size_t r = (i * delta + f) + cols * delta * (j % concat);
        size_t c = j / concat;
        out[r * out_cols + c / d] += (uint32_t)(new_val << (basis * (c % d)));
        val /= modulus;
      }
    }
}

// ---- offline setup (doublepir.rs:76-108) and the matrix operations it is made of -----------------------------------------
struct Mat {                                  // matrix/matrix.rs: row-major u32
  size_t rows = 0, cols = 0;
  std::vector<uint32_t> data;
  Mat() {}
  Mat(size_t r, size_t c) : rows(r), cols(c), data(r * c, 0) {}
};
// matrix/ops.rs:169-191 raw_mat_mul_add / Mul for &Matrix: wrapping u32
inline Mat mul(const Mat& a, const Mat& b) {
  Mat c(a.rows, b.cols);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < a.rows; i++)
    for (size_t k = 0; k < a.cols; k++) {
      const uint32_t av = a.data[a.cols * i + k];
      for (size_t j = 0; j < b.cols; j++) c.data[b.cols * i + j] += av * b.data[b.cols * k + j];
    }
  return c;
}
inline Mat transpose(const Mat& a) {          // matrix/transpose.rs:9-20
  Mat o(a.cols, a.rows);
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < a.cols; j++) o.data[j * o.cols + i] = a.data[i * a.cols + j];
  return o;
}
// matrix/contract.rs:62-78 expand: every value -> delta digits mod `modulus`, each mapped to [-mod/2, mod/2) (raw_to_centered,
// arith.rs:30-32: wrapping subtraction), digit f of row i in row i*delta + f
inline Mat expand(const Mat& a, uint32_t modulus, size_t delta) {
  Mat o(a.rows * delta, a.cols);
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < a.cols; j++) {
      uint32_t val = a.data[i * a.cols + j];
      for (size_t f = 0; f < delta; f++) {
        o.data[(i * delta + f) * a.cols + j] = (val % modulus) - modulus / 2;
        val /= modulus;
      }
    }
  return o;
}
inline Mat concat_cols(const Mat& a, size_t n) {   // matrix/indexing.rs:82-101
  if (n == 1) return a;
  Mat o(a.rows * n, a.cols / n);
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < a.cols; j++) o.data[(i + a.rows * (j % n)) * o.cols + j / n] = a.data[i * a.cols + j];
  return o;
}
inline Mat squish(const Mat& a, uint64_t basis, size_t delta) {   // matrix/squish.rs:52-70
  Mat o(a.rows, (a.cols + delta - 1) / delta);
  for (size_t i = 0; i < o.rows; i++)
    for (size_t j = 0; j < o.cols; j++)
      for (size_t k = 0; k < delta; k++)
        if (delta * j + k < a.cols) o.data[i * o.cols + j] += a.data[i * a.cols + delta * j + k] << (k * basis);
  return o;
}
struct SetupOut { Mat db_squished, h1_squished, a2_t, h2; };
// doublepir.rs:76-108.  db: l x m, entries centered in [-p/2, p/2) (wrapping u32); a1: m x n; a2: (l / x) x n;
// delta = digits of a Z_q value in base p (params.delta()), x = info.x.
inline SetupOut setup(const Mat& db, const Mat& a1, const Mat& a2, uint32_t p, size_t delta, size_t x) {
  SetupOut o;
  Mat h1 = mul(db, a1);                        // (l, m) * (m, n) = (l, n)
  h1 = transpose(h1);                          // (n, l)
  h1 = expand(h1, p, delta);
  h1 = concat_cols(h1, x);                     // (n * delta * x, l / x)
  o.h2 = mul(h1, a2);
  Mat d = db;
  for (auto& v : d.data) v += p / 2;           // db.data += p / 2; db.squish()
  o.db_squished = squish(d, BASIS, COMPRESSION);
  for (auto& v : h1.data) v += p / 2;
  o.h1_squished = squish(h1, BASIS, COMPRESSION);
  Mat a2c = a2;                                // a_2_copy: zero rows up to a multiple of 3, transposed
  if (a2c.rows % 3) { size_t add = 3 - a2c.rows % 3; a2c.data.resize((a2c.rows + add) * a2c.cols, 0); a2c.rows += add; }
  o.a2_t = transpose(a2c);
  return o;
}

}  // namespace dpir
