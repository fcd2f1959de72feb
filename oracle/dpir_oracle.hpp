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

}  // namespace dpir
