// Index / limb arithmetic of the tcgen05 first dimension (tc5_kernels.cu), as __host__ __device__ functions so that the
// operand images and the epilogue can be emulated thread by thread on the CPU (tests/cpp/tc5_emul.cpp) against the
// canonical UMMA shared-memory layout (K-major, no swizzle: cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::K>).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define TC5_HD __host__ __device__ __forceinline__
#else
#define TC5_HD inline
#endif

namespace b200pir {

struct Tc5Geom { int dim0, rows, mt /* ceil(rows/32) */, ks /* ceil(dim0/32) */; };
inline Tc5Geom make_tc5_geom(int dim0, int rows) { return Tc5Geom{dim0, rows, (rows + 31) / 32, (dim0 + 31) / 32}; }

constexpr int TC5_M = 128, TC5_N = 128, TC5_K = 32;
constexpr int TC5_TILE = TC5_M * TC5_K;                 // 4096 bytes, A and B tiles alike
constexpr int TC5_LBO = 128;                            // bytes between the two 16-byte K halves of a core-matrix row group
constexpr int TC5_SBO = 256;                            // bytes between consecutive 8-row groups

// byte (midx, k) of a 128 x 32 tile: 8-row x 16-byte core matrices
TC5_HD int tc5_tile_off(int midx, int k) { return (midx >> 3) * TC5_SBO + (k >> 4) * TC5_LBO + (midx & 7) * 16 + (k & 15); }
// GEMM indices: M = 4 * row_local + l (database row within the 32-row tile, limb), N = 4 * col + m (col = 2 query + ct row)
TC5_HD int tc5_m_index(int row_local, int l) { return 4 * row_local + l; }
TC5_HD int tc5_n_index(int col, int m) { return 4 * col + m; }
TC5_HD size_t tc5_db_tile(const Tc5Geom& T, int slice, int n, int z, int mt, int ks) {
  return ((((size_t)slice * 2 + n) * 2048 + z) * T.mt + mt) * T.ks + ks;
}
TC5_HD size_t tc5_q_tile(const Tc5Geom& T, int n, int z, int ks) { return ((size_t)n * 2048 + z) * T.ks + ks; }
// limb l (7 bits) of four residues as four bytes, lowest address first
TC5_HD uint32_t tc5_limb4(const uint32_t (&r)[4], int l) {
  const int sh = 7 * l;
  return ((r[0] >> sh) & 127u) | (((r[1] >> sh) & 127u) << 8) | (((r[2] >> sh) & 127u) << 16) | (((r[3] >> sh) & 127u) << 24);
}

// ---- descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp; cross-checked against those structs by
// tests/cpp/tc5_desc_check.cu)
// shared-memory matrix descriptor: start address, LBO, SBO in 16-byte units, version 1 (Blackwell), SWIZZLE_NONE (0)
TC5_HD uint64_t tc5_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);           // [0,14)   start address
  d |= (uint64_t)((TC5_LBO >> 4) & 0x3FFF) << 16;       // [16,30)  leading-dimension byte offset (K halves)
  d |= (uint64_t)((TC5_SBO >> 4) & 0x3FFF) << 32;       // [32,46)  stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;                               // [46,48)  descriptor version
  return d;
}
// instruction descriptor: D = s32, A = B = unsigned 8 bit, both K-major, dense, no saturation, M = 128, N = 128
TC5_HD uint32_t tc5_instr_desc() {
  return (2u << 4) /* c_format S32 */ | (0u << 7) /* a u8 */ | (0u << 10) /* b u8 */ | (0u << 15) | (0u << 16) |
         ((uint32_t)(TC5_N >> 3) << 17) | ((uint32_t)(TC5_M >> 4) << 24);
}

// ---- database image: thread tid of the CTA of row tile mt / k-step ks handles one row and four consecutive j
struct Tc5DbThread { int row_local, kq, ii, jp0; };
TC5_HD Tc5DbThread tc5_db_thread(int tid, int mt, int ks) {
  Tc5DbThread t;
  t.row_local = tid >> 3;
  t.kq = tid & 7;
  t.ii = mt * 32 + t.row_local;
  t.jp0 = (ks * 32 + 4 * t.kq) >> 1;                    // the thread reads cells jp0 and jp0 + 1 (two j each)
  return t;
}
// res[i] = residue (one modulus) of j = ks*32 + 4 kq + i; writes the thread's four 4-byte words of one tile
TC5_HD void tc5_db_store(uint8_t* tile, const Tc5DbThread& t, const uint32_t (&res)[4]) {
  for (int l = 0; l < 4; l++) {
    const uint32_t v = tc5_limb4(res, l);
    uint8_t* p = tile + tc5_tile_off(tc5_m_index(t.row_local, l), 4 * t.kq);
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
  }
}

// ---- query image: cell = (query, k, z parity) of the CTA of k-step ks
struct Tc5QueryCell { int zp, k, q; };
TC5_HD Tc5QueryCell tc5_query_cell(int cell) { return Tc5QueryCell{cell & 1, (cell >> 1) & 31, cell >> 6}; }
// value = residue (one modulus) of query q, ciphertext row r at (j = ks*32 + k): its four limb bytes
TC5_HD void tc5_query_store(uint8_t* tile, int q, int r, int k, uint32_t value) {
  for (int m = 0; m < 4; m++) tile[tc5_tile_off(tc5_n_index(2 * q + r, m), k)] = (uint8_t)((value >> (7 * m)) & 127u);
}

// ---- epilogue: TMEM lane = M index; a warp owns the quadrant `quad` (32 lanes), a chunk is 32 consecutive columns
// (8 GEMM columns x 4 query limbs m).  Lane (row, l) holds D[4 row + l][4 col + m] = sum_j a_l(row, j) b_m(j, col) < 2^24.
// The residue is  sum_{l,m} D_{l,m} 2^{7(l+m)} mod q:
//   1. per lane and column: s = sum_m D_{l,m} 2^{7m} (< 2^47, shifts and adds), split at bit 23 and weighted with
//      w_l = 2^{7l} mod q, w'_l = 2^{7l+23} mod q:  P = (s mod 2^23) w_l + (s >> 23) w'_l  < 2^53   (two wide multiply-adds)
//   2. reduce-scatter over the four limb lanes of a row (xor 2, then xor 1): lane l ends with the sums (< 2^55) of the two
//      columns 2l, 2l+1 it stores = both ciphertext rows of query 4 chunk + l
//   3. one Barrett reduction per stored word, with 32-bit operations (tc5_barrett57).
TC5_HD int tc5_lane_row(int quad, int lane) { return quad * 8 + (lane >> 2); }     // row_local of this lane
TC5_HD int tc5_lane_limb(int lane) { return lane & 3; }
TC5_HD int tc5_lane_query(int chunk, int lane) { return chunk * 4 + (lane & 3); }  // the query this lane stores
struct Tc5Weights { uint32_t w, wp, mu; };
TC5_HD Tc5Weights tc5_lane_weights(int l, uint32_t q) {
  Tc5Weights W;
  W.w = (uint32_t)((1ull << (7 * l)) % q);
  W.wp = (uint32_t)((1ull << (7 * l + 23)) % q);
  W.mu = (uint32_t)((1ull << 58) / q);                                             // q > 2^27, so mu < 2^31
  return W;
}
// x mod q for x < 2^57 and 2^27 < q < 2^28: quotient estimate floor((x >> 26) mu / 2^32) is the true quotient or one less
// (x/2^58 + 2^26/q < 1), so the remainder estimate lies in [0, 2q) and its low 32 bits suffice.
TC5_HD uint32_t tc5_barrett57(uint64_t x, uint32_t mu, uint32_t q) {
  const uint32_t a = (uint32_t)(x >> 26);
#if defined(__CUDA_ARCH__)
  const uint32_t qh = __umulhi(a, mu);
#else
  const uint32_t qh = (uint32_t)(((uint64_t)a * mu) >> 32);
#endif
  const uint32_t r = (uint32_t)x - qh * q;
  const uint32_t r2 = r - q;
  return r < r2 ? r : r2;
}
// step 1 for one GEMM column: v4 = the four m-limb accumulators of this lane
TC5_HD uint64_t tc5_lane_partial(const uint32_t* v4, uint32_t w, uint32_t wp) {
  const uint32_t a = v4[0] + (v4[1] << 7), b = v4[2] + (v4[3] << 7);               // < 2^32 each (D < 2^24)
  const uint64_t s = (uint64_t)a + ((uint64_t)b << 14);                            // < 2^47
  return (uint64_t)((uint32_t)s & 0x7FFFFFu) * w + (uint64_t)(uint32_t)(s >> 23) * wp;
}
// step 2, first exchange (partner = lane ^ 2): lanes with limb bit 1 clear keep columns 0..3 and send 4..7, the others the
// opposite; second exchange (partner = lane ^ 1): of the four kept columns, lanes with limb bit 0 clear keep the first two.
TC5_HD void tc5_rs_select_a(int l, const uint64_t (&P)[8], uint64_t (&send)[4], uint64_t (&keep)[4]) {
  const bool up = (l & 2) != 0;
  for (int i = 0; i < 4; i++) { send[i] = up ? P[i] : P[4 + i]; keep[i] = up ? P[4 + i] : P[i]; }
}
TC5_HD void tc5_rs_select_b(int l, const uint64_t (&K)[4], uint64_t (&send)[2], uint64_t (&keep)[2]) {
  const bool odd = (l & 1) != 0;
  for (int i = 0; i < 2; i++) { send[i] = odd ? K[i] : K[2 + i]; keep[i] = odd ? K[2 + i] : K[i]; }
}

}  // namespace b200pir
