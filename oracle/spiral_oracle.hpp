// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, no dependencies) of the Spiral server-side query path of
// blyssprivacy/sdk (reference @ fdb7206), used ONLY as the parity checker for the
// CUDA kernels in sdk_b200/csrc and as the CPU baseline leg of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
//
// Parity pinning: every arithmetic primitive is pinned against the reference's own
// known-answer tests (tests/test_oracle_kats.py):
//   * build_ntt_tables XOR = 519370102, inv table [0]=134184961, [1]=96647580  (ntt.rs:377-398)
//   * ntt_forward(delta*100) -> all 100 ; ntt_inverse(all 100) -> delta*100      (ntt.rs:400-423)
//   * Barrett constants of q0, q1, q                                              (arith.rs:477-490)
//   * barrett_reduction_u128_raw vectors, div2_uint_mod(3,7)=5, calc_index       (arith.rs:457-501, ntt.rs:445-449)
//   * negacyclic 100X * 7X = 700X^2 (poly.rs:731-743), gadget digits of 3 and 6  (gadget.rs:78-95)
// The pipeline stages (expansion, multiply, fold, pack, encode) have NO stored golden
// ciphertexts in the reference (its tests decrypt and compare, with fresh entropy), so the
// pipeline is pinned by (i) the primitive KATs, (ii) line-by-line structural correspondence
// (each function cites the reference lines it follows), (iii) the same decrypt-and-compare
// tests.  Wire-level seed expansion (ChaCha20, rand_chacha 0.3.1: a third-party crate absent from the
// reference tree) is pinned by RFC 8439 / RFC 7539 vectors and the crate's published ChaCha20Rng
// known-answer test instead of reference output (spiral_client.hpp).
//
// The Rust reference cannot be compiled here (no cargo/rustc in the image).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <array>
#include <stdexcept>
#include <algorithm>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace orc {

// set by the CPU-baseline legs of bench.py: use the AVX2 first-dimension kernel inside process_query
static bool g_use_avx2_multiply = false;
static bool g_sparse_fold = false;        // process_query folds like lib/server (compute/fold.rs:15-65) instead of spiral-rs (server.rs:388-427)


typedef unsigned __int128 u128;
typedef __int128 i128;
typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

// ---------------------------------------------------------------- arith.rs
// arith.rs:5-7
inline u64 multiply_uint_mod(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }
// arith.rs:9-11
inline u64 log2_floor(u64 a) { return 63 - __builtin_clzll(a); }
// arith.rs:13-19 (f64 ceil(log2)); exact for the integer ranges used here
// (Rust's `f64 as usize` saturates: log2(0) = -inf -> 0)
inline u64 log2_ceil(u64 a) { return a == 0 ? 0 : (u64)std::ceil(std::log2((double)a)); }
// arith.rs:41-67
inline u64 exponentiate_uint_mod(u64 operand, u64 exponent, u64 m) {
  u64 result = 1 % m, base = operand % m;
  while (exponent) {
    if (exponent & 1) result = multiply_uint_mod(result, base, m);
    base = multiply_uint_mod(base, base, m);
    exponent >>= 1;
  }
  return result;
}
// arith.rs:69-76
inline u64 reverse_bits(u64 x, unsigned bit_count) {
  if (bit_count == 0) return 0;
  u64 r = 0;
  for (unsigned i = 0; i < bit_count; i++) r |= ((x >> i) & 1) << (bit_count - 1 - i);
  return r;
}
// arith.rs:78-89
inline u64 div2_uint_mod(u64 operand, u64 m) {
  if (operand & 1) {
    u128 s = (u128)operand + m;
    return (u64)(s >> 1);
  }
  return operand >> 1;
}
// arith.rs:91-104
inline u64 recenter(u64 val, u64 from_modulus, u64 to_modulus) {
  i64 from = (i64)from_modulus, to = (i64)to_modulus;
  i64 a = (i64)val;
  if (val >= from_modulus / 2) a -= from;
  a = a + (from / to) * to + 2 * to;
  a %= to;
  return (u64)a;
}
// arith.rs:106-111, 335-413: floor(2^128 / modulus) as (lo, hi) 64-bit words.
// Long division, two 64-bit digits.
inline void get_barrett_crs(u64 modulus, u64& cr0, u64& cr1) {
  // 2^128 / m : first digit q_hi = floor(2^64 / m), remainder r; then (r<<64)/m.
  u128 two64 = (u128)1 << 64;
  u128 q_hi = two64 / modulus;      // fits in 64 bits for m>1
  u128 r = two64 % modulus;
  u128 num = r << 64;               // r < m < 2^64
  u128 q_lo = num / modulus;
  // 2^128 = (q_hi*2^64 + q_lo)*m + rem
  cr1 = (u64)q_hi;
  cr0 = (u64)q_lo;
}
// arith.rs:122-134
inline u64 barrett_raw_u64(u64 input, u64 const_ratio_1, u64 modulus) {
  u64 tmp = (u64)(((u128)input * const_ratio_1) >> 64);
  u64 res = input - tmp * modulus;
  return res >= modulus ? res - modulus : res;
}
// arith.rs:165-202 (literal restatement of the SEAL-style 128-bit Barrett)
inline u64 barrett_raw_u128(u128 val, u64 cr0, u64 cr1, u64 modulus) {
  u64 zx = (u64)val, zy = (u64)(val >> 64);
  u64 tmp1 = 0, tmp3, carry;
  u64 prody = (u64)(((u128)zx * cr0) >> 64);
  carry = prody;
  u128 t2 = (u128)zx * cr1;
  u64 tmp2x = (u64)t2, tmp2y = (u64)(t2 >> 64);
  // add_u64 (arith.rs:155-163): returns 1 on overflow and leaves *out untouched
  auto add_u64 = [](u64 a, u64 b, u64* out) -> u64 {
    u64 s;
    if (__builtin_add_overflow(a, b, &s)) return 1;
    *out = s;
    return 0;
  };
  tmp3 = tmp2y + add_u64(tmp2x, carry, &tmp1);
  t2 = (u128)zy * cr0;
  tmp2x = (u64)t2; tmp2y = (u64)(t2 >> 64);
  carry = tmp2y + add_u64(tmp1, tmp2x, &tmp1);
  tmp1 = zy * cr1 + tmp3 + carry;
  tmp3 = zx - tmp1 * modulus;
  return tmp3;
}
inline u64 barrett_reduction_u128_raw(u64 modulus, u64 cr0, u64 cr1, u128 val) {
  u64 r = barrett_raw_u128(val, cr0, cr1, modulus);
  r -= modulus * (u64)(r >= modulus);
  return r;
}
// arith.rs:415-427
inline u64 recenter_mod(u64 val, u64 small_modulus, u64 large_modulus) {
  i64 v = (i64)val;
  if (v > (i64)small_modulus / 2) v -= (i64)small_modulus;
  if (v < 0) v += (i64)large_modulus;
  return (u64)v;
}
// arith.rs:429-444
inline u64 rescale(u64 a, u64 inp_mod, u64 out_mod) {
  i64 inp_mod_i = (i64)inp_mod;
  i128 out_mod_i = (i128)out_mod;
  i64 inp_val = (i64)(a % inp_mod);
  if (inp_val >= inp_mod_i / 2) inp_val -= inp_mod_i;
  i64 sign = inp_val >= 0 ? 1 : -1;
  i128 val = (i128)inp_val * (i128)out_mod;
  i128 result = (val + (i128)(sign * (inp_mod_i / 2))) / (i128)inp_mod;   // truncating division
  result = (result + (i128)((inp_mod / out_mod) * out_mod) + 2 * out_mod_i) % out_mod_i;
  return (u64)((result + out_mod_i) % out_mod_i);
}

// ---------------------------------------------------------------- number_theory.rs
// number_theory.rs:58-96
inline bool invert_uint_mod(u64 value, u64 modulus, u64& out) {
  if (value == 0) return false;
  i128 r0 = value, r1 = modulus, s0 = 1, s1 = 0;
  while (r1 != 0) {
    i128 q = r0 / r1;
    i128 t = r0 - q * r1; r0 = r1; r1 = t;
    t = s0 - q * s1; s0 = s1; s1 = t;
  }
  if (r0 != 1) return false;
  i128 m = modulus;
  s0 %= m; if (s0 < 0) s0 += m;
  out = (u64)s0;
  return true;
}
// number_theory.rs:6-12
inline bool is_primitive_root(u64 root, u64 degree, u64 modulus) {
  if (root == 0) return false;
  return exponentiate_uint_mod(root, degree >> 1, modulus) == modulus - 1;
}
// number_theory.rs:14-55.  The reference draws random candidates and then takes the minimal
// root over all odd powers (:41-55), so the result is deterministic; we scan candidates.
inline bool get_minimal_primitive_root(u64 degree, u64 modulus, u64& out) {
  u64 group = modulus - 1;
  u64 quot = group / degree;
  if (group - quot * degree != 0) return false;
  u64 root = 0;
  bool found = false;
  for (u64 cand = 2; cand < 2000; cand++) {
    root = exponentiate_uint_mod(cand, quot, modulus);
    if (is_primitive_root(root, degree, modulus)) { found = true; break; }
  }
  if (!found) return false;
  u64 gsq = multiply_uint_mod(root, root, modulus);
  u64 cur = root;
  for (u64 i = 0; i < degree; i++) {
    if (cur < root) root = cur;
    cur = multiply_uint_mod(cur, gsq, modulus);
  }
  out = root;
  return true;
}

// ---------------------------------------------------------------- params.rs
static const u64 Q2_VALUES[37] = {   // params.rs:8-46
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    12289ULL, 12289ULL, 61441ULL, 65537ULL, 65537ULL, 520193ULL, 786433ULL, 786433ULL,
    3604481ULL, 7340033ULL, 16515073ULL, 33292289ULL, 67043329ULL, 132120577ULL,
    268369921ULL, 469762049ULL, 1073479681ULL, 2013265921ULL, 4293918721ULL,
    8588886017ULL, 17175674881ULL, 34359214081ULL, 68718428161ULL};

struct Params {
  size_t poly_len = 0, poly_len_log2 = 0;
  // ntt_tables[mod][0..3] = fwd, fwd', inv, inv'   (ntt.rs:39-65)
  std::vector<std::array<std::vector<u64>, 4>> ntt_tables;
  size_t crt_count = 0;
  u64 barrett_cr_0[4] = {0, 0, 0, 0}, barrett_cr_1[4] = {0, 0, 0, 0};
  u64 barrett_cr_0_modulus = 0, barrett_cr_1_modulus = 0;
  u64 mod0_inv_mod1 = 0, mod1_inv_mod0 = 0;
  u64 moduli[4] = {0, 0, 0, 0};
  u64 modulus = 0, modulus_log2 = 0;
  double noise_width = 6.4;
  size_t n = 0; u64 pt_modulus = 0, q2_bits = 0;
  size_t t_conv = 0, t_exp_left = 0, t_exp_right = 0, t_gsw = 0;
  bool expand_queries = true;
  size_t db_dim_1 = 0, db_dim_2 = 0, instances = 1, db_item_size = 0, version = 0;

  size_t num_expanded() const { return (size_t)1 << db_dim_1; }
  size_t num_items() const { return ((size_t)1 << db_dim_1) * ((size_t)1 << db_dim_2); }
  size_t g() const { return (size_t)log2_ceil(t_gsw * db_dim_2 + num_expanded()); }          // params.rs:129-132
  size_t stop_round() const { return (size_t)log2_ceil(t_gsw * db_dim_2); }                 // params.rs:134-136
  size_t bytes_per_chunk() const {                                                          // params.rs:188-193
    size_t chunks = instances * n * n;
    return (db_item_size + chunks - 1) / chunks;
  }
  size_t modp_words_per_chunk() const {                                                     // params.rs:195-200
    size_t logp = log2_floor(pt_modulus);
    return (bytes_per_chunk() * 8 + logp - 1) / logp;
  }
  size_t setup_bytes() const {                                                              // params.rs:146-167
    size_t sz_polys = 0;
    size_t num_packing = version == 0 ? n : 2;
    sz_polys += num_packing * (n * t_conv);
    if (expand_queries) {
      size_t left = g() * t_exp_left;
      size_t right = (stop_round() + 1) * t_exp_right;
      if (version > 0 && t_exp_left == t_exp_right) right = 0;
      sz_polys += left + right + 2 * t_conv;
    }
    return 32 + sz_polys * poly_len * 8;
  }
  size_t query_bytes() const {                                                              // params.rs:169-182
    size_t sz_polys = expand_queries ? 1 : num_expanded() + db_dim_2 * (2 * t_gsw);
    return 32 + sz_polys * poly_len * 8;
  }
  u64 crt_compose_2(u64 x, u64 y) const {                                                   // params.rs:207-214
    u128 val = (u128)x * mod1_inv_mod0 + (u128)y * mod0_inv_mod1;
    return barrett_reduction_u128_raw(modulus, barrett_cr_0_modulus, barrett_cr_1_modulus, val);
  }
  u64 crt_compose(const u64* a, size_t idx) const {                                         // params.rs:216-222
    return crt_count == 1 ? a[idx] : crt_compose_2(a[idx], a[idx + poly_len]);
  }
  u64 barrett_coeff(u64 val, size_t nn) const { return barrett_raw_u64(val, barrett_cr_1[nn], moduli[nn]); }  // arith.rs:140-142
};

// ntt.rs:6-65
inline void build_ntt_tables(Params& p) {
  size_t N = p.poly_len, lg = p.poly_len_log2;
  p.ntt_tables.resize(p.crt_count);
  for (size_t c = 0; c < p.crt_count; c++) {
    u64 m = p.moduli[c];
    u64 root = 0, inv_root = 0;
    if (!get_minimal_primitive_root(2 * N, m, root)) throw std::runtime_error("no primitive root");
    if (!invert_uint_mod(root, m, inv_root)) throw std::runtime_error("no inverse root");
    auto powers = [&](u64 r) {
      std::vector<u64> v(N, 0);
      u64 power = r;
      for (size_t i = 1; i < N; i++) {
        v[reverse_bits(i, lg)] = power;
        power = multiply_uint_mod(power, r, m);
      }
      v[0] = 1;
      return v;
    };
    auto scale32 = [&](const std::vector<u64>& in) {   // ntt.rs:29-37
      std::vector<u64> v(N);
      for (size_t i = 0; i < N; i++) v[i] = (u64)(u32)((in[i] << 32) / (u64)(u32)m);
      return v;
    };
    std::vector<u64> f = powers(root);
    std::vector<u64> fi = powers(inv_root);
    for (size_t i = 0; i < N; i++) fi[i] = div2_uint_mod(fi[i], m);
    p.ntt_tables[c][0] = f;
    p.ntt_tables[c][1] = scale32(f);
    p.ntt_tables[c][2] = fi;
    p.ntt_tables[c][3] = scale32(fi);
  }
}

// params.rs:224-296
inline Params params_init(size_t poly_len, const std::vector<u64>& moduli, double noise_width, size_t n,
                          u64 pt_modulus, u64 q2_bits, size_t t_conv, size_t t_exp_left, size_t t_exp_right,
                          size_t t_gsw, bool expand_queries, size_t db_dim_1, size_t db_dim_2, size_t instances,
                          size_t db_item_size, size_t version) {
  if (q2_bits < 14 || q2_bits > 36) throw std::runtime_error("q2_bits out of range");
  Params p;
  p.poly_len = poly_len;
  p.poly_len_log2 = log2_floor(poly_len);
  p.crt_count = moduli.size();
  if (p.crt_count > 4) throw std::runtime_error("too many moduli");
  p.modulus = 1;
  for (size_t i = 0; i < moduli.size(); i++) { p.moduli[i] = moduli[i]; p.modulus *= moduli[i]; }
  build_ntt_tables(p);
  p.modulus_log2 = log2_ceil(p.modulus);
  for (size_t i = 0; i < moduli.size(); i++) get_barrett_crs(moduli[i], p.barrett_cr_0[i], p.barrett_cr_1[i]);
  get_barrett_crs(p.modulus, p.barrett_cr_0_modulus, p.barrett_cr_1_modulus);
  if (p.crt_count == 2) {
    u64 inv = 0;
    invert_uint_mod(moduli[0], moduli[1], inv); p.mod0_inv_mod1 = moduli[0] * inv;
    invert_uint_mod(moduli[1], moduli[0], inv); p.mod1_inv_mod0 = moduli[1] * inv;
  }
  p.noise_width = noise_width; p.n = n; p.pt_modulus = pt_modulus; p.q2_bits = q2_bits;
  p.t_conv = t_conv; p.t_exp_left = t_exp_left; p.t_exp_right = t_exp_right; p.t_gsw = t_gsw;
  p.expand_queries = expand_queries; p.db_dim_1 = db_dim_1; p.db_dim_2 = db_dim_2;
  p.instances = instances; p.db_item_size = db_item_size; p.version = version;
  return p;
}

// util.rs:224-263 (poly_len and moduli are hard-coded there, :246-247)
inline Params params_from_scalars(size_t n, size_t nu_1, size_t nu_2, u64 p, u64 q2_bits, size_t t_gsw,
                                  size_t t_conv, size_t t_exp_left, size_t t_exp_right, size_t instances,
                                  size_t db_item_size, size_t version, bool expand_queries) {
  if (q2_bits < 14) q2_bits = 14;
  if (instances == 0) instances = 1;
  if (db_item_size == 0) db_item_size = instances * n * n * 2048 * log2_ceil(p) / 8;
  return params_init(2048, {268369921ULL, 249561089ULL}, 6.4, n, p, q2_bits, t_conv, t_exp_left, t_exp_right,
                     t_gsw, expand_queries, nu_1, nu_2, instances, db_item_size, version);
}

// ---------------------------------------------------------------- ntt.rs (scalar = the spec)
// ntt.rs:67-113
inline void ntt_forward_scalar(const Params& p, u64* operand_overall) {
  size_t lg = p.poly_len_log2, n = (size_t)1 << lg;
  for (size_t cm = 0; cm < p.crt_count; cm++) {
    u64* op = operand_overall + cm * n;
    const u64* ft = p.ntt_tables[cm][0].data();
    const u64* ftp = p.ntt_tables[cm][1].data();
    u32 q = (u32)p.moduli[cm];
    u32 two_q = 2 * q;
    for (size_t mm = 0; mm < lg; mm++) {
      size_t m = (size_t)1 << mm, t = n >> (mm + 1);
      for (size_t i = 0; i < m; i++) {
        u64 w = ft[m + i], wp = ftp[m + i];
        u64* o = op + i * 2 * t;
        for (size_t j = 0; j < t; j++) {
          u32 x = (u32)o[j], y = (u32)o[t + j];
          u32 curr_x = x - (two_q * (u32)(x >= two_q));
          u64 q_tmp = ((u64)y * wp) >> 32;
          u64 q_new = w * (u64)y - q_tmp * (u64)q;
          o[j] = (u64)curr_x + q_new;
          o[t + j] = (u64)curr_x + ((u64)two_q - q_new);
        }
      }
    }
    for (size_t i = 0; i < n; i++) {
      op[i] -= (u64)(op[i] >= two_q) * two_q;
      op[i] -= (u64)(op[i] >= q) * q;
    }
  }
}
// ntt.rs:212-258
inline void ntt_inverse_scalar(const Params& p, u64* operand_overall) {
  size_t n = p.poly_len;
  for (size_t cm = 0; cm < p.crt_count; cm++) {
    u64* op = operand_overall + cm * n;
    const u64* it = p.ntt_tables[cm][2].data();
    const u64* itp = p.ntt_tables[cm][3].data();
    u64 q = p.moduli[cm], two_q = 2 * q;
    for (size_t mm = p.poly_len_log2; mm-- > 0;) {
      size_t h = (size_t)1 << mm, t = n >> (mm + 1);
      for (size_t i = 0; i < h; i++) {
        u64 w = it[h + i], wp = itp[h + i];
        u64* o = op + i * 2 * t;
        for (size_t j = 0; j < t; j++) {
          u64 x = o[j], y = o[t + j];
          u64 t_tmp = two_q - y + x;
          u64 curr_x = x + y - (two_q * (u64)((x << 1) >= t_tmp));
          u64 h_tmp = (t_tmp * wp) >> 32;
          u64 res_x = (curr_x + (q * (t_tmp & 1))) >> 1;
          u64 res_y = w * t_tmp - h_tmp * q;
          o[j] = res_x;
          o[t + j] = res_y;
        }
      }
    }
    for (size_t i = 0; i < n; i++) {
      op[i] -= (u64)(op[i] >= two_q) * two_q;
      op[i] -= (u64)(op[i] >= q) * q;
    }
  }
}

#if defined(__AVX2__)
// The reference's AVX2 transforms (ntt.rs:120-210 forward, :260-345 inverse) keep four u64 lanes with 32-bit content
// and use _mm256_mul_epu32 for the 32x32->64 products.  Restated here for the CPU baseline with ONE deliberate
// difference: the reference's vector code compares with `>` where its scalar code (and this oracle) compare with
// `>=`, so it can leave the non-canonical representatives q / 2q where the scalar code produces 0; these versions
// use `>=` (cmpgt against bound - 1) and are therefore bit-identical to ntt_forward_scalar / ntt_inverse_scalar
// (asserted in tests/test_oracle_kats.py).
inline __m256i avx2_sub_if_ge(__m256i x, __m256i bound, __m256i bound_m1) {
  return _mm256_sub_epi64(x, _mm256_and_si256(_mm256_cmpgt_epi64(x, bound_m1), bound));
}
inline void ntt_forward_avx2(const Params& p, u64* operand_overall) {
  size_t lg = p.poly_len_log2, n = (size_t)1 << lg;
  for (size_t cm = 0; cm < p.crt_count; cm++) {
    u64* op = operand_overall + cm * n;
    const u64* ft = p.ntt_tables[cm][0].data();
    const u64* ftp = p.ntt_tables[cm][1].data();
    const u32 q = (u32)p.moduli[cm], two_q = 2 * q;
    const __m256i vq = _mm256_set1_epi64x(q), v2q = _mm256_set1_epi64x(two_q), v2q_m1 = _mm256_set1_epi64x((long long)two_q - 1),
                  vq_m1 = _mm256_set1_epi64x((long long)q - 1);
    for (size_t mm = 0; mm < lg; mm++) {
      size_t m = (size_t)1 << mm, t = n >> (mm + 1);
      for (size_t i = 0; i < m; i++) {
        const u64 w = ft[m + i], wp = ftp[m + i];
        u64* o = op + i * 2 * t;
        if (t < 4) {
          for (size_t j = 0; j < t; j++) {
            u32 x = (u32)o[j], y = (u32)o[t + j];
            u32 curr_x = x - (two_q * (u32)(x >= two_q));
            u64 q_tmp = ((u64)y * wp) >> 32;
            u64 q_new = w * (u64)y - q_tmp * (u64)q;
            o[j] = (u64)curr_x + q_new;
            o[t + j] = (u64)curr_x + ((u64)two_q - q_new);
          }
        } else {
          const __m256i vw = _mm256_set1_epi64x((long long)w), vwp = _mm256_set1_epi64x((long long)wp);
          for (size_t j = 0; j < t; j += 4) {
            __m256i x = _mm256_loadu_si256((const __m256i*)(o + j)), y = _mm256_loadu_si256((const __m256i*)(o + t + j));
            __m256i curr_x = avx2_sub_if_ge(x, v2q, v2q_m1);
            __m256i q_val = _mm256_srli_epi64(_mm256_mul_epu32(y, vwp), 32);
            __m256i q_fin = _mm256_sub_epi64(_mm256_mul_epu32(y, vw), _mm256_mul_epu32(q_val, vq));
            _mm256_storeu_si256((__m256i*)(o + j), _mm256_add_epi64(curr_x, q_fin));
            _mm256_storeu_si256((__m256i*)(o + t + j), _mm256_add_epi64(curr_x, _mm256_sub_epi64(v2q, q_fin)));
          }
        }
      }
    }
    for (size_t i = 0; i < n; i += 4) {
      __m256i x = _mm256_loadu_si256((const __m256i*)(op + i));
      x = avx2_sub_if_ge(x, v2q, v2q_m1);
      x = avx2_sub_if_ge(x, vq, vq_m1);
      _mm256_storeu_si256((__m256i*)(op + i), x);
    }
  }
}
inline void ntt_inverse_avx2(const Params& p, u64* operand_overall) {
  size_t n = p.poly_len;
  for (size_t cm = 0; cm < p.crt_count; cm++) {
    u64* op = operand_overall + cm * n;
    const u64* it = p.ntt_tables[cm][2].data();
    const u64* itp = p.ntt_tables[cm][3].data();
    const u64 q = p.moduli[cm], two_q = 2 * q;
    const __m256i vq = _mm256_set1_epi64x((long long)q), v2q = _mm256_set1_epi64x((long long)two_q),
                  v2q_m1 = _mm256_set1_epi64x((long long)two_q - 1), vq_m1 = _mm256_set1_epi64x((long long)q - 1),
                  one = _mm256_set1_epi64x(1);
    for (size_t mm = p.poly_len_log2; mm-- > 0;) {
      size_t h = (size_t)1 << mm, t = n >> (mm + 1);
      for (size_t i = 0; i < h; i++) {
        const u64 w = it[h + i], wp = itp[h + i];
        u64* o = op + i * 2 * t;
        if (t < 4) {
          for (size_t j = 0; j < t; j++) {
            u64 x = o[j], y = o[t + j];
            u64 t_tmp = two_q - y + x;
            u64 curr_x = x + y - (two_q * (u64)((x << 1) >= t_tmp));
            u64 h_tmp = (t_tmp * wp) >> 32;
            o[j] = (curr_x + (q * (t_tmp & 1))) >> 1;
            o[t + j] = w * t_tmp - h_tmp * q;
          }
        } else {
          const __m256i vw = _mm256_set1_epi64x((long long)w), vwp = _mm256_set1_epi64x((long long)wp);
          for (size_t j = 0; j < t; j += 4) {
            __m256i x = _mm256_loadu_si256((const __m256i*)(o + j)), y = _mm256_loadu_si256((const __m256i*)(o + t + j));
            __m256i t_tmp = _mm256_add_epi64(_mm256_sub_epi64(v2q, y), x);                 // in [1, 4q)
            __m256i sum = _mm256_add_epi64(x, y);
            // (x << 1) >= t_tmp  <=>  x + y >= 2q
            __m256i curr_x = avx2_sub_if_ge(sum, v2q, v2q_m1);
            __m256i h_tmp = _mm256_srli_epi64(_mm256_mul_epu32(t_tmp, vwp), 32);
            __m256i odd = _mm256_cmpeq_epi64(_mm256_and_si256(t_tmp, one), one);
            __m256i res_x = _mm256_srli_epi64(_mm256_add_epi64(curr_x, _mm256_and_si256(odd, vq)), 1);
            __m256i res_y = _mm256_sub_epi64(_mm256_mul_epu32(t_tmp, vw), _mm256_mul_epu32(h_tmp, vq));
            _mm256_storeu_si256((__m256i*)(o + j), res_x);
            _mm256_storeu_si256((__m256i*)(o + t + j), res_y);
          }
        }
      }
    }
    for (size_t i = 0; i < n; i += 4) {
      __m256i x = _mm256_loadu_si256((const __m256i*)(op + i));
      x = avx2_sub_if_ge(x, v2q, v2q_m1);
      x = avx2_sub_if_ge(x, vq, vq_m1);
      _mm256_storeu_si256((__m256i*)(op + i), x);
    }
  }
}
#endif
static bool g_use_avx2_ntt = false;     // CPU-baseline switch (oracle_capi.cpp orc_use_avx2_ntt); parity tests use the scalar path
inline void ntt_forward(const Params& p, u64* operand_overall) {
#if defined(__AVX2__)
  if (g_use_avx2_ntt) { ntt_forward_avx2(p, operand_overall); return; }
#endif
  ntt_forward_scalar(p, operand_overall);
}
inline void ntt_inverse(const Params& p, u64* operand_overall) {
#if defined(__AVX2__)
  if (g_use_avx2_ntt) { ntt_inverse_avx2(p, operand_overall); return; }
#endif
  ntt_inverse_scalar(p, operand_overall);
}

// ---------------------------------------------------------------- poly.rs
struct PolyMatrix {          // poly.rs:59-71 (Raw: poly_len words/poly; NTT: crt_count*poly_len)
  size_t rows = 0, cols = 0;
  bool is_ntt = false;
  size_t words = 0;          // words per poly
  std::vector<u64> data;
  PolyMatrix() {}
  PolyMatrix(const Params& p, size_t r, size_t c, bool ntt)
      : rows(r), cols(c), is_ntt(ntt), words(ntt ? p.poly_len * p.crt_count : p.poly_len), data(r * c * words, 0) {}
  u64* poly(size_t r, size_t c) { return data.data() + (r * cols + c) * words; }
  const u64* poly(size_t r, size_t c) const { return data.data() + (r * cols + c) * words; }
  // poly.rs:41-53
  void copy_into(const PolyMatrix& src, size_t tr, size_t tc) {
    if (tr + src.rows > rows || tc + src.cols > cols) throw std::runtime_error("copy_into out of range");
    for (size_t r = 0; r < src.rows; r++)
      for (size_t c = 0; c < src.cols; c++) std::memcpy(poly(tr + r, tc + c), src.poly(r, c), words * 8);
  }
  PolyMatrix submatrix(const Params& p, size_t tr, size_t tc, size_t r_, size_t c_) const {  // poly.rs:127-141
    PolyMatrix m(p, r_, c_, is_ntt);
    for (size_t r = 0; r < r_; r++)
      for (size_t c = 0; c < c_; c++) std::memcpy(m.poly(r, c), poly(tr + r, tc + c), words * 8);
    return m;
  }
  PolyMatrix pad_top(const Params& p, size_t pad) const {                                    // poly.rs:122-126
    PolyMatrix m(p, rows + pad, cols, is_ntt);
    m.copy_into(*this, pad, 0);
    return m;
  }
};
inline PolyMatrix raw_zero(const Params& p, size_t r, size_t c) { return PolyMatrix(p, r, c, false); }
inline PolyMatrix ntt_zero(const Params& p, size_t r, size_t c) { return PolyMatrix(p, r, c, true); }

// poly.rs:605-623
inline void to_ntt(const Params& p, PolyMatrix& a, const PolyMatrix& b) {
  for (size_t r = 0; r < a.rows; r++)
    for (size_t c = 0; c < a.cols; c++) {
      const u64* src = b.poly(r, c);
      u64* dst = a.poly(r, c);
      for (size_t nn = 0; nn < p.crt_count; nn++)
        for (size_t z = 0; z < p.poly_len; z++) dst[nn * p.poly_len + z] = p.barrett_coeff(src[z], nn);
      ntt_forward(p, dst);
    }
}
// poly.rs:625-638
inline void to_ntt_no_reduce(const Params& p, PolyMatrix& a, const PolyMatrix& b) {
  for (size_t r = 0; r < a.rows; r++)
    for (size_t c = 0; c < a.cols; c++) {
      const u64* src = b.poly(r, c);
      u64* dst = a.poly(r, c);
      for (size_t nn = 0; nn < p.crt_count; nn++) std::memcpy(dst + nn * p.poly_len, src, p.poly_len * 8);
      ntt_forward(p, dst);
    }
}
inline PolyMatrix to_ntt_alloc(const Params& p, const PolyMatrix& b) {
  PolyMatrix a = ntt_zero(p, b.rows, b.cols);
  to_ntt(p, a, b);
  return a;
}
// poly.rs:646-663
inline void from_ntt(const Params& p, PolyMatrix& a, const PolyMatrix& b) {
  std::vector<u64> scratch(p.crt_count * p.poly_len);
  for (size_t r = 0; r < a.rows; r++)
    for (size_t c = 0; c < a.cols; c++) {
      std::memcpy(scratch.data(), b.poly(r, c), scratch.size() * 8);
      ntt_inverse(p, scratch.data());
      u64* dst = a.poly(r, c);
      for (size_t z = 0; z < p.poly_len; z++) dst[z] = p.crt_compose(scratch.data(), z);
    }
}
inline PolyMatrix from_ntt_alloc(const Params& p, const PolyMatrix& b) {
  PolyMatrix a = raw_zero(p, b.rows, b.cols);
  from_ntt(p, a, b);
  return a;
}
// poly.rs:437-458 (scalar: Barrett after every multiply-add; the AVX2 build accumulates and
// reduces once, :460-481 — both yield the canonical residue)
inline void multiply(const Params& p, PolyMatrix& res, const PolyMatrix& a, const PolyMatrix& b) {
  if (res.rows != a.rows || res.cols != b.cols || a.cols != b.rows) throw std::runtime_error("multiply dims");
  size_t W = p.poly_len * p.crt_count;
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < b.cols; j++) {
      u64* rp = res.poly(i, j);
      for (size_t z = 0; z < W; z++) rp[z] = 0;
      for (size_t k = 0; k < a.cols; k++) {
        const u64* p1 = a.poly(i, k);
        const u64* p2 = b.poly(k, j);
        for (size_t c = 0; c < p.crt_count; c++)
          for (size_t z = 0; z < p.poly_len; z++) {
            size_t idx = c * p.poly_len + z;
            rp[idx] = p.barrett_coeff(p1[idx] * p2[idx] + rp[idx], c);
          }
      }
    }
}
inline PolyMatrix mul(const Params& p, const PolyMatrix& a, const PolyMatrix& b) {
  PolyMatrix r = ntt_zero(p, a.rows, b.cols);
  multiply(p, r, a, b);
  return r;
}
// poly.rs:483-498
inline void add(const Params& p, PolyMatrix& res, const PolyMatrix& a, const PolyMatrix& b) {
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < a.cols; j++)
      for (size_t c = 0; c < p.crt_count; c++)
        for (size_t z = 0; z < p.poly_len; z++) {
          size_t idx = c * p.poly_len + z;
          res.poly(i, j)[idx] = p.barrett_coeff(a.poly(i, j)[idx] + b.poly(i, j)[idx], c);
        }
}
inline PolyMatrix add_alloc(const Params& p, const PolyMatrix& a, const PolyMatrix& b) {
  PolyMatrix r = ntt_zero(p, a.rows, a.cols);
  add(p, r, a, b);
  return r;
}
// poly.rs:500-523
inline void add_into_at(const Params& p, PolyMatrix& res, const PolyMatrix& a, size_t tr, size_t tc) {
  for (size_t i = 0; i < a.rows; i++)
    for (size_t j = 0; j < a.cols; j++)
      for (size_t c = 0; c < p.crt_count; c++)
        for (size_t z = 0; z < p.poly_len; z++) {
          size_t idx = c * p.poly_len + z;
          u64* rp = res.poly(tr + i, tc + j);
          rp[idx] = p.barrett_coeff(rp[idx] + a.poly(i, j)[idx], c);
        }
}
inline void add_into(const Params& p, PolyMatrix& res, const PolyMatrix& a) { add_into_at(p, res, a, 0, 0); }
// poly.rs:387-391, 525-537 (NOTE: zero maps to the non-canonical value q, as in the reference)
inline void invert(const Params& p, PolyMatrix& res, const PolyMatrix& a) {
  for (size_t i = 0; i < a.data.size(); i++) res.data[i] = p.modulus - a.data[i];
}
// poly.rs:393-405, 539-551
inline void automorph(const Params& p, PolyMatrix& res, const PolyMatrix& a, size_t t) {
  size_t N = p.poly_len;
  for (size_t r = 0; r < a.rows; r++)
    for (size_t c = 0; c < a.cols; c++) {
      const u64* ap = a.poly(r, c);
      u64* rp = res.poly(r, c);
      for (size_t i = 0; i < N; i++) {
        size_t num = (i * t) / N, rem = (i * t) % N;
        rp[rem] = (num % 2 == 0) ? ap[i] : p.modulus - ap[i];
      }
    }
}
// poly.rs:575-588 : res = b (matrix) * a (1x1), pointwise, Barrett per product
inline void scalar_multiply(const Params& p, PolyMatrix& res, const PolyMatrix& a, const PolyMatrix& b) {
  const u64* p2 = a.poly(0, 0);
  for (size_t i = 0; i < b.rows; i++)
    for (size_t j = 0; j < b.cols; j++)
      for (size_t c = 0; c < p.crt_count; c++)
        for (size_t z = 0; z < p.poly_len; z++) {
          size_t idx = c * p.poly_len + z;
          res.poly(i, j)[idx] = p.barrett_coeff(b.poly(i, j)[idx] * p2[idx], c);
        }
}
// poly.rs:340-349, 567-573
inline PolyMatrix shift_rows_by_one(const Params& p, const PolyMatrix& inp) {
  if (inp.rows == 1) return inp;
  PolyMatrix out(p, inp.rows, inp.cols, inp.is_ntt);
  out.copy_into(inp.submatrix(p, inp.rows - 1, 0, 1, inp.cols), 0, 0);
  out.copy_into(inp.submatrix(p, 0, 0, inp.rows - 1, inp.cols), 1, 0);
  return out;
}
inline PolyMatrix stack(const Params& p, const PolyMatrix& a, const PolyMatrix& b) {   // poly.rs:559-565
  PolyMatrix c(p, a.rows + b.rows, a.cols, a.is_ntt);
  c.copy_into(a, 0, 0);
  c.copy_into(b, a.rows, 0);
  return c;
}

// ---------------------------------------------------------------- gadget.rs
// gadget.rs:3-9
inline size_t get_bits_per(const Params& p, size_t dim) {
  if ((u64)dim == p.modulus_log2) return 1;
  return (size_t)std::floor((double)p.modulus_log2 / (double)dim) + 1;
}
// gadget.rs:11-32
inline PolyMatrix build_gadget(const Params& p, size_t rows, size_t cols) {
  PolyMatrix g = raw_zero(p, rows, cols);
  size_t num_elems = cols / rows;
  size_t bits_per = get_bits_per(p, num_elems);
  for (size_t i = 0; i < rows; i++)
    for (size_t j = 0; j < num_elems; j++) {
      if (bits_per * j >= 64) continue;
      g.poly(i, i + j * rows)[0] = (u64)1 << (bits_per * j);
    }
  return g;
}
// gadget.rs:34-60
inline void gadget_invert_rdim(const Params& p, PolyMatrix& out, const PolyMatrix& inp, size_t rdim) {
  size_t mx = out.rows, num_elems = mx / rdim;
  size_t bits_per = get_bits_per(p, num_elems);
  u64 mask = ((u64)1 << bits_per) - 1;
  for (size_t i = 0; i < inp.cols; i++)
    for (size_t j = 0; j < rdim; j++)
      for (size_t z = 0; z < p.poly_len; z++) {
        u64 val = inp.poly(j, i)[z];
        for (size_t k = 0; k < num_elems; k++) {
          size_t bit_offs = std::min(k * bits_per, (size_t)64);
          u64 piece = bit_offs >= 64 ? 0 : ((val >> bit_offs) & mask);
          out.poly(j + k * rdim, i)[z] = piece;
        }
      }
}
inline void gadget_invert(const Params& p, PolyMatrix& out, const PolyMatrix& inp) { gadget_invert_rdim(p, out, inp, inp.rows); }

// ---------------------------------------------------------------- util.rs
// util.rs:36-44
inline size_t calc_index(const size_t* indices, const size_t* lengths, size_t n) {
  size_t idx = 0, prod = 1;
  for (size_t i = n; i-- > 0;) { idx += indices[i] * prod; prod *= lengths[i]; }
  return idx;
}
// util.rs:289-321.  Native-endian (little-endian on x86) 64/128-bit windows.
inline u64 read_arbitrary_bits(const uint8_t* data, size_t bit_offs, size_t num_bits) {
  size_t word_off = bit_offs / 64, within = bit_offs % 64;
  if (within + num_bits <= 64) {
    u64 v; std::memcpy(&v, data + word_off * 8, 8);
    return (v >> within) & (((u64)1 << num_bits) - 1);
  }
  u128 v; std::memcpy(&v, data + word_off * 8, 16);
  return (u64)((v >> within) & (((u128)1 << num_bits) - 1));
}
inline void write_arbitrary_bits(uint8_t* data, u64 val, size_t bit_offs, size_t num_bits) {
  size_t word_off = bit_offs / 64, within = bit_offs % 64;
  val &= ((u64)1 << num_bits) - 1;
  if (within + num_bits <= 64) {
    u64 cur; std::memcpy(&cur, data + word_off * 8, 8);
    cur &= ~((((u64)1 << num_bits) - 1) << within);
    cur |= val << within;
    std::memcpy(data + word_off * 8, &cur, 8);
  } else {
    u128 cur; std::memcpy(&cur, data + word_off * 8, 16);
    cur &= ~((((u128)1 << num_bits) - 1) << within);
    cur |= (u128)val << within;
    std::memcpy(data + word_off * 8, &cur, 16);
  }
}
// util.rs:323-355 : v_reg[j] (2x1 NTT) -> out[z][j][r] = lo | hi<<32
inline void reorient_reg_ciphertexts(const Params& p, u64* out, const std::vector<PolyMatrix>& v_reg) {
  size_t N = p.poly_len, dim0 = (size_t)1 << p.db_dim_1;
  for (size_t j = 0; j < dim0; j++)
    for (size_t r = 0; r < 2; r++)
      for (size_t z = 0; z < N; z++) {
        size_t idx_in = r * (p.crt_count * N);
        size_t idx_out = z * (dim0 * 2) + j * 2 + r;
        u64 v1 = v_reg[j].data[idx_in + z] % p.moduli[0];
        u64 v2 = v_reg[j].data[idx_in + N + z] % p.moduli[1];
        out[idx_out] = v1 | (v2 << 32);
      }
}

// ---------------------------------------------------------------- server.rs
// params.rs:98-107
inline std::vector<PolyMatrix> get_v_neg1(const Params& p) {
  std::vector<PolyMatrix> v;
  for (size_t i = 0; i < p.poly_len_log2; i++) {
    size_t idx = p.poly_len - ((size_t)1 << i);
    PolyMatrix ng1 = raw_zero(p, 1, 1);
    ng1.data[idx] = 1;
    PolyMatrix neg = raw_zero(p, 1, 1);
    invert(p, neg, ng1);
    v.push_back(to_ntt_alloc(p, neg));
  }
  return v;
}

// server.rs:19-121
inline void coefficient_expansion(const Params& p, std::vector<PolyMatrix>& v, size_t g, size_t stop_round,
                                  const std::vector<PolyMatrix>& v_w_left, const std::vector<PolyMatrix>& v_w_right,
                                  const std::vector<PolyMatrix>& v_neg1, size_t max_bits_to_gen_right) {
  size_t N = p.poly_len;
  for (size_t r = 0; r < g; r++) {
    size_t num_in = (size_t)1 << r, num_out = 2 * num_in;
    size_t t = (N / ((size_t)1 << r)) + 1;
    const PolyMatrix& neg1 = v_neg1[r];
    for (size_t i = 0; i < num_in; i++) scalar_multiply(p, v[num_in + i], neg1, v[i]);     // :105-110
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < num_out; i++) {
      // NOTE: the reference enumerates each half separately (:112-119), so the parity test at
      // :40-44 sees the index WITHIN the half.
      size_t ih = i < num_in ? i : i - num_in;
      if ((stop_round > 0 && r > stop_round && (ih % 2) == 1) ||
          (stop_round > 0 && r == stop_round && (ih % 2) == 1 && (ih / 2) >= max_bits_to_gen_right))
        continue;
      PolyMatrix& v_i = v[i];
      bool left = (r != 0) && (ih % 2 == 0);
      const PolyMatrix& w = left ? v_w_left[r] : v_w_right[r];
      size_t gadget_dim = left ? p.t_exp_left : p.t_exp_right;
      PolyMatrix ct = raw_zero(p, 2, 1), ct_auto = raw_zero(p, 2, 1);
      from_ntt(p, ct, v_i);
      automorph(p, ct_auto, ct, t);
      PolyMatrix gi_ct = raw_zero(p, gadget_dim, 1), gi_ct_ntt = ntt_zero(p, gadget_dim, 1);
      gadget_invert_rdim(p, gi_ct, ct_auto, 1);
      to_ntt_no_reduce(p, gi_ct_ntt, gi_ct);
      PolyMatrix ct_auto_1 = raw_zero(p, 1, 1);
      std::memcpy(ct_auto_1.data.data(), ct_auto.poly(1, 0), N * 8);
      PolyMatrix ct_auto_1_ntt = to_ntt_alloc(p, ct_auto_1);
      PolyMatrix w_times = mul(p, w, gi_ct_ntt);
      size_t idx = 0;
      for (size_t j = 0; j < 2; j++)
        for (size_t nn = 0; nn < p.crt_count; nn++)
          for (size_t z = 0; z < N; z++) {
            u64 sum = v_i.data[idx] + w_times.data[idx] + j * ct_auto_1_ntt.data[nn * N + z];
            v_i.data[idx] = p.barrett_coeff(sum, nn);
            idx++;
          }
    }
  }
}

// server.rs:123-151
inline void regev_to_gsw(const Params& p, std::vector<PolyMatrix>& v_gsw, const std::vector<PolyMatrix>& v_inp,
                         const PolyMatrix& v, size_t idx_factor, size_t idx_offset) {
#pragma omp parallel for schedule(dynamic)
  for (size_t i = 0; i < v_gsw.size(); i++) {
    PolyMatrix& ct = v_gsw[i];
    for (size_t j = 0; j < p.t_gsw; j++) {
      size_t idx_ct = i * p.t_gsw + j;
      size_t idx_inp = idx_factor * idx_ct + idx_offset;
      ct.copy_into(v_inp[idx_inp], 0, 2 * j + 1);
      PolyMatrix tmp_raw = from_ntt_alloc(p, v_inp[idx_inp]);
      PolyMatrix ginv = raw_zero(p, 2 * p.t_conv, 1);
      gadget_invert(p, ginv, tmp_raw);
      PolyMatrix ginv_ntt = to_ntt_alloc(p, ginv);
      PolyMatrix tmp_ct = mul(p, v, ginv_ntt);
      ct.copy_into(tmp_ct, 0, 2 * j);
    }
  }
}

#if defined(__AVX2__)
// AVX2 form of the same product, in the style of the reference's vectorised kernel
// (lib/server/src/compute/dot_product.rs:59-96: _mm256_mul_epu32 + _mm256_add_epi64 on 4 words at a time), but with
// an exact reduction schedule: every 64-bit lane is folded mod q_n before it can overflow (<= 256 products of < 2^56),
// so it equals the u128 path for every input (the reference's counter-based schedule does not, SURVEY 7 "hazards").
// Used only to give the CPU baseline the SIMD the reference builds with (.cargo/config.toml: target-cpu=native).
inline void multiply_reg_by_database_avx2(const Params& p, u64* out, const u64* db, const u64* v_firstdim, size_t dim0,
                                          size_t num_per) {
  size_t N = p.poly_len;
  const u64 q0 = p.moduli[0], q1 = p.moduli[1];
#pragma omp parallel for schedule(static)
  for (size_t z = 0; z < N; z++) {
    const u64* a = v_firstdim + z * dim0 * 2;
    const u64* b = db + z * num_per * dim0;
    for (size_t i = 0; i < num_per; i++) {
      u64 tot[4] = {0, 0, 0, 0};                       // n0r0, n0r1, n1r0, n1r1 (already reduced)
      for (size_t j0 = 0; j0 < dim0; j0 += 512) {
        size_t jend = std::min(dim0, j0 + 512);
        __m256i acc_lo = _mm256_setzero_si256(), acc_hi = _mm256_setzero_si256();   // lanes: (j even r0, r1, j odd r0, r1)
        for (size_t j = j0; j < jend; j += 2) {
          __m128i bw = _mm_loadu_si128((const __m128i*)(b + i * dim0 + j));           // db words j, j+1
          __m256i bb = _mm256_permute4x64_epi64(_mm256_castsi128_si256(bw), 0x50);    // (b_j, b_j, b_j+1, b_j+1)
          __m256i av = _mm256_loadu_si256((const __m256i*)(a + 2 * j));               // (a_j r0, a_j r1, a_j+1 r0, a_j+1 r1)
          acc_lo = _mm256_add_epi64(acc_lo, _mm256_mul_epu32(av, bb));
          acc_hi = _mm256_add_epi64(acc_hi, _mm256_mul_epu32(_mm256_srli_epi64(av, 32), _mm256_srli_epi64(bb, 32)));
        }
        alignas(32) u64 lo[4], hi[4];
        _mm256_store_si256((__m256i*)lo, acc_lo);
        _mm256_store_si256((__m256i*)hi, acc_hi);
        tot[0] = (tot[0] + lo[0] % q0 + lo[2] % q0) % q0;
        tot[1] = (tot[1] + lo[1] % q0 + lo[3] % q0) % q0;
        tot[2] = (tot[2] + hi[0] % q1 + hi[2] % q1) % q1;
        tot[3] = (tot[3] + hi[1] % q1 + hi[3] % q1) % q1;
      }
      u64* o = out + i * 4 * N;
      o[z] = tot[0]; o[2 * N + z] = tot[1]; o[N + z] = tot[2]; o[3 * N + z] = tot[3];
    }
  }
}
#endif

// server.rs:155-221 (u128 accumulate, one % per output)
inline void multiply_reg_by_database(const Params& p, u64* out /*[num_per][4*N]*/, const u64* db, const u64* v_firstdim,
                                     size_t dim0, size_t num_per) {
  size_t N = p.poly_len;
#pragma omp parallel for schedule(static)
  for (size_t z = 0; z < N; z++) {
    const u64* a = v_firstdim + z * dim0 * 2;
    const u64* b = db + z * num_per * dim0;
    for (size_t i = 0; i < num_per; i++) {
      u128 s00 = 0, s01 = 0, s10 = 0, s11 = 0;
      for (size_t j = 0; j < dim0; j++) {
        u64 bw = b[i * dim0 + j];
        u64 a0 = a[2 * j], a1 = a[2 * j + 1];
        u64 b_lo = (u32)bw, b_hi = bw >> 32;
        s00 += (u128)((u64)(u32)a0 * b_lo);
        s01 += (u128)((u64)(u32)a1 * b_lo);
        s10 += (u128)((a0 >> 32) * b_hi);
        s11 += (u128)((a1 >> 32) * b_hi);
      }
      u64* o = out + i * 4 * N;
      o[z] = (u64)(s00 % p.moduli[0]);
      o[2 * N + z] = (u64)(s01 % p.moduli[0]);
      o[N + z] = (u64)(s10 % p.moduli[1]);
      o[3 * N + z] = (u64)(s11 % p.moduli[1]);
    }
  }
}

// server.rs:388-427 ; sparse_shortcut=true adds lib/server/src/compute/fold.rs:37-43
inline void fold_ciphertexts(const Params& p, std::vector<PolyMatrix>& v_cts, const std::vector<PolyMatrix>& v_folding,
                             const std::vector<PolyMatrix>& v_folding_neg, bool sparse_shortcut = false) {
  if (v_cts.size() == 1) return;
  size_t further_dims = log2_floor(v_cts.size());
  size_t ell = v_folding[0].cols / 2;
  auto all_zero = [](const PolyMatrix& m) {
    for (u64 x : m.data) if (x) return false;
    return true;
  };
  size_t num_per = v_cts.size();
  for (size_t cur_dim = 0; cur_dim < further_dims; cur_dim++) {
    num_per /= 2;
    // iterations of one round touch disjoint ciphertext pairs (i, num_per+i): rayon-free but
    // data-parallel, so the CPU baseline may use all cores here.
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < num_per; i++) {
      PolyMatrix ginv_c = raw_zero(p, 2 * ell, 1), ginv_c_ntt = ntt_zero(p, 2 * ell, 1);
      PolyMatrix prod = ntt_zero(p, 2, 1), sum = ntt_zero(p, 2, 1);
      if (sparse_shortcut) {
        if (all_zero(v_cts[i])) { v_cts[i] = v_cts[num_per + i]; continue; }
        else if (all_zero(v_cts[num_per + i])) continue;
      }
      gadget_invert(p, ginv_c, v_cts[i]);
      to_ntt(p, ginv_c_ntt, ginv_c);
      multiply(p, prod, v_folding_neg[further_dims - 1 - cur_dim], ginv_c_ntt);
      gadget_invert(p, ginv_c, v_cts[num_per + i]);
      to_ntt(p, ginv_c_ntt, ginv_c);
      multiply(p, sum, v_folding[further_dims - 1 - cur_dim], ginv_c_ntt);
      add_into(p, sum, prod);
      from_ntt(p, v_cts[i], sum);
    }
  }
}

// server.rs:429-468 (== lib/server/src/compute/pack.rs:5-43, v0)
inline PolyMatrix pack_v0(const Params& p, const PolyMatrix* v_ct, const std::vector<PolyMatrix>& v_w) {
  if (v_w.size() != p.n) throw std::runtime_error("pack_v0: need n packing matrices");
  size_t N = p.poly_len;
  PolyMatrix result = ntt_zero(p, p.n + 1, p.n);
  PolyMatrix ginv = raw_zero(p, p.t_conv, 1), ginv_ntt = ntt_zero(p, p.t_conv, 1);
  PolyMatrix prod = ntt_zero(p, p.n + 1, 1);
  PolyMatrix ct_1 = raw_zero(p, 1, 1), ct_2 = raw_zero(p, 1, 1), ct_2_ntt = ntt_zero(p, 1, 1);
  for (size_t c = 0; c < p.n; c++) {
    PolyMatrix v_int = ntt_zero(p, p.n + 1, 1);
    for (size_t r = 0; r < p.n; r++) {
      const PolyMatrix& w = v_w[r];
      const PolyMatrix& ct = v_ct[r * p.n + c];
      std::memcpy(ct_1.data.data(), ct.poly(0, 0), N * 8);
      std::memcpy(ct_2.data.data(), ct.poly(1, 0), N * 8);
      to_ntt(p, ct_2_ntt, ct_2);
      gadget_invert(p, ginv, ct_1);
      to_ntt(p, ginv_ntt, ginv);
      multiply(p, prod, w, ginv_ntt);
      add_into_at(p, v_int, ct_2_ntt, 1 + r, 0);
      add_into(p, v_int, prod);
    }
    result.copy_into(v_int, 0, c);
  }
  return result;
}
// lib/server/src/compute/pack.rs:45-98 (v1: one key matrix + row shifts)
inline PolyMatrix pack_v1(const Params& p, const PolyMatrix* v_ct, const std::vector<PolyMatrix>& v_w) {
  if (v_w.size() != 2) throw std::runtime_error("pack_v1: need 2 packing matrices");
  size_t N = p.poly_len;
  const PolyMatrix& w_key = v_w[0];
  const PolyMatrix& w_shift = v_w[1];
  PolyMatrix result = ntt_zero(p, p.n + 1, p.n);
  PolyMatrix ginv = raw_zero(p, p.t_conv, 1), ginv_ntt = ntt_zero(p, p.t_conv, 1);
  PolyMatrix ct_1 = raw_zero(p, 1, 1), ct_2 = raw_zero(p, 1, 1), ct_2_ntt = ntt_zero(p, 1, 1);
  for (size_t c = 0; c < p.n; c++) {
    PolyMatrix v_int = ntt_zero(p, p.n + 1, 1);
    for (size_t r = 0; r < p.n; r++) {
      const PolyMatrix& ct = v_ct[r * p.n + c];
      std::memcpy(ct_1.data.data(), ct.poly(0, 0), N * 8);
      std::memcpy(ct_2.data.data(), ct.poly(1, 0), N * 8);
      to_ntt(p, ct_2_ntt, ct_2);
      gadget_invert(p, ginv, ct_1);
      to_ntt(p, ginv_ntt, ginv);
      PolyMatrix prod = mul(p, w_key, ginv_ntt);
      add_into_at(p, prod, ct_2_ntt, 1, 0);
      for (size_t s = 0; s < r; s++) {
        PolyMatrix prod_ct_1 = prod.submatrix(p, 0, 0, 1, 1);
        PolyMatrix prod_rest = prod.submatrix(p, 1, 0, prod.rows - 1, 1);
        PolyMatrix gi = raw_zero(p, p.t_conv, 1);
        gadget_invert(p, gi, from_ntt_alloc(p, prod_ct_1));
        PolyMatrix part1 = mul(p, w_shift, to_ntt_alloc(p, gi));
        PolyMatrix part2 = shift_rows_by_one(p, prod_rest).pad_top(p, 1);
        prod = add_alloc(p, part1, part2);
      }
      add_into(p, v_int, prod);
    }
    result.copy_into(v_int, 0, c);
  }
  return result;
}
inline PolyMatrix pack(const Params& p, const PolyMatrix* v_ct, const std::vector<PolyMatrix>& v_w) {
  if (p.version == 0) return pack_v0(p, v_ct, v_w);       // lib/server/src/compute/pack.rs:100-112
  if (p.version == 1) return pack_v1(p, v_ct, v_w);
  throw std::runtime_error("unknown version");
}

// server.rs:470-503
inline std::vector<uint8_t> encode(const Params& p, const std::vector<PolyMatrix>& v_packed_ct) {
  u64 q1 = 4 * p.pt_modulus;
  size_t q1_bits = log2_ceil(q1);
  u64 q2 = Q2_VALUES[p.q2_bits];
  size_t q2_bits = p.q2_bits;
  size_t N = p.poly_len;
  size_t num_bits = p.instances * ((q2_bits * p.n * N) + (q1_bits * p.n * p.n * N));
  size_t num_bytes = ((num_bits + 63) / 64) * 64 / 8;
  std::vector<uint8_t> result(num_bytes + 8, 0);   // +8: slack for 128-bit windows at the tail
  size_t bit_offs = 0;
  for (size_t inst = 0; inst < p.instances; inst++) {
    const PolyMatrix& ct = v_packed_ct[inst];
    for (size_t i = 0; i < p.n * N; i++) {
      write_arbitrary_bits(result.data(), rescale(ct.data[i], p.modulus, q2), bit_offs, q2_bits);
      bit_offs += q2_bits;
    }
    for (size_t i = 0; i < p.n * p.n * N; i++) {
      write_arbitrary_bits(result.data(), rescale(ct.data[p.n * N + i], p.modulus, q1), bit_offs, q1_bits);
      bit_offs += q1_bits;
    }
  }
  result.resize(num_bytes);
  return result;
}

// server.rs:505-523
inline std::vector<PolyMatrix> get_v_folding_neg(const Params& p, const std::vector<PolyMatrix>& v_folding) {
  PolyMatrix gadget_ntt = to_ntt_alloc(p, build_gadget(p, 2, 2 * p.t_gsw));
  std::vector<PolyMatrix> out;
  for (size_t i = 0; i < p.db_dim_2; i++) {
    PolyMatrix inv = raw_zero(p, 2, 2 * p.t_gsw);
    invert(p, inv, from_ntt_alloc(p, v_folding[i]));
    out.push_back(add_alloc(p, gadget_ntt, to_ntt_alloc(p, inv)));
  }
  return out;
}

struct PublicParameters {       // client.rs:146-152 (all in NTT form)
  std::vector<PolyMatrix> v_packing, v_expansion_left, v_expansion_right, v_conversion;
  bool has_right = false;
};

// server.rs:525-591
inline void expand_query(const Params& p, const PublicParameters& pp, const PolyMatrix& query_ct /*raw 2x1*/,
                         std::vector<u64>& v_reg_reoriented, std::vector<PolyMatrix>& v_folding) {
  size_t dim0 = (size_t)1 << p.db_dim_1, further_dims = p.db_dim_2;
  size_t num_bits_to_gen = p.t_gsw * further_dims + dim0;
  size_t g = log2_ceil(num_bits_to_gen);
  size_t right_expanded = p.t_gsw * further_dims;
  size_t stop_round = log2_ceil(right_expanded);
  std::vector<PolyMatrix> v((size_t)1 << g, ntt_zero(p, 2, 1));
  v[0] = to_ntt_alloc(p, query_ct);
  const std::vector<PolyMatrix>& v_w_left = pp.v_expansion_left;
  const std::vector<PolyMatrix>& v_w_right = pp.has_right ? pp.v_expansion_right : pp.v_expansion_left;
  std::vector<PolyMatrix> v_neg1 = get_v_neg1(p);
  std::vector<PolyMatrix> v_reg_inp, v_gsw_inp;
  if (further_dims > 0) {
    coefficient_expansion(p, v, g, stop_round, v_w_left, v_w_right, v_neg1, p.t_gsw * p.db_dim_2);
    for (size_t i = 0; i < dim0; i++) v_reg_inp.push_back(v[2 * i]);
    for (size_t i = 0; i < right_expanded; i++) v_gsw_inp.push_back(v[2 * i + 1]);
  } else {
    coefficient_expansion(p, v, g, 0, v_w_left, v_w_left, v_neg1, 0);
    for (size_t i = 0; i < dim0; i++) v_reg_inp.push_back(v[i]);
  }
  v_reg_reoriented.assign(dim0 * 2 * p.poly_len, 0);
  reorient_reg_ciphertexts(p, v_reg_reoriented.data(), v_reg_inp);
  v_folding.assign(p.db_dim_2, ntt_zero(p, 2, 2 * p.t_gsw));
  regev_to_gsw(p, v_folding, v_gsw_inp, pp.v_conversion[0], 1, 0);
}

struct Query {                 // client.rs:262-267
  PolyMatrix ct;               // expand_queries: raw 2x1
  std::vector<u64> v_buf;      // direct upload: [z][j][r] packed
  std::vector<PolyMatrix> v_ct;// direct upload: raw 2 x 2t_gsw, one per further dim
};

// server.rs:650-741 (dense).  `stage_out` (optional) receives intermediates for stage-level parity.
struct StageDump {
  std::vector<u64> v_firstdim;               // [z][j][r]
  std::vector<PolyMatrix> v_folding, v_folding_neg;
  std::vector<u64> first_mult;               // slice 0: [num_per][4N]
  std::vector<PolyMatrix> folded;            // one raw 2x1 per (instance,trial)
  std::vector<PolyMatrix> packed;            // one raw (n+1) x n per instance
};
inline std::vector<uint8_t> process_query(const Params& p, const PublicParameters& pp, const Query& query, const u64* db,
                                          StageDump* dump = nullptr) {
  size_t dim0 = (size_t)1 << p.db_dim_1, num_per = (size_t)1 << p.db_dim_2;
  size_t N = p.poly_len;
  size_t db_slice_sz = dim0 * num_per * N;
  std::vector<u64> v_reg;
  std::vector<PolyMatrix> v_folding;
  if (p.expand_queries) {
    expand_query(p, pp, query.ct, v_reg, v_folding);
  } else {
    v_reg = query.v_buf;
    for (auto& m : query.v_ct) v_folding.push_back(to_ntt_alloc(p, m));
  }
  std::vector<PolyMatrix> v_folding_neg = get_v_folding_neg(p, v_folding);
  size_t trials = p.n * p.n;
  std::vector<PolyMatrix> v_ct_all(p.instances * trials);
  std::vector<u64> first_mult;
  for (size_t it = 0; it < p.instances * trials; it++) {
    std::vector<u64> inter(num_per * 4 * N);
    const u64* cur_db = db + it * db_slice_sz;
#if defined(__AVX2__)
    if (g_use_avx2_multiply && (dim0 % 2) == 0) multiply_reg_by_database_avx2(p, inter.data(), cur_db, v_reg.data(), dim0, num_per);
    else
#endif
    multiply_reg_by_database(p, inter.data(), cur_db, v_reg.data(), dim0, num_per);
    if (dump && it == 0) first_mult = inter;
    std::vector<PolyMatrix> inter_raw(num_per, raw_zero(p, 2, 1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < num_per; i++) {
      PolyMatrix m = ntt_zero(p, 2, 1);
      std::memcpy(m.data.data(), inter.data() + i * 4 * N, 4 * N * 8);
      from_ntt(p, inter_raw[i], m);
    }
    fold_ciphertexts(p, inter_raw, v_folding, v_folding_neg, g_sparse_fold);   // lib/server twin: compute/fold.rs shortcut
    v_ct_all[it] = inter_raw[0];
  }
  std::vector<PolyMatrix> v_packed;
  for (size_t inst = 0; inst < p.instances; inst++) {
    PolyMatrix packed = pack(p, v_ct_all.data() + inst * trials, pp.v_packing);
    v_packed.push_back(from_ntt_alloc(p, packed));
  }
  if (dump) {
    dump->v_firstdim = v_reg; dump->v_folding = v_folding; dump->v_folding_neg = v_folding_neg;
    dump->first_mult = first_mult; dump->folded = v_ct_all; dump->packed = v_packed;
  }
  return encode(p, v_packed);
}

}  // namespace orc
