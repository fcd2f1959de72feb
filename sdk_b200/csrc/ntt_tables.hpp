// Host-side construction of the NTT tables (plain C++, no CUDA): an independent re-derivation of Params::init
// (lib/spiral-rs/src/params.rs:224-296) / build_ntt_tables (ntt.rs:39-65); the oracle is never linked into the product.
// Shared by api.cu and the CPU emulation tests (tests/cpp/ntt_core_emul.cpp).
#pragma once
#include <stdint.h>
#include <stdexcept>
#include <vector>
#include "ntt_core.cuh"

namespace b200pir {
namespace tables {

typedef unsigned __int128 u128;
inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a * b % m); }
inline uint64_t powmod(uint64_t a, uint64_t e, uint64_t m) {
  uint64_t r = 1 % m;
  a %= m;
  while (e) { if (e & 1) r = mulmod(r, a, m); a = mulmod(a, a, m); e >>= 1; }
  return r;
}
inline uint64_t invmod(uint64_t a, uint64_t m) {    // m prime or gcd(a,m)=1
  __int128 r0 = a % m, r1 = m, s0 = 1, s1 = 0;
  while (r1 != 0) { __int128 q = r0 / r1, t = r0 - q * r1; r0 = r1; r1 = t; t = s0 - q * s1; s0 = s1; s1 = t; }
  if (r0 != 1) throw std::invalid_argument("invmod: not invertible");
  s0 %= (__int128)m;
  if (s0 < 0) s0 += m;
  return (uint64_t)s0;
}
inline unsigned bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}
// minimal primitive 2N-th root (number_theory.rs:14-55)
inline uint64_t min_primitive_root(uint64_t degree, uint64_t q) {
  if ((q - 1) % degree) throw std::invalid_argument("modulus is not NTT friendly");
  uint64_t quot = (q - 1) / degree, root = 0;
  for (uint64_t c = 2; c < 4096; c++) {
    uint64_t r = powmod(c, quot, q);
    if (powmod(r, degree / 2, q) == q - 1) { root = r; break; }
  }
  if (!root) throw std::invalid_argument("no primitive root found");
  uint64_t gsq = mulmod(root, root, q), cur = root, best = root;
  for (uint64_t i = 0; i < degree; i++) { if (cur < best) best = cur; cur = mulmod(cur, gsq, q); }
  return best;
}
inline Twiddle shoup_pair(uint64_t w, uint64_t q) { return Twiddle{(uint32_t)w, (uint32_t)((w << 32) / q)}; }   // scale_powers_u32, ntt.rs:29-37
// tables of ntt.rs:39-65 as (W, W') pairs: fwd[bitrev(i)] = psi^i, inv[bitrev(i)] = div2(psi^-i) (arith.rs:78-89)
inline void build_tables(uint64_t q, std::vector<Twiddle>& fwd, std::vector<Twiddle>& inv, int N = NTT_N, int LG = NTT_LOG_N) {
  uint64_t root = min_primitive_root(2 * N, q), iroot = invmod(root, q);
  fwd.assign(N, Twiddle{0, 0});
  inv.assign(N, Twiddle{0, 0});
  auto fill = [&](std::vector<Twiddle>& t, uint64_t r, bool halve) {
    uint64_t power = r;
    std::vector<uint64_t> v(N, 0);
    for (int i = 1; i < N; i++) { v[bitrev(i, LG)] = power; power = mulmod(power, r, q); }
    v[0] = 1;
    for (int i = 0; i < N; i++) {
      uint64_t w = v[i];
      if (halve) w = (w & 1) ? (w + q) >> 1 : w >> 1;          // div2_uint_mod
      t[i] = shoup_pair(w, q);
    }
  };
  fill(fwd, root, false);
  fill(inv, iroot, true);
}
// inverse table of the relaxed-range transform (ntt_core.cuh "lz"): un-halved powers psi^-i in the same order, with the
// factor 1/N folded into the last stage: entry [0] = 1/N, entry [1] = psi^{-N/2} / N  (entry 1 is the last stage's only twiddle)
inline void build_inverse_table_lz(uint64_t q, std::vector<Twiddle>& inv, int N = NTT_N, int LG = NTT_LOG_N) {
  uint64_t iroot = invmod(min_primitive_root(2 * N, q), q), ninv = invmod((uint64_t)N % q, q);
  inv.assign(N, Twiddle{0, 0});
  uint64_t power = iroot;
  for (int i = 1; i < N; i++) { inv[bitrev(i, LG)] = shoup_pair(power, q); power = mulmod(power, iroot, q); }
  inv[0] = shoup_pair(ninv, q);
  inv[1] = shoup_pair(mulmod(inv[1].w, ninv, q), q);
}

}  // namespace tables
}  // namespace b200pir
