// CPU emulation of the 512-thread cooperative 4096-point NTT (sdk_b200/csrc/ntt_core4096.cuh), pass by pass, against the
// oracle's scalar transforms instantiated with poly_len = 4096 (the oracle's ntt.rs restatement is generic in the size;
// tables from the same construction, ntt.rs:39-65).  Also checks the shared-memory padding for bank conflicts.
#include "../../sdk_b200/csrc/ntt_core4096.cuh"
#include "../../oracle/spiral_oracle.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace b200pir;

int main() {
  orc::Params p = orc::params_init(4096, {268369921ULL, 249561089ULL}, 6.4, 2, 256, 20, 4, 8, 8, 8, true, 6, 2, 1, 8192, 0);
  std::mt19937_64 rng(321);
  int bad = 0;
  const int N = NTT4K_N, T = NTT4K_THREADS;
  for (int mod = 0; mod < 2; mod++) {
    uint32_t q = (uint32_t)p.moduli[mod], two_q = 2 * q;
    std::vector<Twiddle> fwd(N), inv(N);
    for (int i = 0; i < N; i++) {
      fwd[i] = {(uint32_t)p.ntt_tables[mod][0][i], (uint32_t)p.ntt_tables[mod][1][i]};
      inv[i] = {(uint32_t)p.ntt_tables[mod][2][i], (uint32_t)p.ntt_tables[mod][3][i]};
    }
    for (int trial = 0; trial < 6; trial++) {
      std::vector<uint64_t> ref(2 * N, 0);
      std::vector<uint32_t> in(N);
      for (int i = 0; i < N; i++) {
        uint64_t v = rng() % q;
        if (trial == 1) v = (i == 0) ? 100 : 0;
        if (trial == 2) v = q - 1;
        if (trial == 3) v = rng() % (4ull * q);
        if (trial == 4) v = (i & 1) ? q - 1 : 0;
        in[i] = (uint32_t)v;
        ref[mod * N + i] = v;
      }
      orc::ntt_forward(p, ref.data());
      static uint32_t regs[NTT4K_THREADS][8];
      std::vector<uint32_t> smem(NTT4K_SMEM_WORDS, 0xDEADBEEF);
      for (int t = 0; t < T; t++) for (int a = 0; a < 8; a++) regs[t][a] = in[a * T + t];
      for (int t = 0; t < T; t++) fwd4k_pass_a(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < T; t++) fwd4k_pass_b(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < T; t++) fwd4k_pass_c(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < T; t++) fwd4k_pass_d(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < T; t++) for (int k = 0; k < 8; k++)
        if (regs[t][k] != ref[mod * N + t * 8 + k]) { if (bad < 5) printf("fwd mismatch mod %d trial %d at %d\n", mod, trial, t * 8 + k); bad++; }
      std::vector<uint64_t> ref2 = ref;
      orc::ntt_inverse(p, ref2.data());
      std::fill(smem.begin(), smem.end(), 0xDEADBEEF);
      for (int t = 0; t < T; t++) inv4k_pass_d(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < T; t++) inv4k_pass_c(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < T; t++) inv4k_pass_b(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < T; t++) inv4k_pass_a(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < T; t++) for (int a = 0; a < 8; a++)
        if (regs[t][a] != ref2[mod * N + a * T + t]) { if (bad < 5) printf("inv mismatch mod %d trial %d at %d\n", mod, trial, a * T + t); bad++; }
      // round trip returns the canonical input
      if (trial != 3)
        for (int i = 0; i < N; i++) if (ref2[mod * N + i] != in[i]) { if (bad < 5) printf("round trip mismatch\n"); bad++; }
    }
  }
  auto check32 = [&](auto addr_of_lane, const char* name) {
    for (int warp = 0; warp < T / 32; warp++) {
      int seen[32] = {0};
      for (int lane = 0; lane < 32; lane++) { int b = addr_of_lane(warp * 32 + lane) & 31; if (seen[b]++) { printf("bank conflict in %s\n", name); bad++; return; } }
    }
  };
  for (int a = 0; a < 8; a++) {
    check32([&](int tid) { return ntt_phys(a * 512 + tid); }, "pass A");
    check32([&](int tid) { return ntt_phys((tid >> 6) * 512 + a * 64 + (tid & 63)); }, "pass B");
    check32([&](int tid) { return ntt_phys((tid >> 3) * 64 + a * 8 + (tid & 7)); }, "pass C");
  }
  for (int quarter = 0; quarter < T / 8; quarter++) {
    int seen[8] = {0};
    for (int l = 0; l < 8; l++) {
      int w = ntt_phys((quarter * 8 + l) * 8);
      if (w % 4) { printf("pass D misaligned\n"); bad++; }
      int grp = (w & 31) >> 2;
      if (seen[grp]++) { printf("bank conflict in pass D\n"); bad++; }
    }
  }
  if (ntt_phys(N - 1) >= NTT4K_SMEM_WORDS) { printf("padding exceeds the buffer\n"); bad++; }
  printf(bad ? "FAIL %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
