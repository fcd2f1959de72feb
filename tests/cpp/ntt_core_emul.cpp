// CPU emulation of the 256-thread cooperative NTT in sdk_b200/csrc/ntt_core.cuh: the per-pass
// thread-local functions are run for tid = 0..255 with an explicit "shared memory" array, pass by
// pass (a pass boundary = the group barrier), and compared with the oracle's scalar transforms.
// This validates every index / twiddle / padding computation without a GPU.
#include "../../sdk_b200/csrc/ntt_core.cuh"
#include "../../oracle/spiral_oracle.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace b200pir;

int main() {
  orc::Params p = orc::params_from_scalars(2, 6, 2, 256, 20, 8, 4, 8, 8, 1, 8192, 0, true);
  std::mt19937_64 rng(123);
  int bad = 0;
  for (int mod = 0; mod < 2; mod++) {
    uint32_t q = (uint32_t)p.moduli[mod], two_q = 2 * q;
    std::vector<Twiddle> fwd(2048), inv(2048);
    for (int i = 0; i < 2048; i++) {
      fwd[i] = {(uint32_t)p.ntt_tables[mod][0][i], (uint32_t)p.ntt_tables[mod][1][i]};
      inv[i] = {(uint32_t)p.ntt_tables[mod][2][i], (uint32_t)p.ntt_tables[mod][3][i]};
    }
    for (int trial = 0; trial < 6; trial++) {
      std::vector<uint64_t> ref(4096, 0);
      std::vector<uint32_t> in(2048);
      for (int i = 0; i < 2048; i++) {
        uint64_t v = rng() % q;
        if (trial == 1) v = (i == 0) ? 100 : 0;
        if (trial == 2) v = q - 1;
        if (trial == 3) v = rng() % (4ull * q);     // lazy-range input (to_ntt_no_reduce contract: < 4q)
        in[i] = (uint32_t)v;
        ref[mod * 2048 + i] = v;
      }
      orc::ntt_forward(p, ref.data());
      // ---- forward emulation
      static uint32_t regs[256][8];
      std::vector<uint32_t> smem(NTT_SMEM_WORDS, 0xDEADBEEF);
      for (int t = 0; t < 256; t++) for (int a = 0; a < 8; a++) regs[t][a] = in[a * 256 + t];
      for (int t = 0; t < 256; t++) fwd_pass_a(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < 256; t++) fwd_pass_b(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < 256; t++) fwd_pass_c(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < 256; t++) fwd_pass_d(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
      for (int t = 0; t < 256; t++) for (int k = 0; k < 8; k++)
        if (regs[t][k] != ref[mod * 2048 + t * 8 + k]) { if (bad < 5) printf("fwd mismatch mod %d trial %d at %d\n", mod, trial, t * 8 + k); bad++; }
      // ---- inverse emulation (input: canonical forward output, contiguous layout)
      std::vector<uint64_t> ref2 = ref;
      orc::ntt_inverse(p, ref2.data());
      std::fill(smem.begin(), smem.end(), 0xDEADBEEF);
      for (int t = 0; t < 256; t++) inv_pass_d(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_c(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_b(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_a(t, regs[t], smem.data(), TwArray{inv.data()}, q, two_q);
      for (int t = 0; t < 256; t++) for (int a = 0; a < 8; a++)
        if (regs[t][a] != ref2[mod * 2048 + a * 256 + t]) { if (bad < 5) printf("inv mismatch mod %d trial %d at %d\n", mod, trial, a * 256 + t); bad++; }
    }
  }
  // bank-conflict check of the padded layout: every warp-wide 32-bit access pattern used by the
  // passes must hit 32 distinct banks; the 128-bit pass-D pattern 8 distinct 4-bank groups per quarter warp.
  auto check32 = [&](auto addr_of_lane, const char* name) {
    for (int warp = 0; warp < 8; warp++) {
      int seen[32] = {0};
      for (int lane = 0; lane < 32; lane++) { int b = addr_of_lane(warp * 32 + lane) & 31; if (seen[b]++) { printf("bank conflict in %s\n", name); bad++; return; } }
    }
  };
  for (int a = 0; a < 8; a++) {
    check32([&](int tid) { return ntt_phys(a * 256 + tid); }, "pass A");
    check32([&](int tid) { return ntt_phys((tid >> 5) * 256 + a * 32 + (tid & 31)); }, "pass B");
    check32([&](int tid) { return ntt_phys((tid >> 2) * 32 + a * 4 + (tid & 3)); }, "pass C");
  }
  for (int quarter = 0; quarter < 32; quarter++) {
    int seen[8] = {0};
    for (int l = 0; l < 8; l++) {
      int w = ntt_phys((quarter * 8 + l) * 8);
      if (w % 4) { printf("pass D misaligned\n"); bad++; }
      int grp = (w & 31) >> 2;
      if (seen[grp]++) { printf("bank conflict in pass D\n"); bad++; }
    }
  }
  printf(bad ? "FAIL %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
