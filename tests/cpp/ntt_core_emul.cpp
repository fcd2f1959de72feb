// CPU emulation of the 256-thread cooperative NTT in sdk_b200/csrc/ntt_core.cuh: the per-pass
// thread-local functions are run for tid = 0..255 with an explicit "shared memory" array, pass by
// pass (a pass boundary = the group barrier), and compared with the oracle's scalar transforms.
// This validates every index / twiddle / padding computation without a GPU.
#define NTT_RANGE_CHECK 1
#include "../../sdk_b200/csrc/ntt_core.cuh"
#include "../../sdk_b200/csrc/ntt_tables.hpp"
#include "../../oracle/spiral_oracle.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace b200pir;
namespace b200pir { int ntt_range_violations = 0; }

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
    // the product's own table construction (ntt_tables.hpp) must reproduce the oracle's (= the reference's) tables
    {
      std::vector<Twiddle> f2, i2;
      tables::build_tables(q, f2, i2);
      for (int i = 0; i < 2048; i++)
        if (f2[i].w != fwd[i].w || f2[i].wp != fwd[i].wp || i2[i].w != inv[i].w || i2[i].wp != inv[i].wp) { if (bad < 5) printf("table mismatch mod %d at %d\n", mod, i); bad++; }
    }
    std::vector<Twiddle> inv_lz;
    tables::build_inverse_table_lz(q, inv_lz);
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
      // ---- relaxed-range forward: inputs < 2q (trial 3: < 4q with IN4Q; trial 4: all 2q-1; trial 5: alternating 0 / 2q-1),
      // every output mode; lazy outputs must be congruent and inside their range
      std::vector<uint32_t> in_lz = in;
      if (trial == 4) for (auto& v : in_lz) v = two_q - 1;
      if (trial == 5) for (int i = 0; i < 2048; i++) in_lz[i] = (i & 1) ? two_q - 1 : 0;
      std::vector<uint64_t> ref3(4096, 0);
      for (int i = 0; i < 2048; i++) ref3[mod * 2048 + i] = in_lz[i] % q;
      orc::ntt_forward(p, ref3.data());
      for (int out = 0; out < 3; out++) {
        std::fill(smem.begin(), smem.end(), 0xDEADBEEF);
        for (int t = 0; t < 256; t++) for (int a = 0; a < 8; a++) regs[t][a] = in_lz[a * 256 + t];
        for (int t = 0; t < 256; t++) {
          if (trial == 3) fwd_pass_a_lz<true>(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
          else fwd_pass_a_lz<false>(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
        }
        for (int t = 0; t < 256; t++) fwd_pass_b_lz(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
        for (int t = 0; t < 256; t++) fwd_pass_c_lz(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
        for (int t = 0; t < 256; t++) {
          if (out == 0) fwd_pass_d_lz<NTT_OUT_LAZY16>(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
          else if (out == 1) fwd_pass_d_lz<NTT_OUT_LAZY4>(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
          else fwd_pass_d_lz<NTT_OUT_CANON>(t, regs[t], smem.data(), TwArray{fwd.data()}, q, two_q);
        }
        const uint64_t bound = out == 0 ? 16ull * q : (out == 1 ? 4ull * q : q);
        for (int t = 0; t < 256; t++) for (int k = 0; k < 8; k++) {
          const uint32_t v = regs[t][k];
          if (v >= bound || v % q != ref3[mod * 2048 + t * 8 + k]) { if (bad < 5) printf("lz fwd mismatch mod %d trial %d out %d at %d\n", mod, trial, out, t * 8 + k); bad++; }
        }
      }
      // ---- un-halved inverse: inputs < 2q in the contiguous layout (canonical forward output, + q on odd trials)
      std::vector<uint64_t> ref4 = ref3;
      orc::ntt_inverse(p, ref4.data());
      std::fill(smem.begin(), smem.end(), 0xDEADBEEF);
      for (int t = 0; t < 256; t++) for (int k = 0; k < 8; k++) regs[t][k] = (uint32_t)ref3[mod * 2048 + t * 8 + k] + ((trial & 1) ? q : 0);
      for (int t = 0; t < 256; t++) inv_pass_d_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_c_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_b_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_a_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) for (int a = 0; a < 8; a++)
        if (regs[t][a] != ref4[mod * 2048 + a * 256 + t]) { if (bad < 5) printf("nh inv mismatch mod %d trial %d at %d\n", mod, trial, a * 256 + t); bad++; }
    }
    // worst case for the inverse's ranges: every input 2q-1 (result: N (2q-1) / N at index 0 ... compare mod q)
    {
      static uint32_t regs[256][8];
      std::vector<uint32_t> smem(NTT_SMEM_WORDS, 0);
      std::vector<uint64_t> ref5(4096, 0);
      for (int i = 0; i < 2048; i++) ref5[mod * 2048 + i] = (two_q - 1) % q;
      orc::ntt_inverse(p, ref5.data());
      for (int t = 0; t < 256; t++) for (int k = 0; k < 8; k++) regs[t][k] = two_q - 1;
      for (int t = 0; t < 256; t++) inv_pass_d_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_c_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_b_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) inv_pass_a_nh(t, regs[t], smem.data(), TwArray{inv_lz.data()}, q, two_q);
      for (int t = 0; t < 256; t++) for (int a = 0; a < 8; a++)
        if (regs[t][a] != ref5[mod * 2048 + a * 256 + t]) { if (bad < 5) printf("nh inv (max input) mismatch mod %d at %d\n", mod, a * 256 + t); bad++; }
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
  if (ntt_range_violations) { printf("range violations: %d\n", ntt_range_violations); bad += ntt_range_violations; }
  printf(bad ? "FAIL %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
