// 4096-point negacyclic NTT over one 28-bit CRT modulus, computed cooperatively by 512 threads holding 8 coefficients
// each: the BASELINE.json config #5 sweep ("poly_len in {2048, 4096}").  The reference hard-codes poly_len = 2048
// (lib/spiral-rs/src/util.rs:246), so nothing on the query path uses this size; the transforms follow the same scalar
// definition (ntt.rs:67-113, :212-258, generic in poly_len_log2) with tables built the same way (ntt.rs:39-65), and both
// moduli are 1 mod 8192.  Same butterflies as ntt_core.cuh; twelve stages = four radix-8 passes:
//
//   stage bits   11 10 9 | 8 7 6 | 5 4 3 | 2 1 0
//   pass             A       B       C       D
//   thread owns  e = a*512 + tid               (A: a = bits 11..9)        "strided layout"
//                e = hi*512 + a*64 + lo        (B: tid = hi*64 + lo)
//                e = H*64 + a*8 + l3           (C: tid = H*8 + l3)
//                e = tid*8 + k                 (D: k = bits 2..0)         "contiguous layout"
//
// Shared-memory exchanges use the padding of ntt_core.cuh, phys(e) = e + 4*(e>>5) (4608 words per transform): pass A/B
// accesses are warp-contiguous, pass C's four H values of a warp land 8 banks apart, pass D is the 128-bit pattern already
// used at 2048.  Everything is __host__ __device__ and emulated on the CPU (tests/cpp/ntt_core4096_emul.cpp).
#pragma once
#include "ntt_core.cuh"

namespace b200pir {

static const int NTT4K_N = 4096;
static const int NTT4K_LOG_N = 12;
static const int NTT4K_THREADS = 512;
static const int NTT4K_SMEM_WORDS = 4096 + 4 * 128;

template <typename Tab>
NTT_HD void fwd4k_pass_a(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  radix8_fwd(x, tab, 1, 0, q, two_q);                                   // m = 1,2,4
#pragma unroll
  for (int a = 0; a < 8; a++) smem[ntt_phys(a * 512 + tid)] = x[a];
}
template <typename Tab>
NTT_HD void fwd4k_pass_b(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  int hi = tid >> 6, lo = tid & 63;
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = smem[ntt_phys(hi * 512 + a * 64 + lo)];
  radix8_fwd(x, tab, 8, hi, q, two_q);                                  // m = 8,16,32
#pragma unroll
  for (int a = 0; a < 8; a++) smem[ntt_phys(hi * 512 + a * 64 + lo)] = x[a];
}
template <typename Tab>
NTT_HD void fwd4k_pass_c(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  int H = tid >> 3, l3 = tid & 7;
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = smem[ntt_phys(H * 64 + a * 8 + l3)];
  radix8_fwd(x, tab, 64, H, q, two_q);                                  // m = 64,128,256
#pragma unroll
  for (int a = 0; a < 8; a++) smem[ntt_phys(H * 64 + a * 8 + l3)] = x[a];
}
template <bool CANON = true, typename Tab>
NTT_HD void fwd4k_pass_d(int tid, uint32_t (&x)[8], const uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  int base = ntt_phys(tid * 8);
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = smem[base + k];
  radix8_fwd(x, tab, 512, tid, q, two_q);                               // m = 512,1024,2048
  if (CANON) {
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = ntt_canon(x[k], q, two_q);
  }
}
template <typename Tab>
NTT_HD void inv4k_pass_d(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  radix8_inv(x, tab, 512, tid, q, two_q);
  int base = ntt_phys(tid * 8);
#pragma unroll
  for (int k = 0; k < 8; k++) smem[base + k] = x[k];
}
template <typename Tab>
NTT_HD void inv4k_pass_c(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  int H = tid >> 3, l3 = tid & 7;
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = smem[ntt_phys(H * 64 + a * 8 + l3)];
  radix8_inv(x, tab, 64, H, q, two_q);
#pragma unroll
  for (int a = 0; a < 8; a++) smem[ntt_phys(H * 64 + a * 8 + l3)] = x[a];
}
template <typename Tab>
NTT_HD void inv4k_pass_b(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
  int hi = tid >> 6, lo = tid & 63;
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = smem[ntt_phys(hi * 512 + a * 64 + lo)];
  radix8_inv(x, tab, 8, hi, q, two_q);
#pragma unroll
  for (int a = 0; a < 8; a++) smem[ntt_phys(hi * 512 + a * 64 + lo)] = x[a];
}
template <typename Tab>
NTT_HD void inv4k_pass_a(int tid, uint32_t (&x)[8], const uint32_t* smem, Tab tab, uint32_t q, uint32_t two_q) {
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = smem[ntt_phys(a * 512 + tid)];
  radix8_inv(x, tab, 1, 0, q, two_q);
#pragma unroll
  for (int a = 0; a < 8; a++) x[a] = ntt_canon(x[a], q, two_q);
}

#if defined(__CUDACC__)
// forward: strided layout in (x[a] = coefficient a*512 + tid), contiguous layout out (x[k] = slot tid*8 + k);
// inverse: the reverse.  gsync() synchronises the 512 threads.
template <bool CANON = true, typename Sync, typename Tab>
__device__ __forceinline__ void ntt4k_forward_group(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, Sync gsync) {
  const uint32_t two_q = 2 * q;
  gsync();
  fwd4k_pass_a(tid, x, smem, tab, q, two_q);
  gsync();
  fwd4k_pass_b(tid, x, smem, tab, q, two_q);
  gsync();
  fwd4k_pass_c(tid, x, smem, tab, q, two_q);
  gsync();
  fwd4k_pass_d<CANON>(tid, x, smem, tab, q, two_q);
}
template <typename Sync, typename Tab>
__device__ __forceinline__ void ntt4k_inverse_group(int tid, uint32_t (&x)[8], uint32_t* smem, Tab tab, uint32_t q, Sync gsync) {
  const uint32_t two_q = 2 * q;
  gsync();
  inv4k_pass_d(tid, x, smem, tab, q, two_q);
  gsync();
  inv4k_pass_c(tid, x, smem, tab, q, two_q);
  gsync();
  inv4k_pass_b(tid, x, smem, tab, q, two_q);
  gsync();
  inv4k_pass_a(tid, x, smem, tab, q, two_q);
}
#endif

}  // namespace b200pir
