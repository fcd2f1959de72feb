// Wire formats of the query path: the seed-expanded first rows of serialized PublicParameters / Query
// (lib/spiral-rs/src/client.rs:47-80, 212-259, 303-315).  The client sends a 32-byte seed instead of the uniformly
// random first row of every ciphertext matrix; the server regenerates it as q - (next_u64 % q) from
// ChaCha20Rng::from_seed(seed).  The keystream is counter mode, so every 64-byte block is independent: one thread per
// block, eight u64 per thread, written straight into the raw matrices in HBM.
#include "common.cuh"
#include "kernels.h"

namespace b200pir {

struct ChaChaKey { uint32_t k[8]; };

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

#define B200_QR(a, b, c, d)                                   \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

// raw: n_mats matrices, mat_words apart; the first row_words u64 of each are filled from keystream position
// (block0 * 8 + i * row_words + j).  row_words is a multiple of 2048, so blocks never straddle matrices.
__global__ void __launch_bounds__(256) k_chacha_first_rows(uint64_t* __restrict__ raw, ChaChaKey key, uint64_t block0,
                                                           uint32_t n_mats, uint32_t row_words, uint64_t mat_words,
                                                           uint64_t modulus) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (uint64_t)n_mats * row_words / 8;
  if (b >= total) return;
  const uint64_t ctr = block0 + b;
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                     key.k[4], key.k[5], key.k[6], key.k[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7], x8 = in[8],
           x9 = in[9], x10 = in[10], x11 = in[11], x12 = in[12], x13 = in[13], x14 = in[14], x15 = in[15];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    B200_QR(x0, x4, x8, x12) B200_QR(x1, x5, x9, x13) B200_QR(x2, x6, x10, x14) B200_QR(x3, x7, x11, x15)
    B200_QR(x0, x5, x10, x15) B200_QR(x1, x6, x11, x12) B200_QR(x2, x7, x8, x13) B200_QR(x3, x4, x9, x14)
  }
  const uint32_t o[16] = {x0 + in[0], x1 + in[1], x2 + in[2], x3 + in[3], x4 + in[4], x5 + in[5], x6 + in[6], x7 + in[7],
                          x8 + in[8], x9 + in[9], x10 + in[10], x11 + in[11], x12 + in[12], x13 + in[13], x14 + in[14],
                          x15 + in[15]};
  const uint64_t k = b * 8;
  const uint64_t mat = k / row_words, j = k % row_words;
  uint64_t* dst = raw + mat * mat_words + j;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint64_t v = (uint64_t)o[2 * i] | ((uint64_t)o[2 * i + 1] << 32);      // BlockRng::next_u64: low word first
    dst[i] = modulus - (v % modulus);                                              // client.rs:47-49 (q - 0 stays q)
  }
}

void launch_chacha_first_rows(uint64_t* raw, const uint8_t seed[32], uint64_t word0, uint32_t n_mats, uint32_t row_words,
                              uint64_t mat_words, uint64_t modulus, cudaStream_t s) {
  if (word0 % 8 || row_words % 8) throw Error(-1, "keystream segments must be 64-byte aligned");
  ChaChaKey key;
  for (int i = 0; i < 8; i++)
    key.k[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
  const uint64_t blocks = (uint64_t)n_mats * row_words / 8;
  if (!blocks) return;
  k_chacha_first_rows<<<(unsigned)((blocks + 255) / 256), 256, 0, s>>>(raw, key, word0 / 8, n_mats, row_words, mat_words, modulus);
  g_kernel_launches++;
  B200_CUDA(cudaGetLastError());
}

// Query::deserialize, direct-upload branch (client.rs:316-327 with interleave_rng_data :107-131): the device-format first
// dimension operand q_dev[j][z] = {row 0 mod q0, row 0 mod q1 (regenerated from the seed, transformed), low and high half of
// the uploaded word (z, j)} — the reference's interleaved v_buf[(z dim0 + j) 2 + {0, 1}] without materialising it.
// sig: ntt32 [j][n][z] of the regenerated row-0 polynomials; wire: the uploaded odd-indexed words, [z][j].
__global__ void k_direct_query_to_dev(uint4* __restrict__ q_dev, const uint32_t* __restrict__ sig, const uint64_t* __restrict__ wire,
                                      int dim0) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over dim0 * 2048, z fastest
  if (idx >= (size_t)dim0 * 2048) return;
  const int z = (int)(idx % 2048), j = (int)(idx / 2048);
  const uint64_t w = wire[(size_t)z * dim0 + j];
  q_dev[idx] = make_uint4(sig[((size_t)j * 2 + 0) * 2048 + z], sig[((size_t)j * 2 + 1) * 2048 + z], (uint32_t)w, (uint32_t)(w >> 32));
}
void launch_direct_query_to_dev(uint4* q_dev, const uint32_t* sig_ntt, const uint64_t* wire_words, int dim0, cudaStream_t s) {
  const size_t total = (size_t)dim0 * 2048;
  k_direct_query_to_dev<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(q_dev, sig_ntt, wire_words, dim0);
  g_kernel_launches++;
}

}  // namespace b200pir
