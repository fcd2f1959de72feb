// Compile-and-run check of the C++ host mirror (include/b200pir.hpp): a tiny process_query on cuda:0 with
// synthetic public parameters; prints a checksum of the response.  tests/test_gpu_parity.py compares the
// checksum with the Python path on the same inputs.
#include "../../include/b200pir.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  b200pir_params p{2, 6, 2, 256, 20, 8, 4, 8, 8, 1, 8192, 0, 1};
  try {
    spiral_rs::Params params(p, 0);
    // deterministic pseudo-random residues (same generator as the Python side of the test)
    uint64_t s = 88172645463325252ULL;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    const uint64_t q0 = 268369921ULL, q1 = 249561089ULL;
    auto ntt_mat = [&](size_t polys) {
      std::vector<uint64_t> v(polys * 4096);
      for (size_t i = 0; i < polys; i++) { for (int z = 0; z < 2048; z++) v[i * 4096 + z] = next() % q0; for (int z = 0; z < 2048; z++) v[i * 4096 + 2048 + z] = next() % q1; }
      return v;
    };
    auto pack = ntt_mat(2 * 3 * 4), left = ntt_mat(7 * 2 * 8), right = ntt_mat(5 * 2 * 8), conv = ntt_mat(2 * 8);
    spiral_rs::PublicParameters pp(params, pack, &left, &right, &conv);
    spiral_rs::Database db(params);
    spiral_rs::check(b200pir_db_fill_synthetic(params.ctx, db.h, 0xB1755));
    spiral_rs::Query q;
    q.ct.resize(4096);
    for (auto& x : q.ct) x = next() % (q0 * q1);
    auto resp = spiral_rs::server::process_query(params, pp, q, db);
    uint64_t h = 1469598103934665603ULL;
    for (uint8_t b : resp) { h ^= b; h *= 1099511628211ULL; }
    printf("%zu %llu\n", resp.size(), (unsigned long long)h);
  } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
