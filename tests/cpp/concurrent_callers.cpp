// Concurrent callers of one context from native threads, the way lib/server's actix workers call process_query under a
// read lock (bin/server.rs:98-141): every thread hands a serialized query to b200pir_process_query_bytes (count = 1) and
// gets the response bytes back.  No interpreter lock is involved, unlike the Python-thread version of this test.
//
//   concurrent_callers <dir> <threads> <requests per thread>
// <dir> holds what tests/test_gpu_zz_concurrent_callers.py wrote with the oracle client: params.txt (the twelve Params
// scalars), pp0.bin / pp1.bin (serialized PublicParameters of two clients), queries.bin (threads x query_bytes; thread k
// belongs to client k % 2).  Prints one line: serial_s concurrent_s passes queries mismatches, then writes responses.bin
// (the serial responses, threads x response_bytes) for the caller to decode.
#include "../../include/b200pir.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

static void ok(int rc, const char* what) {
  if (rc != 0) { fprintf(stderr, "%s: rc %d: %s\n", what, rc, b200pir_last_error()); exit(1); }
}
static std::vector<uint8_t> slurp(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(1); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> v((size_t)n);
  if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read %s\n", path.c_str()); exit(1); }
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s dir threads requests\n", argv[0]); return 2; }
  const std::string dir = argv[1];
  const int threads = atoi(argv[2]), per = atoi(argv[3]);
  b200pir_params p;
  memset(&p, 0, sizeof p);
  {
    FILE* f = fopen((dir + "/params.txt").c_str(), "r");
    if (!f) { fprintf(stderr, "no params.txt\n"); return 1; }
    unsigned long long v[12];
    for (int i = 0; i < 12; i++) if (fscanf(f, "%llu", &v[i]) != 1) { fprintf(stderr, "bad params.txt\n"); return 1; }
    fclose(f);
    p.n = v[0]; p.nu_1 = v[1]; p.nu_2 = v[2]; p.p = v[3]; p.q2_bits = v[4]; p.t_gsw = v[5]; p.t_conv = v[6];
    p.t_exp_left = v[7]; p.t_exp_right = v[8]; p.instances = v[9]; p.db_item_size = v[10]; p.version = v[11];
    p.expand_queries = 1;
  }
  b200pir_ctx* ctx = nullptr;
  ok(b200pir_ctx_create(&p, 0, &ctx), "ctx_create");
  uint64_t setup_bytes = 0, query_bytes = 0, response_bytes = 0;
  ok(b200pir_ctx_sizes(ctx, &setup_bytes, &query_bytes, &response_bytes), "ctx_sizes");
  ok(b200pir_ctx_reserve(ctx, 32, (size_t)1 << p.nu_2), "ctx_reserve");
  b200pir_db* db = nullptr;
  ok(b200pir_db_create(ctx, 0, 1, &db), "db_create");
  ok(b200pir_db_fill_synthetic(ctx, db, 0xB1755), "db_fill_synthetic");
  b200pir_pp* pp[2] = {nullptr, nullptr};
  for (int c = 0; c < 2; c++) {
    auto bytes = slurp(dir + "/pp" + std::to_string(c) + ".bin");
    if (bytes.size() != setup_bytes) { fprintf(stderr, "pp%d.bin: %zu bytes, expected %llu\n", c, bytes.size(), (unsigned long long)setup_bytes); return 1; }
    ok(b200pir_pp_create_from_bytes(ctx, bytes.data(), bytes.size(), &pp[c]), "pp_create_from_bytes");
  }
  auto queries = slurp(dir + "/queries.bin");
  if (queries.size() != (size_t)threads * query_bytes) { fprintf(stderr, "queries.bin: wrong size\n"); return 1; }

  auto call = [&](int k, uint8_t* out) {
    size_t n = 0;
    return b200pir_process_query_bytes(ctx, db, pp[k % 2], queries.data() + (size_t)k * query_bytes, query_bytes, 1, out, &n);
  };
  // serial reference (also warms everything up)
  std::vector<uint8_t> serial((size_t)threads * response_bytes);
  for (int k = 0; k < threads; k++) ok(call(k, serial.data() + (size_t)k * response_bytes), "serial process_query_bytes");
  std::vector<uint8_t> scratch(response_bytes);
  auto t0 = std::chrono::steady_clock::now();
  for (int j = 0; j < per; j++)
    for (int k = 0; k < threads; k++) ok(call(k, scratch.data()), "serial process_query_bytes");
  const double serial_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // concurrent: best of three rounds
  double best = 1e30;
  uint64_t best_passes = 0, total_queries = 0;
  std::atomic<int> mismatches{0}, failures{0};
  for (int round = 0; round < 3; round++) {
    uint64_t b0 = 0, q0 = 0, b1 = 0, q1 = 0;
    ok(b200pir_coalesce_stats(ctx, &b0, &q0), "coalesce_stats");
    std::mutex mu;
    std::condition_variable cv;
    int ready = 0;
    bool go = false;
    std::vector<std::thread> ts;
    for (int k = 0; k < threads; k++)
      ts.emplace_back([&, k] {
        std::vector<uint8_t> out(response_bytes);
        {
          std::unique_lock<std::mutex> lk(mu);
          ready++;
          cv.notify_all();
          cv.wait(lk, [&] { return go; });
        }
        for (int j = 0; j < per; j++) {
          if (call(k, out.data()) != 0) { failures++; continue; }
          if (memcmp(out.data(), serial.data() + (size_t)k * response_bytes, response_bytes) != 0) mismatches++;
        }
      });
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return ready == threads; });
      go = true;
      t0 = std::chrono::steady_clock::now();
      cv.notify_all();
    }
    for (auto& t : ts) t.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ok(b200pir_coalesce_stats(ctx, &b1, &q1), "coalesce_stats");
    total_queries = q1 - q0;
    if (s < best) { best = s; best_passes = b1 - b0; }
  }
  printf("%.6f %.6f %llu %llu %d\n", serial_s, best, (unsigned long long)best_passes, (unsigned long long)total_queries,
         mismatches.load() + failures.load());
  FILE* f = fopen((dir + "/responses.bin").c_str(), "wb");
  if (f) { fwrite(serial.data(), 1, serial.size(), f); fclose(f); }
  b200pir_pp_destroy(pp[0]);
  b200pir_pp_destroy(pp[1]);
  b200pir_db_destroy(db);
  b200pir_ctx_destroy(ctx);
  return 0;
}
