"""Where does the time of coalesced batches go?  S8: multi-client batches through b200pir_process_queries vs one-client batches,
then 32 threads x 4 requests through b200pir_process_query."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdk_b200.spiral as S

S8 = dict(n=2, nu_1=9, nu_2=8, p=256, q2_bits=22, t_gsw=8, t_conv=4, t_exp_left=8, t_exp_right=8, instances=1, db_item_size=8192,
          version=0)
G = S.Params(**S8)
gdb = S.Database(G); gdb.fill_synthetic(0xB1755)
rng = np.random.default_rng(1)
def rnd(n):
    a = np.empty((n, 2, 2048), dtype=np.uint64)
    a[:, 0] = rng.integers(0, 268369921, (n, 2048), dtype=np.uint64); a[:, 1] = rng.integers(0, 249561089, (n, 2048), dtype=np.uint64)
    return a.reshape(-1)
W = G.words
def mkpp():
    return S.PublicParameters(G, rnd(W["pack"] // 4096), rnd(W["left"] // 4096), rnd(W["right"] // 4096), rnd(W["conv"] // 4096))
ppa, ppb = mkpp(), mkpp()
mod = 268369921 * 249561089
qs = [rng.integers(0, mod, 4096, dtype=np.uint64) for _ in range(32)]
def t(fn, reps=5):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
for n in (1, 8, 16, 32):
    one = t(lambda: S.process_query_batch(G, ppa, np.concatenate(qs[:n]), gdb))
    multi = t(lambda: S.process_queries(G, [ppa if k % 2 == 0 else ppb for k in range(n)], qs[:n], gdb))
    same = t(lambda: S.process_queries(G, [ppa] * n, qs[:n], gdb))
    print("n=%2d  one-client batch %.2f ms   process_queries same client %.2f ms   two clients %.2f ms" % (n, one, same, multi), flush=True)
ser = t(lambda: S.process_query(G, ppa, S.Query(ct=qs[0]), gdb), 10)
print("single process_query %.2f ms" % ser)
for window, nthreads in [(w, n) for w in (0, 200, 1000) for n in (2, 8, 32)]:
    G.set_option("coalesce_window_us", window)
    b0, q0 = S.coalesce_stats(G)
    start = threading.Barrier(nthreads)
    def worker(k):
        start.wait()
        for _ in range(4):
            S.process_query(G, ppa if k % 2 == 0 else ppb, S.Query(ct=qs[k]), gdb)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(nthreads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    b1, q1 = S.coalesce_stats(G)
    print("window %4d us  %2d threads x 4: %.1f ms total, %.2f ms per query, %d batches for %d queries" % (window, nthreads, dt * 1e3, dt * 1e3 / (4 * nthreads), b1 - b0, q1 - q0), flush=True)
