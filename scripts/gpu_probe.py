"""Quick GPU probe: full-size S8 process_query with per-stage device times, mul-variant sweep."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
import sdk_b200.spiral as S

name = sys.argv[1] if len(sys.argv) > 1 else "S8"
P = O.Params.named(name)
t0 = time.time()
cl = O.Client(P, 5)
pp = cl.generate_keys()
print("keygen %.1fs" % (time.time() - t0), flush=True)
G = S.Params(**P.kw)
gdb = S.Database(G)
t0 = time.time()
gdb.fill_synthetic(0xB1755)
print("db synth %.2fs  (%.2f GiB)" % (time.time() - t0, P.slices * P.dim0 * P.num_per * 2048 * 8 / 2**30), flush=True)
gpp = S.PublicParameters(G, pp["pack"], pp.get("left"), pp.get("right"), pp.get("conv"))
G.set_option("profile", 1)
idx = 12345 % (P.dim0 * P.num_per)
q = cl.generate_query(idx)
db_bytes = P.slices * P.dim0 * P.num_per * 2048 * 8
for variant in (0, 1, 2, 3):
    G.set_option("mul_variant", variant)
    for rep in range(3):
        t0 = time.time()
        resp = S.process_query(G, gpp, S.Query(ct=q["ct"]), gdb)
        wall = time.time() - t0
    ms = G.last_stage_ms()
    print("variant", variant, "wall %.2f ms" % (wall * 1e3), json.dumps({k: round(v, 3) for k, v in ms.items()}),
          "mul GB/s %.0f" % (db_bytes / ms["multiply"] / 1e6), flush=True)
ok = np.array_equal(cl.decode_response(resp), P.db_plain_item(0xB1755, idx))
print("decode ok:", ok, flush=True)
G.set_option("mul_variant", 0)
for batch in (2, 4):
    G.set_option("batch", batch)
    qs = np.concatenate([cl.generate_query((idx + 7 * k) % (P.dim0 * P.num_per))["ct"] for k in range(batch)])
    for rep in range(2):
        t0 = time.time()
        out = S.process_query_batch(G, gpp, qs, gdb)
        wall = time.time() - t0
    ms = G.last_stage_ms()
    okb = all(np.array_equal(cl.decode_response(out[k]), P.db_plain_item(0xB1755, (idx + 7 * k) % (P.dim0 * P.num_per))) for k in range(batch))
    print("batch", batch, "wall %.2f ms" % (wall * 1e3), json.dumps({k: round(v, 3) for k, v in ms.items()}), "decode ok:", okb, flush=True)
