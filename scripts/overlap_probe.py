"""Experiment: two contexts (two CUDA streams) sharing one database, steps alternate between them.
Does the HBM-bound first dimension of one batch overlap the ALU-bound expansion/fold of the other?"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sdk_b200.spiral as S
from sdk_b200._lib import LIB, check
import bench as Bn
kw = dict(Bn.S8)
fmt = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nctx = int(sys.argv[3]) if len(sys.argv) > 3 else 2
steps = 24
ctxs = [S.Params(**kw) for _ in range(nctx)]
streams = [torch.cuda.Stream() for _ in range(nctx)]
for c, st in zip(ctxs, streams):
    c.set_stream(st.cuda_stream)
    c.set_option("batch", 4)
gdb = S.Database(ctxs[0], fmt=fmt)
gdb.fill_synthetic(0xB1755)
rng = np.random.default_rng(1)
pp = Bn.synthetic_pp(kw, rng)
gpp = S.PublicParameters(ctxs[0], pp["pack"], pp["left"], pp["right"], pp["conv"])
rb = ctxs[0].response_bytes
qs = [torch.from_numpy(rng.integers(0, Bn.Q0 * Bn.Q1, B * 4096, dtype=np.uint64).view(np.int64)).cuda() for _ in range(nctx)]
outs = [torch.zeros(B * rb, dtype=torch.uint8, device="cuda") for _ in range(nctx)]
def run(n):
    for k in range(n):
        i = k % nctx
        check(LIB.b200pir_process_query_batch_dev(ctxs[i]._h, gdb._h, gpp._h, qs[i].data_ptr(), B, outs[i].data_ptr()))
run(2 * nctx)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"db_format": fmt, "batch": B, "contexts": nctx, "qps": B * steps / dt, "ms_per_batch": dt / steps * 1e3}))
