"""N>1 host logic on CPU (gloo, world_size 2 and 4): row sharding ii = s (mod G), local fold rounds, ONE
all-gather of the survivors, final log2(G) rounds — composed exactly as bench.py / b200pir_query_stage_{a,b}_dev do
on GPUs — equals the single-node fold byte for byte.  The oracle stands in for the kernels here; the GPU twin of
this test is tests/test_gpu_parity.py::test_sharded_stages_equal_single_gpu."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O.LIB.orc_set_num_threads(1)
    P = O.Params.named("T0")                       # nu_2 = 3 -> 8 second-dimension rows, 9 slices
    cl = O.Client(P, 4321)
    pp = cl.generate_keys()
    db = P.generate_db(0xB1755)
    q = cl.generate_query(57)
    _, d = P.process_query(pp, q, db, dump=True)   # every rank derives the same expansion (replicated, as on GPUs)
    N, logw = P.N, world.bit_length() - 1
    mat = 2 * 2 * P.t_gsw * P.W
    slice_words = P.dim0 * P.num_per * N
    partial = np.zeros((P.slices, 2 * N), dtype=np.uint64)
    for s in range(P.slices):
        # this rank's rows of the first-dimension product: global rows ii = il*world + rank
        full = P.multiply_reg_by_database(db[s * slice_words:(s + 1) * slice_words], d["v_firstdim"]).reshape(P.num_per, 4 * N)
        local = np.ascontiguousarray(full[rank::world])
        raw = P.from_ntt(local.reshape(-1))
        # local rounds use v_folding[nu_2-1 .. logw]  (stage A)
        folded = P.fold_ciphertexts(raw, d["v_folding"][logw * mat:], d["v_folding_neg"][logw * mat:]) if local.shape[0] > 1 else raw
        partial[s] = folded[: 2 * N]
    t = torch.from_numpy(partial.view(np.int64))
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)                   # the single collective
    g = np.stack([x.numpy().view(np.uint64) for x in gathered])      # [world][slices][ct]
    final = np.zeros((P.slices, 2 * N), dtype=np.uint64)
    for s in range(P.slices):
        cts = np.ascontiguousarray(g[:, s, :]).reshape(-1)             # survivor of rank s' = global ct[s']
        final[s] = P.fold_ciphertexts(cts, d["v_folding"], d["v_folding_neg"])[: 2 * N] if world > 1 else cts[: 2 * N]
    ok = np.array_equal(final.reshape(-1), d["folded"])
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_fold_equals_single_node(world, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r))[0]), "rank %d disagrees" % r
