"""One process, contexts on two devices (INTEGRATION.md describes a single Rust server process; the header allows one context
per GPU): every kernel that needs more than 48 KiB of dynamic shared memory must have opted in on EACH device (the opt-in is a
per-device attribute), and a context must leave the caller's current device alone.  Skipped on a single-GPU box."""
import threading

import numpy as np
import pytest

from test_gpu_parity import setup_case

pytestmark = [pytest.mark.gpu]


def test_contexts_on_two_devices_in_one_process():
    S, P, cl, pp, db, G0, gdb0, gpp0 = setup_case("T")
    from sdk_b200._lib import LIB
    if LIB.b200pir_device_count() < 2:
        pytest.skip("needs two GPUs")
    q = cl.generate_query(5)
    ref = P.process_query(pp, q, db)
    G1 = S.Params(device=1, **P.kw)
    handles = []
    try:
        for fmt in (2, 1, 0):                                   # every first-dimension layout on the second device
            gdb1 = S.Database.from_words(G1, db, fmt=fmt)
            handles.append(gdb1)
            gpp1 = S.PublicParameters(G1, pp["pack"], pp.get("left"), pp.get("right"), pp.get("conv"))
            handles.append(gpp1)
            assert np.array_equal(S.process_query(G1, gpp1, S.Query(ct=q["ct"]), gdb1), ref), fmt
            assert np.array_equal(S.process_query(G0, gpp0, S.Query(ct=q["ct"]), gdb0), ref), fmt      # device 0 still fine
        # both devices at once from two host threads
        out = {}

        def run(tag, G, gpp, gdb):
            out[tag] = [S.process_query(G, gpp, S.Query(ct=q["ct"]), gdb).copy() for _ in range(8)]

        t = [threading.Thread(target=run, args=(0, G0, gpp0, gdb0)), threading.Thread(target=run, args=(1, G1, handles[-1], handles[-2]))]
        for x in t:
            x.start()
        for x in t:
            x.join()
        assert all(np.array_equal(o, ref) for tag in (0, 1) for o in out[tag]) and len(out[0]) == len(out[1]) == 8
    finally:
        for h in reversed(handles):
            h.close()
        G1.close()
