"""Deterministic end-to-end cases shared by tests/golden/make_golden.py (which freezes the oracle's outputs as fixtures)
and the tests that replay them (oracle: test_oracle_protocol.py, CUDA path: test_gpu_parity.py)."""
import hashlib

import numpy as np

import oracle_lib as O

GOLDEN_SEED_CLIENT = 20260923
GOLDEN_SEED_DB = 0x60D1
GOLDEN_CASES = {          # name -> (parameter set, expand_queries, item indices)
    "T_expand": ("T", True, [0, 77, 255]),
    "T1_expand": ("T1", True, [1, 100]),
    "T0_expand": ("T0", True, [2, 31]),
    "T_direct": ("T", False, [42]),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_case(case):
    """-> (P, client, pp dict, db words, [(idx, query dict)]) — same RNG call order every time."""
    name, expand, idxs = GOLDEN_CASES[case]
    P = O.Params.named(name, expand_queries=expand)
    cl = O.Client(P, GOLDEN_SEED_CLIENT)
    pp = cl.generate_keys()
    db = P.generate_db(GOLDEN_SEED_DB)
    total = P.dim0 * P.num_per
    queries = [(i % total, cl.generate_query(i % total)) for i in idxs]
    return P, cl, pp, db, queries


def oracle_record(case):
    P, cl, pp, db, queries = build_case(case)
    rec = {"params": P.kw, "expand_queries": bool(GOLDEN_CASES[case][1]), "db_sha256": sha(db),
           "pp_sha256": {k: sha(v) for k, v in pp.items() if v is not None}, "queries": []}
    for idx, q in queries:
        resp = P.process_query(pp, q, db)
        item = cl.decode_response(resp)
        assert np.array_equal(item, P.db_plain_item(GOLDEN_SEED_DB, idx))
        rec["queries"].append({"idx": int(idx), "query_sha256": {k: sha(v) for k, v in q.items() if v is not None},
                               "response_sha256": sha(resp), "response_head": [int(x) for x in resp[:16]],
                               "item_sha256": sha(item)})
    return rec
