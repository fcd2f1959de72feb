"""Freeze the oracle's end-to-end outputs for the deterministic cases of tests/golden_cases.py.

    python tests/golden/make_golden.py        # rewrites tests/golden/spiral_golden.json

The reference (Rust) cannot run in this image and stores no ciphertext-level vectors of its own (its tests encrypt with
fresh entropy and compare decryptions), so these fixtures are produced by the oracle AFTER it has been pinned against the
reference's known-answer tests (tests/test_oracle_kats.py).  They guard against drift: the oracle must keep reproducing
them (tests/test_oracle_protocol.py) and the CUDA path must produce the same response bytes (tests/test_gpu_parity.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

import golden_cases as GC  # noqa: E402


def main():
    out = {"generator": "tests/golden/make_golden.py (oracle = C++ restatement of blyssprivacy/sdk spiral-rs)",
           "seed_client": GC.GOLDEN_SEED_CLIENT, "seed_db": GC.GOLDEN_SEED_DB,
           "cases": {c: GC.oracle_record(c) for c in GC.GOLDEN_CASES}}
    path = os.path.join(HERE, "spiral_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
