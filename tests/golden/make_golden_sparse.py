"""tests/golden/make_golden_sparse.py -- regenerates tests/golden/sparse_vectors.json.

ZipLinearCode (zip/code.rs:77-215) commits produced by oracle/pyoracle.py: the matrices are sampled by the restated
transcripts (KeccakTranscript transcript.rs:161-201, whose Keccak-256 is pinned to hashlib in test_oracle_pins.py, and
the reference tests' MockTranscript pcs/tests.rs:24-56), the products are exact Python ints, every digest comes from
the `blake3` PyPI package (a binding of the Rust crate the reference depends on).  The reference ships no known-answer
vectors for this code either (its tests check success / determinism only, commit.rs:218-300).

    python tests/golden/make_golden_sparse.py
"""
import json
import os
import random
import sys

import blake3

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

I64_MAX, I64_MIN = (1 << 63) - 1, -(1 << 63)


def make(nv, transcript_name, pat):
    n = 1 << nv
    t = po.MockTranscript() if transcript_name == "mock" else po.KeccakTranscript()
    row_len, cw, a, b = po.zip_linear_code_new(n, t)
    num_rows = po.num_rows_for(n, row_len)
    rnd = random.Random(0x5A + nv)
    evals = {"one_to_n": list(range(1, n + 1)),  # commit.rs:234
             "const42": [42] * n,                # commit.rs:281
             "random": [rnd.randrange(I64_MIN, I64_MAX + 1) for _ in range(n)]}[pat]
    K, half, d = 4, cw // 2, row_len // 2
    rows = []
    for r in range(num_rows):
        rows += po.sparse_encode_row(evals[r * row_len:(r + 1) * row_len], half, d, a, b)
    depth = (cw - 1).bit_length() if cw > 1 else 0
    roots, layers_h = [], blake3.blake3()
    for r in range(num_rows):
        root, layers = po.merkle_tree(depth, rows[r * cw:(r + 1) * cw], K)
        for d_ in layers:
            layers_h.update(d_)
        roots.append(root.hex())
    rows_bytes = b"".join(w.to_bytes(8, "little") for v in rows for w in po.to_words(v, K))
    return {"nv": nv, "transcript": transcript_name, "pattern": pat, "row_len": row_len, "cw": cw, "num_rows": num_rows,
            "cells_per_row": d, "cols_a": a[0], "coef_a": a[1], "cols_b": b[0], "coef_b": b[1],
            "evals": [str(v) for v in evals], "rows_blake3": blake3.blake3(rows_bytes).hexdigest(),
            "rows_head": [str(v) for v in rows[:8]], "layers_blake3": layers_h.hexdigest(), "roots": roots}


def main():
    out = []
    for nv, pat in ((2, "one_to_n"), (3, "one_to_n"), (4, "const42"), (5, "random"), (6, "random")):
        out.append(make(nv, "mock", pat))
    for nv, pat in ((3, "one_to_n"), (4, "random"), (6, "random"), (8, "random")):
        out.append(make(nv, "keccak", pat))
    with open(os.path.join(HERE, "sparse_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "sparse vectors")


if __name__ == "__main__":
    main()
