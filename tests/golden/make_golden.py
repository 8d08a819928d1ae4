"""tests/golden/make_golden.py -- regenerates tests/golden/*.json.

The fixtures are produced by oracle/pyoracle.py: arbitrary-precision Python ints for the RAA code and the
`blake3` PyPI package (a binding of the Rust `blake3` crate the reference depends on, Cargo.toml:30) for every
digest.  The reference itself ships no known-answer vectors for this path and cannot be built
here (no Rust toolchain), so these vectors pin the C oracle and the CUDA path to the real hash crate and to an
independent big-int restatement of code_raa.rs.  The permutations are part of each fixture (they come from the
rand-0.9.2 restatement, whose parity with the real crate is unpinned).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import blake3

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

I64_MAX, I64_MIN = (1 << 63) - 1, -(1 << 63)


def pattern(name, n, nv):
    import random

    rnd = random.Random(0x21C0 + nv)
    return {
        "random": [rnd.randrange(I64_MIN, I64_MAX + 1) for _ in range(n)],
        "one_to_n": list(range(1, n + 1)),          # commit.rs:234
        "zeros": [0] * n,                            # commit.rs:475
        "i64_max": [I64_MAX] * n,                    # commit.rs:620-621
        "i64_min": [I64_MIN] * n,
        "alternating": [1 if i % 2 == 0 else -1 for i in range(n)],  # commit.rs:487-489
    }[name]


def hexint(v, limbs):
    return b"".join(w.to_bytes(8, "little") for w in po.to_words(v, limbs)).hex()


def make(nv, seeds, pat, full):
    n = 1 << nv
    row_len = po.raa_row_len(n)
    num_rows = po.num_rows_for(n, row_len)
    rep, K = 2, 4
    cw = row_len * rep
    perm1, perm2 = po.perm_from_seed(cw, seeds[0]), po.perm_from_seed(cw, seeds[1])
    evals = pattern(pat, n, nv)
    rows, layers, roots = po.commit(evals, num_rows, row_len, rep, perm1, perm2, K)
    rows_bytes = b"".join(bytes.fromhex(hexint(v, K)) for v in rows)
    layers_bytes = b"".join(b"".join(l) for l in layers)
    fx = {
        "nv": nv, "seeds": list(seeds), "pattern": pat, "row_len": row_len, "num_rows": num_rows, "cw": cw,
        "in_limbs": 1, "out_limbs": K, "rep": rep,
        "perm1_blake3": blake3.blake3(b"".join(p.to_bytes(4, "little") for p in perm1)).hexdigest(),
        "perm2_blake3": blake3.blake3(b"".join(p.to_bytes(4, "little") for p in perm2)).hexdigest(),
        "rows_blake3": blake3.blake3(rows_bytes).hexdigest(),
        "layers_blake3": blake3.blake3(layers_bytes).hexdigest(),
        "roots": [r.hex() for r in roots],
    }
    if full:
        fx.update({"evals": [str(v) for v in evals], "perm1": perm1, "perm2": perm2,
                   "rows": [hexint(v, K) for v in rows],
                   "layers": [[d.hex() for d in l] for l in layers]})
    return fx


def main():
    out = []
    for seeds in ((1, 2), (0xB9736F582676E7E8, 0xD7397E6260CE9C3E)):
        for nv in (2, 3, 4, 5, 6):
            for pat in ("random", "one_to_n", "zeros", "i64_max", "i64_min", "alternating"):
                out.append(make(nv, seeds, pat, full=nv <= 4))
        for nv in (8, 10):
            out.append(make(nv, seeds, "random", full=False))
    with open(os.path.join(HERE, "commit_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    # hashing KATs: leaf = blake3(Int<K>.to_bytes()), node = blake3(l || r)
    kats = {"leaf": [], "node": [], "blake3": []}
    for limbs in (1, 2, 3, 4, 8):
        for v in (0, 1, -1, I64_MAX, I64_MIN, (1 << 62) + 12345, -(1 << 61) - 999):
            kats["leaf"].append({"limbs": limbs, "value": str(v), "bytes": po.int_to_bytes(v, limbs).hex(),
                                 "digest": blake3.blake3(po.int_to_bytes(v, limbs)).hexdigest()})
    z = blake3.blake3(po.int_to_bytes(0, 4)).digest()
    o = blake3.blake3(po.int_to_bytes(1, 4)).digest()
    for l, r in ((z, z), (z, o), (o, z)):
        kats["node"].append({"left": l.hex(), "right": r.hex(), "digest": blake3.blake3(l + r).hexdigest()})
    for n in (0, 1, 32, 63, 64, 65, 128, 1024, 1025, 2049):
        data = bytes((i * 7 + 3) % 251 for i in range(n))
        kats["blake3"].append({"len": n, "digest": blake3.blake3(data).hexdigest()})
    with open(os.path.join(HERE, "hash_kats.json"), "w") as f:
        json.dump(kats, f, indent=0)
    # rand restatement snapshot (NOT a pin against the real crate: guards against accidental drift only)
    snap = {"note": "snapshot of the rand 0.9.2 restatement; parity with the real crate is unpinned",
            "stdrng_first_words": {str(s): [po.StdRng(s).next_u32() for _ in range(1)] for s in (0, 1, 2, 42)},
            "perm16": {str(s): po.perm_from_seed(16, s) for s in (1, 2, 0xB9736F582676E7E8, 0xD7397E6260CE9C3E)},
            "keccak_transcript_seeds": [hex(x) for x in (lambda t: (t.get_u64(), t.get_u64()))(po.KeccakTranscript())]}
    with open(os.path.join(HERE, "rand_snapshot.json"), "w") as f:
        json.dump(snap, f, indent=0)
    print("wrote", len(out), "commit vectors")


if __name__ == "__main__":
    main()
