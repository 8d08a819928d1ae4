#!/usr/bin/env python
"""scripts/e2e_sizes.py -- host-to-host commit time (zipgpu_commit_resident: pinned evaluations in, roots out, prover data
resident) by size, with the library's per-chunk timeline for one call when ZIPGPU_TIMELINE=1.

    python scripts/e2e_sizes.py [--nv 20,22,24] [--rows-of-24 512,1024]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", default="20,22,24")
    ap.add_argument("--shards", default="512,1024,2048", help="row counts of a sharded nv=24 commit (row_len 4096)")
    args = ap.parse_args()
    import torch

    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    L = nat.lib()
    ctx = Context(0)

    def run(row_len, rows, label):
        cw = 2 * row_len
        code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
        h = code.native(ctx, 1, 4)
        pinned = torch.empty(rows * row_len, dtype=torch.int64).pin_memory()
        pinned.numpy()[:] = np.random.default_rng(0).integers(-2**63, 2**63 - 1, size=rows * row_len)
        roots = torch.empty(rows * 32, dtype=torch.uint8).pin_memory()

        def e2e():
            hh = C.c_void_p()
            nat.check(L.zipgpu_commit_resident(h, rows, pinned.data_ptr(), roots.data_ptr(), C.byref(hh)))
            L.zipgpu_data_free(hh)

        for _ in range(5):
            e2e()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e()
        dt = (time.perf_counter() - t0) / reps
        mib = rows * row_len * 8 / 2**20
        print(json.dumps({"case": label, "rows": rows, "row_len": row_len, "e2e_ms": round(dt * 1e3, 4), "h2d_MiB": mib,
                          "copy_only_ms_at_55GBps": round(mib * 2**20 / 55.4e9 * 1e3, 4)}), flush=True)
        ctx.drop_code(code)

    for nv in [int(x) for x in args.nv.split(",") if x]:
        row_len = 1 << ((nv + 1) // 2)
        run(row_len, (1 << nv) // row_len, f"nv={nv}")
    for rows in [int(x) for x in args.shards.split(",") if x]:
        run(4096, rows, f"nv=24 shard of {rows} rows")


if __name__ == "__main__":
    main()
