#!/usr/bin/env python
"""scripts/mgpu_bench.py -- zipgpu_mgpu_* (ONE process, all visible GPUs) timed through the C ABI with host buffers.

    python scripts/mgpu_bench.py [--nv 24] [--reps 20]

  commit_resident : one 2^nv commit sharded by row range over the GPUs; every call copies the host slices over all PCIe
                    links, runs the kernels with the in-kernel roots exchange, and returns all roots (one D2H from device 0)
  batch_commit    : 64 x 2^18, polynomial p on device p mod n
Roots are checked against a single-GPU commit.  One JSON line."""
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
    ap.add_argument("--nv", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--devices", type=int, default=0, help="use the first n devices (0 = all)")
    args = ap.parse_args()
    import torch

    from zinc_b200 import Context, MultiContext, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    L = nat.lib()
    m = MultiContext(n=args.devices)
    nv = args.nv
    row_len = 1 << ((nv + 1) // 2)
    rows = (1 << nv) // row_len
    cw = 2 * row_len
    seeds = (0xB9736F582676E7E8, 0xD7397E6260CE9C3E)
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, seeds[0]), shuffle_seeded_indices(cw, seeds[1]))
    hm = code.native(m, 1, 4)
    pinned = torch.empty(1 << nv, dtype=torch.int64).pin_memory()
    pinned.numpy().view(np.uint64)[:] = np.random.Generator(np.random.PCG64(nv)).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    roots = torch.empty(rows * 32, dtype=torch.uint8).pin_memory()

    def commit():
        h = C.c_void_p()
        nat.check(L.zipgpu_mgpu_commit_resident(hm, rows, pinned.data_ptr(), roots.data_ptr(), C.byref(h)))
        L.zipgpu_mgpu_data_free(h)

    for _ in range(5):
        commit()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        commit()
    ms = (time.perf_counter() - t0) / args.reps * 1e3
    # single-GPU reference
    c0 = Context(0)
    h0 = code.native(c0, 1, 4)
    ref = torch.empty(rows * 32, dtype=torch.uint8).pin_memory()
    hd = C.c_void_p()
    nat.check(L.zipgpu_commit_resident(h0, rows, pinned.data_ptr(), ref.data_ptr(), C.byref(hd)))
    L.zipgpu_data_free(hd)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        hd = C.c_void_p()
        nat.check(L.zipgpu_commit_resident(h0, rows, pinned.data_ptr(), ref.data_ptr(), C.byref(hd)))
        L.zipgpu_data_free(hd)
    ms1 = (time.perf_counter() - t0) / args.reps * 1e3
    out = {"n_devices": m.num_devices, "nv": nv, "mgpu_commit_resident_ms": round(ms, 4), "single_gpu_commit_resident_ms": round(ms1, 4),
           "evals_per_s": (1 << nv) / (ms * 1e-3), "roots_equal_single_gpu": bool(torch.equal(roots, ref))}
    # batch: 64 x 2^18
    bnv, npoly = 18, 64
    brl = 1 << ((bnv + 1) // 2)
    brows, bcw = (1 << bnv) // brl, 2 * brl
    bcode = RaaCode.with_permutations(ZipTypes(), brl, 2, shuffle_seeded_indices(bcw, seeds[0]), shuffle_seeded_indices(bcw, seeds[1]))
    bm, b0 = bcode.native(m, 1, 4), bcode.native(c0, 1, 4)
    polys = [torch.from_numpy(np.random.Generator(np.random.PCG64(100 + i)).integers(0, 1 << 63, size=1 << bnv, dtype=np.int64)).pin_memory()
             for i in range(npoly)]
    r_m = [torch.empty(brows * 32, dtype=torch.uint8).pin_memory() for _ in range(npoly)]
    r_1 = [torch.empty(brows * 32, dtype=torch.uint8).pin_memory() for _ in range(npoly)]
    ev = (C.c_void_p * npoly)(*[p.data_ptr() for p in polys])
    am = (C.c_void_p * npoly)(*[r.data_ptr() for r in r_m])
    a1 = (C.c_void_p * npoly)(*[r.data_ptr() for r in r_1])
    for fn, h, arr, key in ((L.zipgpu_mgpu_batch_commit, bm, am, "mgpu_batch_64x2^18_ms"), (L.zipgpu_batch_commit, b0, a1, "single_gpu_batch_64x2^18_ms")):
        for _ in range(3):
            nat.check(fn(h, npoly, brows, ev, None, None, arr))
        t0 = time.perf_counter()
        for _ in range(10):
            nat.check(fn(h, npoly, brows, ev, None, None, arr))
        out[key] = round((time.perf_counter() - t0) / 10 * 1e3, 4)
    out["batch_roots_equal_single_gpu"] = all(torch.equal(a, b) for a, b in zip(r_m, r_1))
    print(json.dumps(out), flush=True)
    c0.close()
    m.close()


if __name__ == "__main__":
    main()
