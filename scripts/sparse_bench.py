"""scripts/sparse_bench.py -- times the ZipLinearCode (sparse code) encoder and commit on one GPU.

    python scripts/sparse_bench.py [--nv 24] [--steps 10]

Synthetic 0/1 matrices (row_len/2 cells per matrix row, every coefficient 1: the densest case the reference can
sample) and uniform i64 evaluations.  CUDA events inside the library (zipgpu_profile_*) on its launch stream.
Prints one JSON line; mma_tops counts 2*cw*row_len*(rows*8 planes) integer operations.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--generic", action="store_true")
    args = ap.parse_args()
    if args.generic:
        os.environ["ZIPGPU_SPARSE_GENERIC"] = "1"
    import torch

    from zinc_b200 import SparseMatrixZ, ZipLinearCode, ZipTypes, _native as nat, default_context

    nv = args.nv
    row_len = 1 << ((nv + 1) // 2)
    num_rows = (1 << nv) // row_len
    cw, d = 2 * row_len, row_len // 2
    rng = np.random.default_rng(nv)

    def matrix():
        cols = np.empty((cw // 2, d), dtype=np.uint32)
        for i in range(cw // 2):
            cols[i] = np.sort(rng.permutation(row_len)[:d])
        return SparseMatrixZ(cw // 2, row_len, d, cols, np.ones(cw // 2 * d, dtype=np.int64))

    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, matrix(), matrix())
    ctx = default_context()
    L = nat.lib()
    h = code.native(ctx, 1, 4)
    kind = code.kernel_kind(ctx)
    depth = cw.bit_length() - 1
    dev = torch.device("cuda:0")
    evals = torch.from_numpy(rng.integers(-(1 << 63), (1 << 63) - 1, size=1 << nv, dtype=np.int64)).to(dev)
    rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
    layers = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
    roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        nat.check(L.zipgpu_profile_read(ctx.handle, None, None, None, 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 1))
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        e_, h_, c_ = C.c_double(), C.c_double(), C.c_uint64()
        nat.check(L.zipgpu_profile_read(ctx.handle, C.byref(e_), C.byref(h_), C.byref(c_), 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 0))
        return e_.value / max(c_.value, 1), h_.value / max(c_.value, 1)

    enc_ms, _ = timed(lambda: nat.check(L.zipgpu_encode_rows_device(h, num_rows, evals.data_ptr(), rows.data_ptr(), None)))
    c_enc, c_hash = timed(lambda: nat.check(L.zipgpu_commit_device(h, num_rows, evals.data_ptr(), rows.data_ptr(),
                                                                   layers.data_ptr(), roots.data_ptr(), None)))
    ops = 2.0 * cw * row_len * num_rows * 8
    print(json.dumps({"nv": nv, "kernel": kind, "row_len": row_len, "num_rows": num_rows, "cw": cw,
                      "encode_ms": enc_ms, "commit_encode_ms": c_enc, "commit_merkle_ms": c_hash,
                      "commit_ms": c_enc + c_hash, "evals_per_s": (1 << nv) / ((c_enc + c_hash) * 1e-3),
                      "mma_tops": ops / (enc_ms * 1e-3) / 1e12 if kind == "tensor" else None}))


if __name__ == "__main__":
    main()
