"""scripts/size_sweep.py -- device-resident commit / encode / tree times for a range of MLE sizes on one GPU.

    python scripts/size_sweep.py [--nv 17 18 ... ] [--reps 20]

One JSON line per size: ms per call (CUDA events on the launch stream), evals/s, and the ratio to the INT32-alu floor
of the hashing (4 - 1/row_len compressions per evaluation at 480 alu lane-instructions, 64 lanes/clk/SM).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, nargs="+", default=[16, 17, 18, 19, 20, 21, 22, 23, 24, 25])
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch

    from helpers import KECCAK_SEEDS, shape_for
    from zinc_b200 import RaaCode, ZipTypes, _native as nat, default_context, shuffle_seeded_indices

    ctx = default_context()
    L = nat.lib()
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    def ms(fn):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(args.reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.reps

    for nv in args.nv:
        row_len, num_rows, cw = shape_for(nv)
        depth = cw.bit_length() - 1
        code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, KECCAK_SEEDS[0]),
                                         shuffle_seeded_indices(cw, KECCAK_SEEDS[1]))
        h = code.native(ctx, 1, 4)
        ev = torch.from_numpy(np.random.default_rng(nv).integers(-(1 << 63), (1 << 63) - 1, size=1 << nv, dtype=np.int64)).to(dev)
        rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
        lay = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
        roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
        commit = ms(lambda: nat.check(L.zipgpu_commit_device(h, num_rows, ev.data_ptr(), rows.data_ptr(), lay.data_ptr(), roots.data_ptr(), sptr)))
        enc = ms(lambda: nat.check(L.zipgpu_encode_rows_device(h, num_rows, ev.data_ptr(), rows.data_ptr(), sptr)))
        tree = ms(lambda: nat.check(L.zipgpu_merkle_rows_device(ctx.handle, num_rows, depth, 4, rows.data_ptr(), lay.data_ptr(), roots.data_ptr(), sptr)))
        floor_ms = num_rows * (2 * cw - 1) * 480.0 * 32 / 32 / (64.0 * sms * 1.965e9) * 1e3
        print(json.dumps({"nv": nv, "rows": num_rows, "cw": cw, "commit_ms": round(commit, 4), "encode_ms": round(enc, 4),
                          "tree_ms": round(tree, 4), "evals_per_s": float("%.3e" % ((1 << nv) / (commit * 1e-3))),
                          "alu_floor_ms": round(floor_ms, 4), "commit_over_floor": round(commit / floor_ms, 3)}), flush=True)
        del ev, rows, lay, roots
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
