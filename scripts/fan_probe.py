#!/usr/bin/env python
"""scripts/fan_probe.py -- cost of the fused roots exchange on ONE GPU (world = 1: the launch that produces the roots also
stores them into the exchange buffer and its last CTA shakes hands with itself) against the plain commit.

    python scripts/fan_probe.py [--row-len 4096] [--rows 512,2048]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--row-len", type=int, default=4096)
    ap.add_argument("--rows", default="512,2048")
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    import torch

    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat
    from zinc_b200.dist import PeerRoots

    L = nat.lib()
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    row_len, cw = args.row_len, 2 * args.row_len
    depth = cw.bit_length() - 1
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    h = code.native(ctx, 1, 4)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    for rows in [int(x) for x in args.rows.split(",")]:
        ev = torch.from_numpy(np.random.default_rng(rows).integers(0, 1 << 63, size=rows * row_len, dtype=np.int64)).to(dev)
        d_rows = torch.empty(rows * cw * 4, dtype=torch.int64, device=dev)
        d_lay = torch.empty(rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
        d_roots = torch.empty(rows * 32, dtype=torch.uint8, device=dev)
        peer = PeerRoots(ctx, rows)
        plain = lambda: nat.check(L.zipgpu_commit_device(h, rows, ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(),
                                                         d_roots.data_ptr(), sptr))
        fused = lambda: peer.commit_device(h, 0, rows, ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(), sptr)
        out = {"row_len": row_len, "rows": rows}
        for name, fn in (("plain_ms", plain), ("sharded_world1_ms", fused), ("plain2_ms", plain), ("sharded_world1_2_ms", fused)):
            for _ in range(5):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(args.reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            out[name] = round(a.elapsed_time(b) / args.reps, 4)
        peer.status()
        print(json.dumps(out), flush=True)
        peer.close()


if __name__ == "__main__":
    main()
