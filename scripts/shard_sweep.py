#!/usr/bin/env python
"""scripts/shard_sweep.py -- device time of zipgpu_commit_device for the row counts a sharded commit gives one GPU.

    python scripts/shard_sweep.py [--row-len 4096] [--rows 256,512,1024,2048,4096] [--reps 20]

One line per row count: ms per commit, the fused-kernel / upper-pass split from the library's own CUDA events, and
the time relative to the alu-pipe floor (456 lane-ops per compression at 64 lanes/clk/SM).  Environment knobs
(ZIPGPU_WS_UNITS, ZIPGPU_FUSE_MIN_ROWS, ZIPGPU_NO_FUSE, ZIPGPU_CTA_TREE_LOG2) select the variant under test."""
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
    ap.add_argument("--rows", default="256,512,1024,2048,4096")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch

    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    L = nat.lib()
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    row_len, cw = args.row_len, 2 * args.row_len
    depth = cw.bit_length() - 1
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    h = code.native(ctx, 1, 4)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    knobs = {k: v for k, v in os.environ.items() if k.startswith("ZIPGPU_")}
    try:
        import pynvml

        pynvml.nvmlInit()
        nvh = pynvml.nvmlDeviceGetHandleByIndex(0)
    except Exception:
        pynvml = None
    for rows in [int(x) for x in args.rows.split(",")]:
        ev = torch.from_numpy(np.random.default_rng(rows).integers(0, 1 << 63, size=rows * row_len, dtype=np.int64)).to(dev)
        d_rows = torch.empty(rows * cw * 4, dtype=torch.int64, device=dev)
        d_lay = torch.empty(rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
        d_roots = torch.empty(rows * 32, dtype=torch.uint8, device=dev)
        run = lambda: nat.check(L.zipgpu_commit_device(h, rows, ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(),
                                                       d_roots.data_ptr(), sptr))
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        nat.check(L.zipgpu_profile_read(ctx.handle, None, None, None, 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 1))
        l0 = ctx.launch_count
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.reps):
            run()
        b.record(stream)
        mhz = None
        if pynvml is not None:  # the GPU is still busy with the queued launches: the clock under load
            clk = []
            while not b.query():
                clk.append(pynvml.nvmlDeviceGetClockInfo(nvh, pynvml.NVML_CLOCK_SM))
            mhz = float(np.median(clk)) if clk else None
        torch.cuda.synchronize()
        e_, h_, c_ = C.c_double(), C.c_double(), C.c_uint64()
        nat.check(L.zipgpu_profile_read(ctx.handle, C.byref(e_), C.byref(h_), C.byref(c_), 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 0))
        ms = a.elapsed_time(b) / args.reps
        floor_ms = rows * (2 * cw - 1) * 456.0 / (64.0 * sms * 1.965e9) * 1e3
        print(json.dumps({"row_len": row_len, "rows": rows, "ms": round(ms, 4), "first_kernel_ms": round(e_.value / c_.value, 4),
                          "rest_ms": round(h_.value / c_.value, 4), "launches": (ctx.launch_count - l0) // args.reps,
                          "x_alu_floor": round(ms / floor_ms, 3), "sm_mhz": mhz, "knobs": knobs}), flush=True)
        del ev, d_rows, d_lay, d_roots
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
