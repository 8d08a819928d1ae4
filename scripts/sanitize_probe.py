#!/usr/bin/env python
"""scripts/sanitize_probe.py -- a few small commits through every fused kernel variant, for compute-sanitizer
(memcheck / racecheck / synccheck):  compute-sanitizer --tool racecheck python scripts/sanitize_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ZIPGPU_FUSE_MIN_ROWS", "1")
os.environ.setdefault("ZIPGPU_WS16K_MIN_ROWS", "1")


def main():
    import ctypes as C

    import torch

    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    L = nat.lib()
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    cases = [(256, 40, "1", "0"), (512, 40, "1", "1"), (1024, 30, "1", "1"), (2048, 20, "2", "1"), (4096, 9, "2", "1"),
             (4096, 5, "1", "1"), (4096, 7, "1", "0"), (8192, 3, "1", "0")]
    if len(sys.argv) > 1:
        cases = cases[: int(sys.argv[1])]
    for row_len, rows, units, tops in cases:
        os.environ["ZIPGPU_WS_UNITS"] = units
        os.environ["ZIPGPU_WS_TOPS"] = tops
        cw = 2 * row_len
        depth = cw.bit_length() - 1
        code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
        h = code.native(ctx, 1, 4)
        ev = torch.from_numpy(np.random.default_rng(rows).integers(-2**63, 2**63 - 1, size=rows * row_len)).to(dev)
        d_rows = torch.empty(rows * cw * 4, dtype=torch.int64, device=dev)
        d_lay = torch.empty(rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
        d_roots = torch.empty(rows * 32, dtype=torch.uint8, device=dev)
        nat.check(L.zipgpu_commit_device(h, rows, ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(), d_roots.data_ptr(), None))
        ctx.sync()
        print("ok", row_len, rows, units, tops, d_roots[:4].cpu().numpy().tolist(), flush=True)
        ctx.drop_code(code)


if __name__ == "__main__":
    main()
