#!/usr/bin/env python
"""scripts/combine_probe.py -- device time of zipgpu_combine_rows_device (open_z.rs:100-113) by size, against the HBM floor"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zinc_b200 import Context
from zinc_b200 import _native as nat

L = nat.lib()
ctx = Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)
for nv in (16, 20, 22, 24, 26):
    row_len = 1 << ((nv + 1) // 2)
    rows = (1 << nv) // row_len
    ev = torch.from_numpy(np.random.default_rng(nv).integers(-2**63, 2**63 - 1, size=1 << nv)).to(dev)
    co = torch.from_numpy(np.random.default_rng(nv + 1).integers(-2**63, 2**63 - 1, size=rows)).to(dev)
    out = torch.empty(row_len * 8, dtype=torch.int64, device=dev)
    run = lambda: nat.check(L.zipgpu_combine_rows_device(ctx.handle, rows, row_len, ev.data_ptr(), co.data_ptr(), 8, out.data_ptr(), sptr))
    for _ in range(5):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(50):
        run()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(json.dumps({"nv": nv, "rows": rows, "row_len": row_len, "ms": round(ms, 4), "GBps": round((8 << nv) / ms / 1e6, 1)}), flush=True)
