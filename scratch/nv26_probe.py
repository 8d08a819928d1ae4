import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
for nv, rows in ((26, 8192), (25, 4096)):
    row_len = 8192; num_rows = rows; cw = 16384; depth = 14
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    h = code.native(ctx, 1, 4)
    d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=num_rows*row_len)).cuda()
    dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda')
    dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro = torch.empty(num_rows*32, dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
    for mode in ("fused", "two-kernel"):
        os.environ.pop("ZIPGPU_NO_FUSE", None)
        if mode == "two-kernel": os.environ["ZIPGPU_NO_FUSE"] = "1"
        f = lambda: nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), sp))
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): f()
        e1.record(s); torch.cuda.synchronize()
        print(f"nv={nv} {mode}: {e0.elapsed_time(e1)/5:.3f} ms")
    del d, dr, dl, dro; torch.cuda.empty_cache()
