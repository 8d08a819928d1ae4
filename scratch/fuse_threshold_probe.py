import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
def run(nv, rows_override=None):
    row_len = 1 << ((nv + 1) // 2); num_rows = (1 << nv) // row_len; cw = 2 * row_len; depth = cw.bit_length() - 1
    if rows_override: num_rows = rows_override
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    h = code.native(ctx, 1, 4)
    d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=num_rows*row_len)).cuda()
    dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda')
    dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro = torch.empty(num_rows*32, dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
    out = {}
    for mode in ("fused", "two-kernel"):
        os.environ.pop("ZIPGPU_NO_FUSE", None); os.environ["ZIPGPU_FUSE_MIN_ROWS"] = "1"
        if mode == "two-kernel": os.environ["ZIPGPU_NO_FUSE"] = "1"
        f = lambda: nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), sp))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): f()
        e1.record(s); torch.cuda.synchronize()
        out[mode] = e0.elapsed_time(e1)/20
    print(f"cw={cw} rows={num_rows}: fused {out['fused']:.4f} ms, two-kernel {out['two-kernel']:.4f} ms")
run(16); run(18); run(20); run(22)
for r in (64, 128, 256, 512, 1024, 2048): run(24, r)
