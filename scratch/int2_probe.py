import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, RandomFieldZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
for nv in (20, 22, 24):
    row_len = 1 << ((nv + 1) // 2); num_rows = (1 << nv) // row_len; cw = 2 * row_len; depth = cw.bit_length() - 1
    code = RaaCode.with_permutations(RandomFieldZipTypes(2), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    h = code.native(ctx, 2, 8)
    d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=2 << nv)).cuda()
    dr = torch.empty(num_rows*cw*8, dtype=torch.int64, device='cuda')
    dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro = torch.empty(num_rows*32, dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
    for name, f in (("encode", lambda: nat.check(L.zipgpu_encode_rows_device(h, num_rows, d.data_ptr(), dr.data_ptr(), sp))),
                    ("commit", lambda: nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), sp)))):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): f()
        e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/10
        print(f"Int<2> nv={nv} {name}: {ms:.4f} ms" + (f"  ({(16+128)*(1<<nv)/ms/1e6:.0f} GB/s of 144 B/eval)" if name=="encode" else f"  ({(1<<nv)/ms/1e6:.2f} Gevals/s)"))
    del d, dr, dl, dro
    torch.cuda.empty_cache()
