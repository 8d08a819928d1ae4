import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).cuda()
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda')
s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
p2 = shuffle_seeded_indices(cw, 2)
i = np.arange(cw); t = i // 16; k = i % 16
perms = {"random perm1": shuffle_seeded_indices(cw, 1), "conflict-free perm1 (transpose)": (k * 512 + t).astype(np.uint32)}
for name, p1 in perms.items():
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    h = code.native(ctx,1,4)
    f = lambda: nat.check(L.zipgpu_encode_rows_device(h, num_rows, d.data_ptr(), dr.data_ptr(), sp))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20): f()
    e1.record(s); torch.cuda.synchronize()
    print(f"{name}: encode {e0.elapsed_time(e1)/20:.4f} ms")
