import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).cuda()
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda')
s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
f = lambda: nat.check(L.zipgpu_encode_rows_device(h, num_rows, d.data_ptr(), dr.data_ptr(), sp))
for _ in range(3): f()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(20): f()
e1.record(s); torch.cuda.synchronize()
print(f"encode {e0.elapsed_time(e1)/20:.4f} ms")
