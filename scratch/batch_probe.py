import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv=18; k=64; row_len=512; num_rows=512; cw=1024
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
pin = torch.empty((k, 1<<nv), dtype=torch.int64).pin_memory()
pin.numpy()[:] = np.random.default_rng(0).integers(-2**63, 2**63-1, size=(k, 1<<nv))
roots = torch.empty((k, num_rows*32), dtype=torch.uint8).pin_memory()
ev = (C.c_void_p * k)(*[pin[i].data_ptr() for i in range(k)])
ro = (C.c_void_p * k)(*[roots[i].data_ptr() for i in range(k)])
f = lambda: nat.check(L.zipgpu_batch_commit(h, k, num_rows, ev, None, None, ro))
for _ in range(3): f()
t0=time.perf_counter()
for _ in range(10): f()
dt=(time.perf_counter()-t0)/10
print(f"batch_commit 64 x 2^18 (host in, roots out): {dt*1e3:.3f} ms = {k*(1<<nv)/dt/1e9:.2f} Gevals/s")
