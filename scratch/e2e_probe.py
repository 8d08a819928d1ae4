import ctypes as C, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
pinned = torch.empty(1<<nv, dtype=torch.int64).pin_memory()
pinned.numpy()[:] = np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)
roots = torch.empty(num_rows*32, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<nv, dtype=torch.int64, device='cuda')
for _ in range(3): d.copy_(pinned, non_blocking=True)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): d.copy_(pinned, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print(f"H2D 128MiB: {dt*1e3:.2f} ms = {(1<<27)/dt/1e9:.1f} GB/s")
def e2e():
    hh = C.c_void_p()
    nat.check(L.zipgpu_commit_resident(h, num_rows, pinned.data_ptr(), roots.data_ptr(), C.byref(hh)))
    L.zipgpu_data_free(hh)
for chunk in [0, 4096, 2048, 1024, 512, 256]:
    if chunk: os.environ['ZIPGPU_CHUNK_ROWS']=str(chunk)
    for _ in range(3): e2e()
    t0=time.perf_counter()
    for _ in range(10): e2e()
    dt=(time.perf_counter()-t0)/10
    print(f"chunk_rows={chunk or 'default'}: e2e {dt*1e3:.2f} ms")
# device-only for comparison
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda'); dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro=torch.empty(num_rows*32,dtype=torch.uint8,device='cuda')
for _ in range(3): nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), None))
ctx.sync(); t0=time.perf_counter()
for _ in range(10): nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), None))
ctx.sync(); print(f"device commit: {(time.perf_counter()-t0)/10*1e3:.2f} ms")
for _ in range(3): nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), None, None, dro.data_ptr(), None))
ctx.sync(); t0=time.perf_counter()
for _ in range(10): nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), None, None, dro.data_ptr(), None))
ctx.sync(); print(f"device commit (scratch alloc): {(time.perf_counter()-t0)/10*1e3:.2f} ms")
