import ctypes as C, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv = int(sys.argv[1])
row_len = 1 << ((nv + 1) // 2); num_rows = (1 << nv) // row_len; cw = 2 * row_len
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx, 1, 4)
pinned = torch.empty(1 << nv, dtype=torch.int64).pin_memory()
pinned.numpy()[:] = np.random.default_rng(0).integers(-2**63, 2**63-1, size=1 << nv)
roots = torch.empty(num_rows * 32, dtype=torch.uint8).pin_memory()
def e2e():
    hh = C.c_void_p()
    nat.check(L.zipgpu_commit_resident(h, num_rows, pinned.data_ptr(), roots.data_ptr(), C.byref(hh)))
    L.zipgpu_data_free(hh)
for _ in range(5): e2e()
t0 = time.perf_counter()
for _ in range(30): e2e()
dt = (time.perf_counter() - t0) / 30
print(f"nv={nv} chunk={os.environ.get('ZIPGPU_CHUNK_ROWS','default')}: e2e {dt*1e3:.3f} ms")
