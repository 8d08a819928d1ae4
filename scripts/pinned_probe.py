"""scripts/pinned_probe.py -- the nv = 20 host-to-host commit (8 MiB in) at several points of a bench-like sequence, from a torch-pinned buffer and from a zipgpu_host_alloc (2 MiB-page-backed) one: shows whether the DMA rate of freshly pinned memory depends on the history of the process on this box"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
from zinc_b200 import _native as nat
L = nat.lib()
ctx = Context(0)
dev = torch.device("cuda", 0)

def code_for(row_len):
    cw = 2 * row_len
    c = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    return c, c.native(ctx, 1, 4)

c20, h20 = code_for(1024)
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
class LibPin:
    def __init__(self, n):
        p = C.c_void_p(); nat.check(L.zipgpu_host_alloc(n * 8, C.byref(p))); self.p = p
        arr = np.frombuffer((C.c_int64 * n).from_address(p.value), dtype=np.int64)
        arr[:] = np.random.default_rng(0).integers(-2**63, 2**63 - 1, size=n)
    def data_ptr(self): return self.p.value
    def __del__(self): L.zipgpu_host_free(self.p)
def dbg(tag, pin=None):
    if pin is None:
        pin = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63 - 1, size=1 << 20)).pin_memory()
        lp = LibPin(1 << 20)
        dbg(tag + " [zipgpu_host_alloc]", lp)
    roots = torch.empty(1024 * 32, dtype=torch.uint8).pin_memory()
    def call():
        hh = C.c_void_p()
        nat.check(L.zipgpu_commit_resident(h20, 1024, pin.data_ptr(), roots.data_ptr(), C.byref(hh)))
        L.zipgpu_data_free(hh)
    for _ in range(5): call()
    t0 = time.perf_counter()
    for _ in range(20): call()
    print(tag, round((time.perf_counter() - t0) / 20 * 1e3, 4), "ms", flush=True)
    return pin

p0 = dbg("A fresh process")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
dbg("B torch side stream current")
c24, h24 = code_for(4096)
ev = torch.from_numpy(np.random.default_rng(1).integers(-2**63, 2**63 - 1, size=1 << 24)).to(dev)
rows = torch.empty(4096 * 8192 * 4, dtype=torch.int64, device=dev)
lay = torch.empty(4096 * (2 * 8192 - 2) * 32, dtype=torch.uint8, device=dev)
roots = torch.empty(4096 * 32, dtype=torch.uint8, device=dev)
sptr = C.c_void_p(stream.cuda_stream)
for _ in range(10):
    nat.check(L.zipgpu_commit_device(h24, 4096, ev.data_ptr(), rows.data_ptr(), lay.data_ptr(), roots.data_ptr(), sptr))
torch.cuda.synchronize()
dbg("C after nv=24 device commits + 3 GiB of torch buffers")
dbg("C' same, reusing the first pinned buffer", p0)
pin24 = torch.empty(1 << 24, dtype=torch.int64).pin_memory()
r24 = torch.empty(4096 * 32, dtype=torch.uint8).pin_memory()
for _ in range(5):
    hh = C.c_void_p()
    nat.check(L.zipgpu_commit_resident(h24, 4096, pin24.data_ptr(), r24.data_ptr(), C.byref(hh)))
    L.zipgpu_data_free(hh)
dbg("D after nv=24 host commits (128 MiB pinned)")
dbg("D' reusing the first pinned buffer", p0)
peak = C.c_double()
try:
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 2, C.byref(peak), None))
except Exception as e:
    print("microbench", e)
dbg("E after microbench")
c16, h16 = code_for(256)
e16 = torch.from_numpy(np.random.default_rng(2).integers(-2**63, 2**63 - 1, size=1 << 16)).to(dev)
r16 = torch.empty(256 * 512 * 4, dtype=torch.int64, device=dev); l16 = torch.empty(256 * 1022 * 32, dtype=torch.uint8, device=dev); o16 = torch.empty(256 * 32, dtype=torch.uint8, device=dev)
for _ in range(10):
    nat.check(L.zipgpu_commit_device(h16, 256, e16.data_ptr(), r16.data_ptr(), l16.data_ptr(), o16.data_ptr(), sptr))
torch.cuda.synchronize()
del e16, r16, l16, o16
torch.cuda.empty_cache()
dbg("F after nv=16 + empty_cache")
e20 = torch.from_numpy(np.random.default_rng(3).integers(-2**63, 2**63 - 1, size=1 << 20)).to(dev)
r20 = torch.empty(1024 * 2048 * 4, dtype=torch.int64, device=dev); l20 = torch.empty(1024 * 4094 * 32, dtype=torch.uint8, device=dev); o20 = torch.empty(1024 * 32, dtype=torch.uint8, device=dev)
for _ in range(10):
    nat.check(L.zipgpu_commit_device(h20, 1024, e20.data_ptr(), r20.data_ptr(), l20.data_ptr(), o20.data_ptr(), sptr))
    nat.check(L.zipgpu_encode_rows_device(h20, 1024, e20.data_ptr(), r20.data_ptr(), sptr))
torch.cuda.synchronize()
dbg("G after nv=20 device commits on the torch stream (bench order)")
