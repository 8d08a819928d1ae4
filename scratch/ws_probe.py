import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
G=148
dbg = torch.zeros(G*4, dtype=torch.int64, device='cuda')
os.environ['ZIPGPU_DBG_PTR'] = str(dbg.data_ptr())
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).cuda()
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda'); dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro=torch.empty(num_rows*32,dtype=torch.uint8,device='cuda')
for _ in range(2):
    dbg.zero_(); torch.cuda.synchronize()
    nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), None)); ctx.sync()
t = dbg.cpu().numpy().reshape(G,4)/1e3
print("ENC : wait-for-empty us median %.0f, work us median %.0f" % (np.median(t[:,0]), np.median(t[:,1])))
print("HASH: wait-for-full  us median %.0f, work us median %.0f" % (np.median(t[:,2]), np.median(t[:,3])))
print("per row (27.7 rows/SM): ENC work %.1f us, HASH work %.1f us, HASH wait %.1f us" % (np.median(t[:,1])/27.7, np.median(t[:,3])/27.7, np.median(t[:,2])/27.7))
