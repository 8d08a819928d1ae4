import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
dbg = torch.zeros(3*296, dtype=torch.int64, device='cuda')
os.environ['ZIPGPU_DBG_PTR'] = str(dbg.data_ptr())
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192; depth=13
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).cuda()
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda'); dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro=torch.empty(num_rows*32,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize()
for _ in range(3):
    nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), None)); ctx.sync()
t = dbg.cpu().numpy()
end, start, sm = t[:296], t[296:592], t[592:]
t0 = start.min()
print("start spread us:", (start.max()-t0)/1e3, " end min/median/max us:", (end.min()-t0)/1e3, (np.median(end)-t0)/1e3, (end.max()-t0)/1e3)
dur = (end-start)/1e3
print("duration per CTA us: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (dur.min(), np.percentile(dur,10), np.median(dur), np.percentile(dur,90), dur.max()))
rows = np.array([len(range(b, 4096, 296)) for b in range(296)])
print("14-row CTAs: median dur %.1f; 13-row CTAs: median dur %.1f" % (np.median(dur[rows==14]), np.median(dur[rows==13])))
# SM pairing
from collections import defaultdict
bysm = defaultdict(list)
for b in range(296): bysm[int(sm[b])].append(b)
print("CTAs per SM histogram:", np.bincount([len(v) for v in bysm.values()]))
pairs = [v for v in bysm.values() if len(v)==2]
print("example pairs:", pairs[:6])
print("resident-time fraction: %.3f" % (dur.sum() / (296 * (end.max()-t0)/1e3)))
