import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
G=296
dbg = torch.zeros(G*64*8, dtype=torch.int64, device='cuda')
os.environ['ZIPGPU_DBG_PTR'] = str(dbg.data_ptr())
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).cuda()
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device='cuda'); dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device='cuda'); dro=torch.empty(num_rows*32,dtype=torch.uint8,device='cuda')
fused = len(sys.argv) < 2
for _ in range(2):
    dbg.zero_(); torch.cuda.synchronize()
    if fused: nat.check(L.zipgpu_commit_device(h, num_rows, d.data_ptr(), dr.data_ptr(), dl.data_ptr(), dro.data_ptr(), None))
    else: nat.check(L.zipgpu_encode_rows_device(h, num_rows, d.data_ptr(), dr.data_ptr(), None))
    ctx.sync()
t = dbg.cpu().numpy().reshape(G, 64, 8)
rows_done = (t[:, :, 4] > 0).sum(axis=1)
t0 = t[:, 0, 0].min()
print("rows per CTA: min %d median %d max %d; histogram:" % (rows_done.min(), np.median(rows_done), rows_done.max()), np.bincount(rows_done))
end = np.array([t[b, rows_done[b]-1, 4] for b in range(G)])
print("CTA end times us: min %.0f median %.0f max %.0f" % ((end.min()-t0)/1e3, (np.median(end)-t0)/1e3, (end.max()-t0)/1e3))
sm = t[:, 0, 6]
start_last = np.array([t[b, rows_done[b]-1, 0] for b in range(G)])
order = np.argsort(end)
print("last 12 CTAs to finish: (cta, rows, last-row start us, end us, last-row duration us)")
for b in order[-12:]:
    print("  ", b, rows_done[b], round((start_last[b]-t0)/1e3), round((end[b]-t0)/1e3), round((end[b]-start_last[b])/1e3))
pct = np.percentile((end - t0)/1e3, [0, 10, 25, 50, 75, 90, 100])
print("end-time percentiles us:", np.round(pct))
# per-SM finish = max over its CTAs
bysm = {}
for b in range(G): bysm.setdefault(int(sm[b]), []).append(end[b])
smend = np.array([max(v) for v in bysm.values()])
print("per-SM finish us: min %.0f median %.0f max %.0f; mean idle before kernel end %.0f us" % ((smend.min()-t0)/1e3, (np.median(smend)-t0)/1e3, (smend.max()-t0)/1e3, (smend.max()-smend.mean())/1e3))
for s_ in sorted(set(sm))[:3]:
    mates = [b for b in range(G) if sm[b] == s_]
    for b in mates:
        n = rows_done[b]
        enc = (t[b, :n, 1] - t[b, :n, 0]) / 1e3; wo = (t[b, :n, 2] - t[b, :n, 1]) / 1e3; hs = (t[b, :n, 3] - t[b, :n, 2]) / 1e3; st = (t[b, :n, 4] - t[b, :n, 3]) / 1e3
        print(f"SM {int(s_)} CTA {b}: rows {n}; per-row us median: encode {np.median(enc):.1f} writeout {np.median(wo):.1f} hash {np.median(hs):.1f} stage {np.median(st):.1f}; total {np.median(enc+wo+hs+st):.1f}")
