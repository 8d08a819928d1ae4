import ctypes as C, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices, _native as nat
L = nat.lib(); ctx = Context(0)
nv=24; row_len=4096; num_rows=4096; cw=8192; depth=13
code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
h = code.native(ctx,1,4)
dev=torch.device('cuda')
d = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63-1, size=1<<nv)).to(dev)
dr = torch.empty(num_rows*cw*4, dtype=torch.int64, device=dev); dl = torch.empty(num_rows*(2*cw-2)*32, dtype=torch.uint8, device=dev); dro=torch.empty(num_rows*32,dtype=torch.uint8,device=dev)
ref_roots=None
def run(nchunks, prio):
    sa = torch.cuda.Stream(priority=prio); sb = torch.cuda.Stream()
    rows_per = num_rows//nchunks
    def step():
        evs=[]
        for c in range(nchunks):
            r0=c*rows_per
            nat.check(L.zipgpu_encode_rows_device(h, rows_per, d.data_ptr()+r0*row_len*8, dr.data_ptr()+r0*cw*32, C.c_void_p(sa.cuda_stream)))
            ev=torch.cuda.Event(); ev.record(sa); evs.append(ev)
        for c in range(nchunks):
            r0=c*rows_per
            sb.wait_event(evs[c])
            nat.check(L.zipgpu_merkle_rows_device(ctx.handle, rows_per, depth, 4, dr.data_ptr()+r0*cw*32, dl.data_ptr()+r0*(2*cw-2)*32, dro.data_ptr()+r0*32, C.c_void_p(sb.cuda_stream)))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    main=torch.cuda.current_stream()
    e0.record(main); sa.wait_event(e0); sb.wait_event(e0)
    for _ in range(10): step()
    ee=torch.cuda.Event(); ee.record(sb); main.wait_event(ee); ee2=torch.cuda.Event(); ee2.record(sa); main.wait_event(ee2)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/10
base=None
for nchunks in (1,2,4,8,16):
    for prio in (0,-1):
        ms=run(nchunks,prio)
        r=dro.cpu().numpy().copy()
        if ref_roots is None: ref_roots=r
        print(f"chunks={nchunks} enc_prio={prio}: {ms:.3f} ms  roots_ok={np.array_equal(r,ref_roots)}")
