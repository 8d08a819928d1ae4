timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "warp_specialised" 2>&1 | tail -2
echo "--- WS forced"; ZIPGPU_FUSE_MIN_ROWS=1 timeout 100 python scripts/size_sweep.py --nv 17 18 2>&1 | cut -c1-120
echo "--- old fused forced"; ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_NO_WS=1 timeout 100 python scripts/size_sweep.py --nv 17 18 2>&1 | cut -c1-120
echo "--- default"; timeout 100 python scripts/size_sweep.py --nv 17 18 2>&1 | cut -c1-120
python - <<'P'
import ctypes as C, numpy as np, torch, os, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from helpers import KECCAK_SEEDS
from zinc_b200 import RaaCode, ZipTypes, _native as nat, default_context, shuffle_seeded_indices
ctx=default_context(); L=nat.lib(); dev=torch.device('cuda:0')
row_len,cw,rows=512,1024,32768; depth=10
code=RaaCode.with_permutations(ZipTypes(),row_len,2,shuffle_seeded_indices(cw,KECCAK_SEEDS[0]),shuffle_seeded_indices(cw,KECCAK_SEEDS[1]))
h=code.native(ctx,1,4)
ev=torch.randint(-2**62,2**62,(rows*row_len,),dtype=torch.int64,device=dev)
r=torch.empty(rows*cw*4,dtype=torch.int64,device=dev); l=torch.empty(rows*((2<<depth)-2)*32,dtype=torch.uint8,device=dev); ro=torch.empty(rows*32,dtype=torch.uint8,device=dev)
for knob in (None,'ZIPGPU_NO_WS','ZIPGPU_NO_FUSE'):
    if knob: os.environ[knob]='1'
    f=lambda: nat.check(L.zipgpu_commit_device(h,rows,ev.data_ptr(),r.data_ptr(),l.data_ptr(),ro.data_ptr(),None))
    for _ in range(3): f()
    ctx.sync(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    import time; t0=time.perf_counter()
    for _ in range(10): f()
    ctx.sync(); print('batch-shape 32768 x 1024', knob, round((time.perf_counter()-t0)/10*1e3,4),'ms')
P
