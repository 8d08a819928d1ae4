timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "warp_specialised or fused" 2>&1 | tail -3
echo "--- forced fused (WS)"; ZIPGPU_FUSE_MIN_ROWS=1 timeout 200 python scripts/size_sweep.py --nv 18 19 20 21 22 2>&1 | cut -c1-140
echo "--- forced fused, no WS"; ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_NO_WS=1 timeout 200 python scripts/size_sweep.py --nv 19 20 21 22 2>&1 | cut -c1-140
echo "--- no fuse"; ZIPGPU_NO_FUSE=1 timeout 200 python scripts/size_sweep.py --nv 19 20 21 22 2>&1 | cut -c1-140
