timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "warp_specialised" 2>&1 | tail -2
echo "--- WS forced"; ZIPGPU_FUSE_MIN_ROWS=1 timeout 100 python scripts/size_sweep.py --nv 12 13 14 15 16 --reps 100 2>&1 | cut -c1-120
echo "--- old fused forced"; ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_NO_WS=1 timeout 100 python scripts/size_sweep.py --nv 12 13 14 15 16 --reps 100 2>&1 | cut -c1-120
echo "--- default"; timeout 100 python scripts/size_sweep.py --nv 12 13 14 15 16 --reps 100 2>&1 | cut -c1-120
