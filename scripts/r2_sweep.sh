set -x
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_reference_tests.py -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests3.log
S=gpurun_out/r2_shard_sweep2.jsonl; : > $S
R="--rows 64,128,256,512,1024,2048,4096"
ZIPGPU_FUSE_MIN_ROWS=1 python scripts/shard_sweep.py $R >> $S 2>>gpurun_out/sweep.err
ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_WS_UNITS=1 python scripts/shard_sweep.py $R >> $S 2>>gpurun_out/sweep.err
ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_WS_UNITS=2 python scripts/shard_sweep.py $R >> $S 2>>gpurun_out/sweep.err
ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_WS_UNITS=4 python scripts/shard_sweep.py $R >> $S 2>>gpurun_out/sweep.err
for RL in 2048 1024 512 256; do
R2="--row-len $RL --rows 64,128,256,512,1024,2048,4096"
ZIPGPU_FUSE_MIN_ROWS=1 python scripts/shard_sweep.py $R2 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_WS_UNITS=1 python scripts/shard_sweep.py $R2 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_FUSE_MIN_ROWS=1 ZIPGPU_WS_UNITS=2 python scripts/shard_sweep.py $R2 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_NO_FUSE=1 python scripts/shard_sweep.py $R2 >> $S 2>>gpurun_out/sweep.err
done
tail -3 gpurun_out/r2_tests3.log
