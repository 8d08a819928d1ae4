S=gpurun_out/r2_tops3.jsonl; : > $S
python scripts/shard_sweep.py --rows 512,1024,4096 --reps 30 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_WS_TOPS=0 python scripts/shard_sweep.py --rows 512,1024 --reps 30 >> $S 2>>gpurun_out/sweep.err
python scripts/shard_sweep.py --row-len 2048 --rows 512,1024,2048 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_WS_TOPS=0 python scripts/shard_sweep.py --row-len 2048 --rows 512,1024,2048 >> $S 2>>gpurun_out/sweep.err
python scripts/shard_sweep.py --row-len 1024 --rows 512,1024,2048 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_WS_TOPS=0 python scripts/shard_sweep.py --row-len 1024 --rows 512,1024,2048 >> $S 2>>gpurun_out/sweep.err
python scripts/shard_sweep.py --row-len 512 --rows 256,512,1024 >> $S 2>>gpurun_out/sweep.err
ZIPGPU_WS_TOPS=0 python scripts/shard_sweep.py --row-len 512 --rows 256,512,1024 >> $S 2>>gpurun_out/sweep.err
cat $S | cut -c1-230
timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
