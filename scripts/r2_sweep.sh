S=gpurun_out/r2_ab4.jsonl; : > $S
for V in "" pfleaf pfstack pfboth ""; do
  LIB=""; [ -n "$V" ] && LIB=$PWD/zinc_b200/libzipgpu_$V.so
  ZIPGPU_LIB=$LIB python scripts/shard_sweep.py --rows 512,4096 --reps 30 | sed "s/\"knobs\"/\"variant\": \"$V\", \"knobs\"/" >> $S 2>>gpurun_out/sweep.err
done
for V in pfleaf pfstack pfboth; do ZIPGPU_LIB=$PWD/zinc_b200/libzipgpu_$V.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large" 2>&1 | tail -1; done
cat $S | cut -c1-175
