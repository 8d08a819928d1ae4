S=gpurun_out/r2_ab3.jsonl; : > $S
for V in "" onecomp ""; do
  LIB=""; [ -n "$V" ] && LIB=$PWD/zinc_b200/libzipgpu_$V.so
  ZIPGPU_LIB=$LIB python scripts/shard_sweep.py --rows 512,4096 --reps 30 | sed "s/\"knobs\"/\"variant\": \"$V\", \"knobs\"/" >> $S 2>>gpurun_out/sweep.err
done
ZIPGPU_LIB=$PWD/zinc_b200/libzipgpu_onecomp.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large" 2>&1 | tail -2
cat $S | cut -c1-200
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests5.log; tail -3 gpurun_out/r2_tests5.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
