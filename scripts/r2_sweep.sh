timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests7.log; tail -4 gpurun_out/r2_tests7.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
