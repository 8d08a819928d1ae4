S=gpurun_out/r2_pdl.jsonl; : > $S
for P in "" 1 "" 1; do
  export ZIPGPU_NO_PDL=$P; [ -z "$P" ] && unset ZIPGPU_NO_PDL
  python scripts/shard_sweep.py --rows 512,4096 --reps 30 >> $S 2>>gpurun_out/sweep.err
  python scripts/shard_sweep.py --row-len 1024 --rows 1024 --reps 50 >> $S 2>>gpurun_out/sweep.err
  python scripts/shard_sweep.py --row-len 256 --rows 256 --reps 100 >> $S 2>>gpurun_out/sweep.err
  python scripts/shard_sweep.py --row-len 64 --rows 64 --reps 100 >> $S 2>>gpurun_out/sweep.err
done
unset ZIPGPU_NO_PDL
cat $S | cut -c1-210
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
