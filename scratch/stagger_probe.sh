for ns in 0 10000 20000 40000 60000 100000 200000; do echo -n "stagger=$ns: "; ZIPGPU_FUSE_STAGGER_NS=$ns python bench.py --kernels-only --steps 10 --warmup 3 | cut -c60-180; done
