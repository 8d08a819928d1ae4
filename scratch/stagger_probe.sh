for ns in 0 20000 40000 70000 100000; do echo -n "stagger=$ns: "; ZIPGPU_STAGGER_NS=$ns python bench.py --kernels-only --steps 10 --warmup 3 2>/dev/null | cut -c60-200; done
