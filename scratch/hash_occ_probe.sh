for b in 2 3 4 5; do echo "minb=$b"; ZIPGPU_HASH_MINB=$b python bench.py --kernels-only --steps 10 --warmup 3 | cut -c60-200; done
