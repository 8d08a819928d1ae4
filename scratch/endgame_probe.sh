for q in 0 1 2 3 4 6 8 16; do echo -n "endgame_q=$q: "; ZIPGPU_ENDGAME_Q=$q python bench.py --kernels-only --steps 10 --warmup 3 2>/dev/null | cut -c60-190; done
